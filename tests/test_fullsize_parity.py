"""`-m gpu`: BASELINE.json's configurations at FULL size against the unmodified reference binary.

Each test synthesises a 1 GiB capture on the GPU, decodes it through the C ABI (device-resident input, one batch),
writes the same bytes to a scratch file and runs oracle/_ref/rtl_wmbus over it on all host cores in time chunks
(tests/ref_chunked.py: halo + prefix rule, validated against whole-capture runs in tests/test_ref_chunked.py).
The comparison is the whole line list, in order, timestamps blanked: MODE;CRC_OK;3OUTOF6OK;TS;PACKET_RSSI;
CURRENT_RSSI;LINK_LAYER_IDENT_NO;DATAGRAM, with the -v prefix so that the four bit-sync streams are told apart.
Tolerance 0 (the RSSI columns are integers)."""
import hashlib
import importlib
import os
import time

import numpy as np
import pytest

import orc
import ref_chunked as rc

pytestmark = pytest.mark.gpu

GIB = 1 << 30


def _decode_and_compare(pkg, gpu_lib, emitters, flags, d, fs, seed, shift=0.0, nbytes=GIB, extra_check=None):
    import torch
    if orc.ref_binary() is None:
        pytest.skip("oracle/_ref/rtl_wmbus not shipped")
    synth = importlib.import_module("rtl-wmbus_b200.synth")
    em = synth.default_emitters(emitters)
    t0 = time.perf_counter()
    cap, plan = synth.synth_capture(nbytes, fs=fs, emitters=em, seed=seed, device="cuda", center_shift_hz=shift)
    torch.cuda.synchronize()
    with pkg.WmbusB200(flags, lib=gpu_lib, max_batch_mib=nbytes >> 20) as ctx:
        got = ctx.process_device(cap.data_ptr(), nbytes, flush=True)
        st = ctx.stats()
    t1 = time.perf_counter()
    path = os.path.join(rc.scratch_dir(nbytes), f"wmbus_fullsize_{os.getpid()}.cu8")
    try:
        cap.cpu().numpy().tofile(path)
        t2 = time.perf_counter()
        want = rc.ref_lines_chunked(path, nbytes, flags, d=d)
    finally:
        if os.path.exists(path):
            os.unlink(path)
    t3 = time.perf_counter()
    print(f"[fullsize {flags!r}] synth+gpu {t1 - t0:.1f} s, file {t2 - t1:.1f} s, reference on {rc.host_cpus()} cpus "
          f"{t3 - t2:.1f} s, {len(want)} lines, {st.lanes_rerun} lanes re-run")
    diff = [(i, a, b) for i, (a, b) in enumerate(zip(got, want)) if a != b]
    assert len(got) == len(want) and len(diff) == 0, (len(got), len(want), diff[:3])
    assert len(want) > 100
    if extra_check:
        extra_check(cap, em, plan, got)
    del cap
    return got


def test_config2_t1x2_1gib(pkg, gpu_lib):
    """config 2: 1 GiB, 1.6 MS/s, two T1 emitters, T1+C1 chain (-p S); plus size-independent properties:
    batch-size invariance (sha of the lines), every CRC-ok datagram is a planted one, the strong emitter's
    telegrams are all recovered."""
    def props(cap, em, plan, got):
        with pkg.WmbusB200("-v -p S", lib=gpu_lib, max_batch_mib=128) as ctx:
            small = ctx.process_device(cap.data_ptr(), GIB, flush=True)
            assert ctx.stats().batches == 8
        sha = lambda ls: hashlib.sha256("\n".join(ls).encode()).hexdigest()
        assert sha(small) == sha(got)
        planted = {em[p.emitter].expected_fields(p.k)[2] for p in plan}
        ok = [l for l in got if l.split(";")[2] == "1"]
        assert all(l.split(";")[8] in planted for l in ok)
        strong = {em[0].expected_fields(p.k)[2] for p in plan if p.emitter == 0}
        assert len(strong - {l.split(";")[8] for l in ok}) <= 0.02 * len(strong)
    _decode_and_compare(pkg, gpu_lib, "t1x2", "-v -p S", 2, 1.6e6, 0xB2000020, extra_check=props)


def test_config3_s1_1gib(pkg, gpu_lib):
    """config 3: 1 GiB, 1.6 MS/s input (the S1 chain itself runs at 800 kS/s), S1 emitters, -p T"""
    _decode_and_compare(pkg, gpu_lib, "s1", "-v -p T", 2, 1.6e6, 0xB2000030)


def test_default_flags_both_chains_1gib(pkg, gpu_lib):
    """the drop-in's default: both chains, both bit-sync algorithms, dense T1/C1-A/C1-B/S1 traffic"""
    _decode_and_compare(pkg, gpu_lib, "mixed", "-v", 2, 1.6e6, 0xB2000021)


def test_config4_d3_1gib(pkg, gpu_lib):
    """config 4's signal: 2.4 MS/s, -d 3 (general front end)"""
    _decode_and_compare(pkg, gpu_lib, "mixed", "-v -d 3", 3, 2.4e6, 0xB2000040)


def test_config4_d3_mixer_1gib(pkg, gpu_lib):
    """... and its -s variant: capture centred on 868.625 MHz, T1/C1 at +325 kHz, S1 at -325 kHz"""
    _decode_and_compare(pkg, gpu_lib, "mixed", "-v -d 3 -s", 3, 2.4e6, 0xB2000041, shift=325e3)
