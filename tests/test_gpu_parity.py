"""`-m gpu`: parity of the CUDA path (libwmbus_b200.so on a B200) with the CPU oracle and the
reference's goldens, through the C ABI.  Bit-exact everywhere: the path is integer work plus fp32
arithmetic restated operation by operation (tolerance 0; the RSSI columns are integers)."""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np
import pytest

import pipeline_checks as pc
from conftest import ROOT, load_fixture

pytestmark = pytest.mark.gpu


def test_device_atan2f_and_discriminator_bit_exact(pkg, gpu_lib, orc_mod):
    """The demod kernel's dphi output exercises the device atan2f on every sample; a capture of
    full-scale uniform noise covers all octants and reduction ranges."""
    rng = np.random.default_rng(9)
    cu8 = rng.integers(0, 256, 1 << 22).astype(np.uint8)
    pc.check_stages(pkg, gpu_lib, cu8, "")
    pc.check_stages(pkg, gpu_lib, cu8, "-d 1")
    edge = np.tile(np.array([127, 128, 0, 255, 128, 127, 255, 0, 127, 127, 128, 128], np.uint8), 1 << 16)[:1 << 19]
    pc.check_stages(pkg, gpu_lib, np.ascontiguousarray(edge), "")


def test_device_arith_operand_by_operand(pkg, gpu_lib, orc_mod):
    """device atan2f (bounded / general), slow-path-free IEEE division and sqrt, discriminator scaling invariance"""
    import test_device_arith
    test_device_arith.check_device_arith(pkg, gpu_lib, orc_mod, 400000)


def test_stage_outputs_bit_exact(pkg, gpu_lib):
    pc.check_stages(pkg, gpu_lib, load_fixture("excerpt_samples2_a.cu8"), "")
    pc.check_stages(pkg, gpu_lib, load_fixture("excerpt_samples2_a.cu8"), "-a")
    pc.check_stages(pkg, gpu_lib, load_fixture("excerpt_issue48_2m4.cu8"), "-d 3 -s -o")
    pc.check_stages(pkg, gpu_lib, load_fixture("synth_mixed_2m4_shift.cu8"), "-d 3 -s")
    pc.check_stages(pkg, gpu_lib, load_fixture("synth_mixed_1m6.cu8"), "-d 4 -s")


def test_bitsync_stage_taps(pkg, gpu_lib):
    """Each bit-sync stage on its own: a6/a7 slicer (+DC block), a9 clock signs, a10 lock strobes, a11 deglitch +
    a12 run-length events, a13 time2 shift register / access code -- every sample and every event vs the oracle."""
    cu8 = load_fixture("synth_mixed_1m6.cu8")
    counts = pc.check_bitsync_stages(pkg, gpu_lib, cu8, "")
    assert all(n > 1000 and syncs >= 1 for n, syncs, _ in counts.values()) and counts[(0, 0)][2] > 100
    pc.check_bitsync_stages(pkg, gpu_lib, cu8, "-o")
    pc.check_bitsync_stages(pkg, gpu_lib, cu8, "-r 0")
    pc.check_bitsync_stages(pkg, gpu_lib, cu8, "-t 0 -a")
    pc.check_bitsync_stages(pkg, gpu_lib, load_fixture("excerpt_samples2_a.cu8"), "-p S")
    pc.check_bitsync_stages(pkg, gpu_lib, load_fixture("excerpt_issue48_2m4.cu8"), "-d 3 -s -o")
    pc.check_bitsync_stages(pkg, gpu_lib, load_fixture("excerpt_issue47_c1.cu8"), "")
    rng = np.random.default_rng(11)
    pc.check_bitsync_stages(pkg, gpu_lib, rng.integers(0, 256, 1 << 22).astype(np.uint8), "")      # full-scale noise


def test_golden_lines_all_flags(pkg, gpu_lib, golden_lines):
    assert pc.check_golden(pkg, gpu_lib, golden_lines) > 200


def test_full_reference_captures(pkg, gpu_lib, orc_mod):
    sdir = os.path.join(ROOT, "oracle", "_ref", "samples")
    if not os.path.isdir(sdir):
        pytest.skip("reference sample captures not shipped (oracle/_ref/samples)")
    for name in sorted(os.listdir(sdir)):
        cu8 = np.fromfile(os.path.join(sdir, name), np.uint8)
        flagsets = ["", "-v", "-o", "-a -o", "-r 0", "-t 0"] if "1M6" in name else ["-d 3 -s -o -v", "-d 3 -s"]
        for flags in flagsets:
            got, _ = pc.run_lines(pkg, gpu_lib, cu8, flags)
            assert got == pc.oracle_lines(cu8, flags), (name, flags)


def test_geometry_and_push_invariance(pkg, gpu_lib):
    cu8 = load_fixture("synth_mixed_1m6.cu8")
    variants = [dict(chunk_samples=1024, warmup_samples=32768, max_batch_mib=1),
                dict(chunk_samples=4096, warmup_samples=256, max_batch_mib=1),
                dict(chunk_samples=2048, warmup_samples=1024),
                dict(pushes=[4096] * 7 + [12288, 4096 * 33, 1, 5000, 8191]),
                dict(pushes=[100000, 300000], max_batch_mib=1, chunk_samples=8192, warmup_samples=512)]
    assert pc.check_invariance(pkg, gpu_lib, cu8, "-v", variants) > 0
    pc.check_invariance(pkg, gpu_lib, cu8, "-v -o", [dict(chunk_samples=2048, warmup_samples=2048, max_batch_mib=1)])
    pc.check_invariance(pkg, gpu_lib, load_fixture("synth_mixed_2m4_shift.cu8"), "-v -d 3 -s",
                        [dict(chunk_samples=1024, warmup_samples=512, max_batch_mib=1, pushes=[4096 * 3] * 9 + [7, 4096 * 50])])


def test_manual_frame_api(pkg, gpu_lib):
    pc.check_manual_frames(pkg, gpu_lib, load_fixture("synth_mixed_1m6.cu8"), "-v")


def test_second_reset_rule_falls_back_to_monolithic_lanes(pkg, gpu_lib):
    pc.check_type2_fallback(pkg, gpu_lib)


def test_degenerate_inputs(pkg, gpu_lib):
    pc.check_degenerate(pkg, gpu_lib)


def test_device_push_with_ragged_tail(pkg, gpu_lib):
    import torch
    pc.check_device_push_ragged(pkg, gpu_lib, to_device=lambda a: torch.from_numpy(a).cuda())


def test_sample_index_wrap_at_2_pow_40(pkg, gpu_lib):
    pc.check_sample_index_wrap(pkg, gpu_lib)


def test_synthetic_64mib_vs_oracle_and_device_input(pkg, gpu_lib):
    """A 64 MiB synthetic capture (all telegram types), generated on the GPU, decoded (a) from host memory,
    (b) from device memory in one batch, (c) in 8 MiB batches: identical lines, equal to the oracle's."""
    import torch
    synth = importlib.import_module("rtl-wmbus_b200.synth")
    em = synth.default_emitters("mixed")
    buf, plan = synth.synth_capture(64 << 20, emitters=em, seed=123, device="cuda")
    torch.cuda.synchronize()
    host = buf.cpu().numpy()
    want = pc.oracle_lines(host, "-v")
    assert len(want) > 400
    a, _ = pc.run_lines(pkg, gpu_lib, host, "-v")
    assert a == want
    with pkg.WmbusB200("-v", lib=gpu_lib) as ctx:
        b = ctx.process_device(buf.data_ptr(), buf.numel(), flush=True)
    assert b == want
    with pkg.WmbusB200("-v", lib=gpu_lib, max_batch_mib=8) as ctx:
        c = ctx.process_device(buf.data_ptr(), buf.numel(), flush=True)
        assert ctx.stats().batches == 8
    assert c == want
    # every planted telegram that the reference algorithm can decode is there, byte for byte
    good = {l.split(";")[8] for l in a if l.split(";")[2] == "1"}
    planted = {em[p.emitter].expected_fields(p.k)[2] for p in plan}
    assert len(good & planted) > 0.7 * len(planted)      # overlapping emitters collide by design


def test_cli_drop_in(pkg, gpu_lib, golden_lines):
    """The C host program keeps the reference's stdin -> stdout contract."""
    exe = os.path.join(os.path.dirname(pkg.library_path()), "rtl_wmbus_b200")
    for name, flags in [("excerpt_samples2_a.cu8", "-v"), ("synth_mixed_1m6.cu8", ""), ("excerpt_issue48_2m4.cu8", "-d 3 -s -o")]:
        data = load_fixture(name).tobytes()
        r = subprocess.run([exe] + flags.split(), input=data, capture_output=True, check=True)
        import orc
        got = [orc.blank_ts(l) for l in r.stdout.decode().split("\n") if l]
        assert got == golden_lines[name][flags]
        # timestamps look like the reference's
        for l in r.stdout.decode().split("\n"):
            if l:
                ts = l.split(";")[4 if flags.startswith("-v") else 3]
                assert len(ts) == 26 and ts[4] == "-" and ts[19] == "."


def test_table_overflow_costs_lines_not_the_stream(pkg, gpu_lib):
    pc.check_overflow_degrades(pkg, gpu_lib)


def test_many_carriers_per_capture(pkg, gpu_lib):
    import torch
    pc.check_carriers(pkg, gpu_lib, to_device=lambda a: torch.from_numpy(a).cuda())


def test_cw_interferer_refutes_lanes_not_lines(pkg, gpu_lib):
    pc.check_cw_interferer(pkg, gpu_lib)
