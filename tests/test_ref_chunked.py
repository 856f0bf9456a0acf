"""`not gpu`: the chunked full-size parity helper (tests/ref_chunked.py) reproduces whole-capture runs of the
unmodified reference binary, line for line and in order -- so the `-m gpu` full-size tests may use it as the oracle
for 1 GiB captures."""
import importlib
import os

import numpy as np
import pytest

import orc
import ref_chunked as rc

pytestmark = pytest.mark.skipif(orc.ref_binary() is None, reason="oracle/_ref/rtl_wmbus not built")


def _capture_file(tmp_path, nbytes, emitters, seed, fs=1.6e6, shift=0.0):
    synth = importlib.import_module("rtl-wmbus_b200.synth")
    buf, _ = synth.synth_capture(nbytes, fs=fs, emitters=synth.default_emitters(emitters), seed=seed,
                                 center_shift_hz=shift)
    p = os.path.join(tmp_path, "cap.cu8")
    buf.numpy().tofile(p)
    return p


@pytest.mark.parametrize("flags,d,emitters,fs,shift", [
    ("-v", 2, "mixed", 1.6e6, 0.0),
    ("-p S", 2, "t1x2", 1.6e6, 0.0),
    ("-v -d 3 -s -o", 3, "mixed", 2.4e6, 325e3),
])
def test_chunked_reference_equals_whole_run(tmp_path, flags, d, emitters, fs, shift):
    n = 12 << 20
    path = _capture_file(str(tmp_path), n, emitters, 0xB2000061, fs, shift)
    want = rc.ref_lines_whole(path, n, flags)
    assert len(want) > 8
    # five chunks, halo of 2^18 decimated samples (shorter than a chunk, so the halo logic is exercised)
    got = rc.ref_lines_chunked(path, n, flags, d=d, procs=4, chunks=5, halo_m=1 << 18)
    assert got == want


def test_chunk_plan_alignment_and_ragged_tail():
    plan = rc.chunk_plan((64 << 20) + 1234, 3, 7, 1 << 18)
    assert plan[0][0] == 0 and plan[0][1] == 0
    assert plan[-1][2] == (64 << 20)                     # whole 4096-byte items only
    for (h, b, e), nxt in zip(plan, plan[1:] + [None]):
        assert h % (4096 * 3) == 0 and b % (4096 * 3) == 0 and h <= b < e
        if nxt:
            assert nxt[1] == e
    assert rc.host_cpus() >= 1
