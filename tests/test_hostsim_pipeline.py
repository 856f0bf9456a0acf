"""`not gpu`: the library's host logic (batching, history carry-over, lane verification/re-run,
candidate hand-over, framers, ordering) and the kernels' phase functions, executed through the C ABI
of the CPU simulation build (tests/hostsim -- test infrastructure, not a product path)."""
import numpy as np
import pytest

import pipeline_checks as pc
from conftest import load_fixture


def test_golden_lines_all_flags(pkg, hostsim_lib, golden_lines):
    assert pc.check_golden(pkg, hostsim_lib, golden_lines) > 200


def test_stage_outputs_bit_exact(pkg, hostsim_lib):
    pc.check_stages(pkg, hostsim_lib, load_fixture("excerpt_samples2_a.cu8"), "")
    pc.check_stages(pkg, hostsim_lib, load_fixture("excerpt_issue48_2m4.cu8"), "-d 3 -s -o")
    pc.check_stages(pkg, hostsim_lib, load_fixture("synth_mixed_1m6.cu8")[:1 << 19], "-a -d 1")


def test_geometry_and_push_invariance(pkg, hostsim_lib):
    cu8 = load_fixture("synth_mixed_1m6.cu8")
    variants = [dict(chunk_samples=1024, warmup_samples=32768, max_batch_mib=1),
                dict(chunk_samples=4096, warmup_samples=256, max_batch_mib=1),      # forces lane re-runs
                dict(chunk_samples=2048, warmup_samples=1024),
                dict(pushes=[4096] * 7 + [12288, 4096 * 33, 1, 5000, 8191]),
                dict(pushes=[100000, 300000], max_batch_mib=1, chunk_samples=8192, warmup_samples=512)]
    reruns = pc.check_invariance(pkg, hostsim_lib, cu8, "-v", variants)
    assert reruns > 0, "the short warm-up variants must exercise the re-run path"
    pc.check_invariance(pkg, hostsim_lib, cu8, "-v -o", [dict(chunk_samples=2048, warmup_samples=2048, max_batch_mib=1)])
    cu8s = load_fixture("synth_mixed_2m4_shift.cu8")
    pc.check_invariance(pkg, hostsim_lib, cu8s, "-v -d 3 -s",
                        [dict(chunk_samples=1024, warmup_samples=512, max_batch_mib=1, pushes=[4096 * 3] * 9 + [7, 4096 * 50])])


def test_manual_frame_api(pkg, hostsim_lib):
    pc.check_manual_frames(pkg, hostsim_lib, load_fixture("synth_mixed_1m6.cu8"), "-v")


def test_second_reset_rule_falls_back_to_monolithic_lanes(pkg, hostsim_lib):
    pc.check_type2_fallback(pkg, hostsim_lib)


def test_degenerate_inputs(pkg, hostsim_lib):
    pc.check_degenerate(pkg, hostsim_lib)


def test_error_paths(pkg, hostsim_lib):
    import ctypes as C
    o = pkg.opts_from_flags(hostsim_lib, "-d 200")
    ctx = C.c_void_p()
    assert hostsim_lib.wmb_create(C.byref(o), 0, C.byref(ctx)) == -1        # WMB_E_INVAL
    assert b"decimation" in hostsim_lib.wmb_last_error()
    o = pkg.opts_from_flags(hostsim_lib, "-d 26")                           # one more than a block's shared memory holds
    assert hostsim_lib.wmb_create(C.byref(o), 0, C.byref(ctx)) == -1 and b"max 25" in hostsim_lib.wmb_last_error()
    o = pkg.opts_from_flags(hostsim_lib, "-d 25")
    assert hostsim_lib.wmb_create(C.byref(o), 0, C.byref(ctx)) == 0
    hostsim_lib.wmb_destroy(ctx)
    o = pkg.opts_from_flags(hostsim_lib, "")
    assert hostsim_lib.wmb_create(C.byref(o), 7, C.byref(ctx)) == -2        # WMB_E_NODEVICE


def test_sample_index_wrap_at_2_pow_40(pkg, hostsim_lib):
    pc.check_sample_index_wrap(pkg, hostsim_lib)


def test_bitsync_stage_taps(pkg, hostsim_lib):
    """a6/a7 slicer (+DC block), a9 clock signs, a10 lock strobes, a11-a13 bit events of all four streams"""
    cu8 = load_fixture("synth_mixed_1m6.cu8")
    counts = pc.check_bitsync_stages(pkg, hostsim_lib, cu8, "")
    assert all(n > 1000 and syncs >= 1 for n, syncs, _ in counts.values()) and counts[(0, 0)][2] > 100
    pc.check_bitsync_stages(pkg, hostsim_lib, cu8, "-o")
    pc.check_bitsync_stages(pkg, hostsim_lib, load_fixture("excerpt_issue48_2m4.cu8"), "-d 3 -s -o")


def test_device_push_with_ragged_tail(pkg, hostsim_lib):
    pc.check_device_push_ragged(pkg, hostsim_lib)


def test_table_overflow_costs_lines_not_the_stream(pkg, hostsim_lib):
    pc.check_overflow_degrades(pkg, hostsim_lib)


def test_many_carriers_per_capture(pkg, hostsim_lib):
    pc.check_carriers(pkg, hostsim_lib)


def test_cw_interferer_refutes_lanes_not_lines(pkg, hostsim_lib):
    pc.check_cw_interferer(pkg, hostsim_lib)


def test_dormant_prefilter_front_end(pkg, hostsim_lib):
    pc.check_prefilter(pkg, hostsim_lib)


def test_lane_event_overflow_costs_bits_not_the_stream(pkg, hostsim_lib):
    pc.check_lane_event_overflow(pkg, hostsim_lib)


@pytest.mark.parametrize("order", [1, 2])
def test_results_do_not_depend_on_the_order_of_simulated_threads(hostsim_lib, order):
    """On the device the threads of a phase, the lanes and the blocks of a launch run concurrently; the CPU build runs
    them one after another, in index order by default.  WMB_HOSTSIM_ORDER=1 runs every phase backwards, 2 in a scrambled
    order (tests/hostsim/hostsim_launch.inl: hs_for): a missing barrier between two phases of the demod block, or a lane
    that reads what another lane of the same launch writes, would change a result.  The core of this file again, in a
    process of its own per order (the variable is read once)."""
    import os, subprocess, sys
    from conftest import ROOT
    env = dict(os.environ, WMB_HOSTSIM_ORDER=str(order))
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "not gpu", "-p", "no:cacheprovider",
                        os.path.join(ROOT, "tests", "test_hostsim_pipeline.py"), os.path.join(ROOT, "tests", "test_framer.py"),
                        "-k", "golden or stage or invariance or second_reset or prefilter or table_overflow or device_framer or carriers"],
                       env=env, cwd=ROOT, capture_output=True, timeout=1200)
    assert r.returncode == 0, r.stdout.decode()[-3000:]
    assert b"passed" in r.stdout


def test_decimations_up_to_the_limit(pkg, hostsim_lib):
    """-d beyond what an RTL-SDR delivers (8: 6.4 MS/s, 25: 20 MS/s = WMB_MAX_DECIMATION, where a tile's raw samples still
    fit a block's shared memory): demod stages sample by sample and the lines, one shot and in ragged pushes."""
    import importlib
    synth = importlib.import_module("rtl-wmbus_b200.synth")
    for d, flags in ((8, "-v -d 8 -s -o"), (25, "-v -d 25")):
        cap, _ = synth.synth_capture(4096 * d * 96, fs=800e3 * d, emitters=synth.default_emitters("mixed"), seed=1234 + d)
        cu8 = np.ascontiguousarray(cap.numpy())
        pc.check_stages(pkg, hostsim_lib, cu8, flags)
        want = pc.oracle_lines(cu8, flags)
        got, _ = pc.run_lines(pkg, hostsim_lib, cu8, flags, max_batch_mib=1, pushes=[12345, 1 << 18, 4096 * d * 3])
        assert got == want and (len(want) >= 10 or "-s" in flags)
