"""Parity checks shared by the CPU-simulation tests (host logic, `not gpu`) and the GPU tests proper.
Every check drives the product through its C ABI and compares with the CPU oracle / reference goldens."""
import ctypes as C
import importlib

import numpy as np

import orc
from conftest import load_fixture


def run_lines(pkg, lib, cu8, flags, pushes=None, **tuning):
    """Decode a capture through the C ABI.  pushes: list of byte counts to split the input into."""
    data = np.ascontiguousarray(cu8, np.uint8)
    with pkg.WmbusB200(flags, lib=lib, **tuning) as ctx:
        if pushes is None:
            lines = ctx.process(data.ctypes.data, len(data), flush=True)
        else:
            lines, off = [], 0
            for n in pushes:
                n = min(n, len(data) - off)
                lines += ctx.process(data.ctypes.data + off, n, flush=False)
                off += n
            if off < len(data):
                lines += ctx.process(data.ctypes.data + off, len(data) - off, flush=False)
            lines += ctx.process(0, 0, flush=True)
        st = ctx.stats()
    return lines, st


def oracle_lines(cu8, flags):
    return [orc.blank_ts(l) for l in orc.run_lines(cu8, orc.opts_from_flags(flags))]


def check_golden(pkg, lib, golden_lines, only=None, **tuning):
    total = 0
    for name, per_flags in golden_lines.items():
        if only and name not in only:
            continue
        cu8 = load_fixture(name)
        for flags, want in per_flags.items():
            got, _ = run_lines(pkg, lib, cu8, flags, **tuning)
            assert got == want, (name, flags, len(got), len(want))
            total += len(want)
    return total


def check_stages(pkg, lib, cu8, flags, **tuning):
    """dphi (post-FIR) and (unsigned)rssi of the demod kernel vs the oracle, bit for bit."""
    o = orc.opts_from_flags(flags)
    gran = 4096 * max(1, o.decimation)          # one batch: the stage tap returns the last batch only
    data = np.ascontiguousarray(cu8[:len(cu8) // gran * gran], np.uint8)
    with pkg.WmbusB200(flags, lib=lib, **tuning) as ctx:
        ctx.process(data.ctypes.data, len(data), flush=True)
        for chain in (0, 1):
            if (chain == 0 and not o.t1c1_enabled) or (chain == 1 and not o.s1_enabled):
                continue
            st = orc.stages(data, o, chain)
            dphi, rssi = ctx.debug_stage(chain, st["M"])
            assert len(dphi) == st["M"]
            bad = np.nonzero(dphi.view(np.uint32) != st["fir"].view(np.uint32))[0]
            assert len(bad) == 0, (flags, chain, "dphi", bad[:5], dphi[bad[:5]], st["fir"][bad[:5]])
            want = st["rssi"].astype(np.uint32).astype(np.uint8)
            bad = np.nonzero(rssi != want)[0]
            assert len(bad) == 0, (flags, chain, "rssi", bad[:5])


def check_invariance(pkg, lib, cu8, flags, variants):
    """Lane/batch geometry and push granularity never change the output."""
    want = oracle_lines(cu8, flags)
    reruns = 0
    for v in variants:
        v = dict(v)
        pushes = v.pop("pushes", None)
        got, st = run_lines(pkg, lib, cu8, flags, pushes=pushes, **v)
        assert got == want, (flags, v, pushes, len(got), len(want))
        reruns += st.lanes_rerun
    return reruns


def check_manual_frames(pkg, lib, cu8, flags):
    want = oracle_lines(cu8, flags)
    data = np.ascontiguousarray(cu8, np.uint8)
    lines = []
    with pkg.WmbusB200(flags, lib=lib, manual_frames=1, max_batch_mib=1) as ctx:
        step = 1 << 19
        for off in range(0, len(data), step):
            n = min(step, len(data) - off)
            ctx.push(data.ctypes.data + off, n)
            arr, k = ctx.poll(flush=False)
            ctx.decode_frames(arr, k)
            lines += ctx.take_lines()
        arr, k = ctx.poll(flush=True)
        ctx.decode_frames(arr, k)
        lines += ctx.take_lines()
    assert lines == want


def degenerate_inputs():
    rng = np.random.default_rng(5)
    out = {
        "empty": np.zeros(0, np.uint8),
        "short": np.full(4095, 128, np.uint8),
        "one_item": rng.integers(0, 256, 4096).astype(np.uint8),
        "ragged": rng.integers(100, 156, 4096 * 5 + 1234).astype(np.uint8),
        "zeros": np.zeros(1 << 18, np.uint8),
        "const": np.full(1 << 18, 200, np.uint8),
        "noise": rng.integers(0, 256, 1 << 19).astype(np.uint8),
        "square": np.tile(np.repeat(np.array([[90, 160], [160, 90]], np.uint8), 40, axis=0).reshape(-1), 1700)[:1 << 18],
    }
    out["square"] = np.ascontiguousarray(out["square"][:len(out["square"]) // 4096 * 4096])
    return out


def check_degenerate(pkg, lib, flags_list=("-v", "-v -o -d 3 -s", "-v -a -d 1")):
    for name, cu8 in degenerate_inputs().items():
        for flags in flags_list:
            got, _ = run_lines(pkg, lib, cu8, flags, max_batch_mib=1)
            assert got == oracle_lines(cu8, flags), (name, flags)
