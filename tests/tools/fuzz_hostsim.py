"""Differential fuzzing of the library's host logic and kernel phase functions on the CPU build (tests/hostsim) against
the oracle: random captures (noise level, emitters, interferers, dead air), random flags, lane geometry and push sizes
(tests/fuzz_cases.py draws them).
    python tests/tools/fuzz_hostsim.py [seconds] [seed] [tone|manual]     prints one line per case; exits 1 at the first mismatch
`tone` forces every case into the regime the first campaign found: a CW carrier inside the channel,
noise sigma <= 3 -- the run-length tracker's bit length collapses, event buffers and rings fill."""
import importlib, sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import orc, pipeline_checks as pc, fuzz_cases
from conftest import HOSTSIM_SO
pkg = importlib.import_module("rtl-wmbus_b200")
lib = pkg.load_library(HOSTSIM_SO)
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 600.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
tone = len(sys.argv) > 3 and sys.argv[3] == "tone"
manual = len(sys.argv) > 3 and sys.argv[3] == "manual"       # the caller drives the framers: wmb_push / wmb_poll / wmb_decode_frames


def run_manual(cu8, flags, pushes, tuning):
    data = np.ascontiguousarray(cu8, np.uint8)
    lines, off = [], 0
    with pkg.WmbusB200(flags, lib=lib, manual_frames=1, **tuning) as ctx:
        for n in list(pushes or []) + [1 << 19] * (len(data) // (1 << 19) + 1):
            n = min(n, len(data) - off)
            if n <= 0:
                break
            ctx.push(data.ctypes.data + off, n); off += n
            arr, k = ctx.poll(flush=False)
            ctx.decode_frames(arr, k)
            lines += ctx.take_lines()
        arr, k = ctx.poll(flush=True)
        ctx.decode_frames(arr, k)
        lines += ctx.take_lines()
        return lines, ctx.stats()
rng_tone = np.random.default_rng(seed + 1000003)             # (its own generator: the cases keep their numbers)
t_end = time.time() + budget
k = 0
while time.time() < t_end:
    k += 1
    c = fuzz_cases.draw_case(rng)
    if tone:
        c["cw"] = (float(rng_tone.uniform(-60e3, 60e3)), float(rng_tone.uniform(10, 60)))
        c["sigma"] = float(rng_tone.choice([1.0, 3.0]))
    print("start %d flags=%r d=%d n=%d sigma=%g tuning=%r pushes=%r" % (k, c["flags"], c["d"], c["n"], c["sigma"], c["tuning"], c["pushes"]), flush=True)
    cu8 = fuzz_cases.build_capture(c)
    o = orc.opts_from_flags(c["flags"]); o.prefilter = c["prefilter"]
    want = [orc.blank_ts(l) for l in orc.run_lines(cu8, o)]
    if manual:
        got, st = run_manual(cu8, c["flags"], c["pushes"], c["tuning"])
    else:
        got, st = pc.run_lines(pkg, lib, cu8, c["flags"], pushes=c["pushes"], **c["tuning"])
    ok = got == want
    lost = 0
    if not ok and st.overflow_batches:                       # a device table was full: lines may be missing, none may be invented
        it = iter(want)
        ok = all(any(l == w for w in it) for l in got)
        lost = len(want) - len(got)
    print("case %d (seed %d) %s lines=%d rerun=%d fallbacks=%d overflow_batches=%d lost=%d" % (
        k, seed, "ok" if ok else "MISMATCH", len(want), st.lanes_rerun, st.rl_fallbacks, st.overflow_batches, lost), flush=True)
    if not ok:
        print("got", len(got), "want", len(want)); sys.exit(1)
print("done", k, "cases")
