"""Differential fuzzing of the library's host logic and kernel phase functions on the CPU build (tests/hostsim) against
the oracle: random captures (noise level, emitters, carriers, interferers), random flags, lane geometry and push sizes.
    python tools/fuzz_hostsim.py [seconds] [seed]      prints one line per case; exits 1 at the first mismatch"""
import ctypes as C, importlib, sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import orc, pipeline_checks as pc
from conftest import HOSTSIM_SO
pkg = importlib.import_module("rtl-wmbus_b200"); synth = importlib.import_module("rtl-wmbus_b200.synth")
lib = pkg.load_library(HOSTSIM_SO)
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 600.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
t_end = time.time() + budget
case = 0
while time.time() < t_end:
    case += 1
    d = int(rng.choice([1, 2, 2, 2, 3, 4]))
    fs = 800e3 * d
    flags = ["-v"] if rng.random() < 0.8 else []
    if d != 2: flags += ["-d", str(d)]
    shift = 0.0
    if rng.random() < 0.3 and d >= 2: flags.append("-s"); shift = 325e3
    if rng.random() < 0.25: flags.append("-o")
    if rng.random() < 0.15: flags.append("-a")
    if rng.random() < 0.1: flags += ["-r", "0"]
    if rng.random() < 0.1: flags += ["-t", "0"]
    p = rng.random()
    if p < 0.15: flags += ["-p", "S"]
    elif p < 0.3: flags += ["-p", "T"]
    flags = " ".join(flags)
    n = int(rng.integers(40, 400)) * 4096 * d
    em = []
    for k in range(int(rng.integers(1, 5))):
        mode = str(rng.choice(["T1", "C1A", "C1B", "S1"]))
        em.append(synth.Emitter(mode, int(rng.integers(0, 99999999)) // 1 * 1 % 0x99999999 & 0x77777777, amp=float(rng.uniform(20, 100)),
                                offset_hz=float(rng.uniform(-15e3, 15e3)), l_field=int(rng.integers(10, 120)),
                                period_s=float(rng.uniform(0.05, 0.2)), start_s=float(rng.uniform(0.002, 0.05)), seed=int(rng.integers(1, 1000))))
    sigma = float(rng.choice([1.0, 3.0, 8.0, 20.0]))
    cap, _ = synth.synth_capture(n, fs=fs, emitters=em, seed=int(rng.integers(1, 1 << 30)), noise_sigma=sigma, center_shift_hz=shift)
    x = cap.numpy().astype(np.float64).reshape(-1, 2)
    if rng.random() < 0.3:                                    # CW interferer
        f = float(rng.uniform(-0.45, 0.45)) * fs
        tone = float(rng.uniform(10, 60)) * np.exp(2j * np.pi * f / fs * np.arange(len(x)))
        x[:, 0] += tone.real; x[:, 1] += tone.imag
    if rng.random() < 0.1:                                    # a stretch of dead air
        a = int(rng.integers(0, len(x) // 2)); x[a:a + len(x) // 4] = 127.0
    cu8 = np.ascontiguousarray(np.clip(np.round(x), 0, 255).astype(np.uint8).reshape(-1))
    tuning = {}
    if rng.random() < 0.7: tuning["max_batch_mib"] = int(rng.choice([1, 1, 2, 4]))
    if rng.random() < 0.5: tuning["chunk_samples"] = int(rng.choice([1024, 2048, 4096, 8192]))
    if rng.random() < 0.3: tuning["warmup_samples"] = int(rng.choice([256, 1024, 8192, 32768]))
    pre = 0
    if d == 2 and rng.random() < 0.15: pre = 1; tuning["prefilter"] = 1
    pushes = None
    if rng.random() < 0.6:
        pushes, left = [], len(cu8)
        while left > 0 and len(pushes) < 12:
            k = int(rng.choice([1, 4095, 4096, 12288, 100000, 1 << 18, 1 << 20])); k = min(k, left); pushes.append(k); left -= k
    o = orc.opts_from_flags(flags); o.prefilter = pre
    want = [orc.blank_ts(l) for l in orc.run_lines(cu8, o)]
    got, st = pc.run_lines(pkg, lib, cu8, flags, pushes=pushes, **tuning)
    ok = got == want
    print("case %d %s flags=%r d=%d n=%d sigma=%g tuning=%r pushes=%r lines=%d rerun=%d fallbacks=%d" % (
        case, "ok" if ok else "MISMATCH", flags, d, n, sigma, tuning, pushes if pushes is None else len(pushes), len(want), st.lanes_rerun, st.rl_fallbacks), flush=True)
    if not ok:
        np.save("/tmp/fuzz_fail.npy", cu8)
        print("got", len(got), "want", len(want)); sys.exit(1)
print("done", case, "cases")
