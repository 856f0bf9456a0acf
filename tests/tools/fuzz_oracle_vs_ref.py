"""Differential fuzzing of the ORACLE (oracle/wmbus_oracle.c, the restatement every parity test compares with) against the
UNMODIFIED reference binary (oracle/_ref/rtl_wmbus, compiled from /root/reference by oracle/Makefile): the same random
captures and flag sets as tests/tools/fuzz_hostsim.py (tests/fuzz_cases.py draws them), so that the chain
reference binary == oracle == product is closed on random input and not only on the fixtures.  Needs the reference
binary, i.e. the container that holds /root/reference.
    python tests/tools/fuzz_oracle_vs_ref.py [seconds] [seed] [tone]     one line per case; exits 1 at the first mismatch"""
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import orc, fuzz_cases

if not orc.ref_binary():
    sys.exit("oracle/_ref/rtl_wmbus is missing (make -C oracle ref, where /root/reference exists)")
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 600.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
tone = len(sys.argv) > 3 and sys.argv[3] == "tone"           # every case with a CW carrier inside the channel, noise sigma <= 3 (as fuzz_hostsim.py)
rng_tone = np.random.default_rng(seed + 1000003)
t_end = time.time() + budget
k = lines = 0
while time.time() < t_end:
    k += 1
    c = fuzz_cases.draw_case(rng)
    if c["prefilter"]:
        continue                                             # the reference cannot be switched into that mode
    if tone:
        c["cw"] = (float(rng_tone.uniform(-60e3, 60e3)), float(rng_tone.uniform(10, 60)))
        c["sigma"] = float(rng_tone.choice([1.0, 3.0]))
    cu8 = fuzz_cases.build_capture(c)
    want = orc.ref_lines(cu8, c["flags"])
    got = [orc.blank_ts(l) for l in orc.run_lines(cu8, orc.opts_from_flags(c["flags"]))]
    lines += len(want)
    print("case %d (seed %d) %s flags=%r n=%d sigma=%g cw=%r lines=%d" % (k, seed, "ok" if got == want else "MISMATCH",
                                                                        c["flags"], c["n"], c["sigma"], c["cw"], len(want)), flush=True)
    if got != want:
        print("oracle", len(got), "reference", len(want)); sys.exit(1)
print("done", k, "cases,", lines, "lines")
