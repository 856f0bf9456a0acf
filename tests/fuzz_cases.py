"""Random test cases for the differential fuzzer (tests/tools/fuzz_hostsim.py) and for the regression tests that came out of
it: draw_case() only draws parameters (cheap, so a test can skip to case k of seed s), build_capture() makes the cu8."""
import importlib

import numpy as np


def draw_case(rng):
    """One case: flags, tuning, pushes and everything build_capture() needs.  The order of the draws is part of the
    contract (case k of seed s stays the same capture)."""
    c = {}
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
    n = int(rng.integers(40, 400)) * 4096 * d
    em = []
    for k in range(int(rng.integers(1, 5))):
        mode = str(rng.choice(["T1", "C1A", "C1B", "S1"]))
        em.append(dict(mode=mode, ident=int(rng.integers(0, 99999999)) // 1 * 1 % 0x99999999 & 0x77777777, amp=float(rng.uniform(20, 100)),
                       offset_hz=float(rng.uniform(-15e3, 15e3)), l_field=int(rng.integers(10, 120)),
                       period_s=float(rng.uniform(0.05, 0.2)), start_s=float(rng.uniform(0.002, 0.05)), seed=int(rng.integers(1, 1000))))
    sigma = float(rng.choice([1.0, 3.0, 8.0, 20.0]))
    cap_seed = int(rng.integers(1, 1 << 30))
    cw = None
    if rng.random() < 0.3:                                    # CW interferer
        cw = (float(rng.uniform(-0.45, 0.45)) * fs, float(rng.uniform(10, 60)))
    dead = None
    if rng.random() < 0.1:                                    # a stretch of dead air
        dead = int(rng.integers(0, n // 4))
    tuning = {}
    if rng.random() < 0.7: tuning["max_batch_mib"] = int(rng.choice([1, 1, 2, 4]))
    if rng.random() < 0.5: tuning["chunk_samples"] = int(rng.choice([1024, 2048, 4096, 8192]))
    if rng.random() < 0.3: tuning["warmup_samples"] = int(rng.choice([256, 1024, 8192, 32768]))
    pre = 0
    if d == 2:                                                # (one draw at d = 2 only, as when there was one dormant filter: the cases keep their numbers)
        u = rng.random()
        if u < 0.15: pre = 1 + int(u / 0.15 * 4); tuning["prefilter"] = pre
    pushes = None
    if rng.random() < 0.6:
        pushes, left = [], n
        while left > 0 and len(pushes) < 12:
            k = int(rng.choice([1, 4095, 4096, 12288, 100000, 1 << 18, 1 << 20])); k = min(k, left); pushes.append(k); left -= k
    c.update(d=d, fs=fs, flags=" ".join(flags), shift=shift, n=n, emitters=em, sigma=sigma, cap_seed=cap_seed, cw=cw, dead=dead,
             tuning=tuning, prefilter=pre, pushes=pushes)
    return c


def build_capture(c):
    synth = importlib.import_module("rtl-wmbus_b200.synth")
    em = [synth.Emitter(e["mode"], e["ident"], amp=e["amp"], offset_hz=e["offset_hz"], l_field=e["l_field"],
                        period_s=e["period_s"], start_s=e["start_s"], seed=e["seed"]) for e in c["emitters"]]
    cap, _ = synth.synth_capture(c["n"], fs=c["fs"], emitters=em, seed=c["cap_seed"], noise_sigma=c["sigma"], center_shift_hz=c["shift"])
    x = cap.numpy().astype(np.float64).reshape(-1, 2)
    if c["cw"]:
        f, amp = c["cw"]
        tone = amp * np.exp(2j * np.pi * f / c["fs"] * np.arange(len(x)))
        x[:, 0] += tone.real; x[:, 1] += tone.imag
    if c["dead"] is not None:
        a = c["dead"]; x[a:a + len(x) // 4] = 127.0
    return np.ascontiguousarray(np.clip(np.round(x), 0, 255).astype(np.uint8).reshape(-1))


def case(seed, k):
    """parameters of case k (1-based) of the fuzzer run with this seed"""
    rng = np.random.default_rng(seed)
    c = None
    for _ in range(k):
        c = draw_case(rng)
    return c


def draw_time_chunk_case(rng):
    """One case of tests/tools/fuzz_time_chunks.py: a capture long enough for 2-4 chunks with their halos, the number of
    chunks, the halo.  None where the draw is skipped (the order of the draws is part of the contract here too)."""
    c = draw_case(rng)
    if c["prefilter"]:
        return None
    c["n"] = int(rng.integers(600, 1500)) * 4096 * c["d"]
    c["world"] = int(rng.integers(2, 5)); c["halo"] = int(rng.choice([1 << 16, 1 << 17, 1 << 18]))
    return c


def time_chunk_case(seed, k):
    """parameters of case k (1-based, skipped draws counted) of the time-chunk fuzzer run with this seed"""
    rng = np.random.default_rng(seed)
    c = None
    for _ in range(k):
        c = draw_time_chunk_case(rng)
    return c
