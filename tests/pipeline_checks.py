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


def check_stages(pkg, lib, cu8, flags, orc_extra=None, **tuning):
    """dphi (post-FIR) and (unsigned)rssi of the demod kernel vs the oracle, bit for bit."""
    o = orc.opts_from_flags(flags)
    for k, v in (orc_extra or {}).items():
        setattr(o, k, v)
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


def type2_capture():
    """A capture that drives the T1/C1 run-length tracker to a bit length ~37 % above nominal (runs of
    11 samples) and then presents 5/6-sample runs: the reference's second reset rule
    (rtl_wmbus.c:756-762) fires, which the two-phase path must hand to the monolithic lanes.
    Real telegrams follow the pattern so that the fallback also has lines to get right."""
    import math
    synth = importlib.import_module("rtl-wmbus_b200.synth")
    n_iq = 1 << 19
    x = np.random.default_rng(4).normal(127.4, 4.0, (n_iq, 2))
    x[:1 << 16] = np.random.default_rng(8).normal(127.4, 4.0, (1 << 16, 2))   # seed found by search: 2 hits
    runs = [11] * 200 + [6] * 2 + [11] * 200 + [5] * 3 + [11] * 100
    lev, f = 0, []
    for r in runs:
        f += [50e3 if lev else -50e3] * (r * 2)
        lev ^= 1
    phase = 2 * math.pi * np.cumsum(np.array(f)) / 1.6e6
    s0 = 20000
    x[s0:s0 + len(f), 0] += 80 * np.cos(phase)
    x[s0:s0 + len(f), 1] += 80 * np.sin(phase)
    for k, (mode, ident, at) in enumerate([("T1", 0x71200023, 100000), ("C1A", 0x20338739, 200000),
                                           ("S1", 0x19131290, 300000), ("T1", 0x71200023, 420000)]):
        e = synth.Emitter(mode, ident, amp=80.0, offset_hz=3e3, l_field=0x19, seed=40 + k)
        b = synth.fsk_burst(e.chips(k), e.chip_rate, 1.6e6, e.dev_hz, e.offset_hz, e.amp)
        x[at:at + len(b)] += b
    return np.clip(np.round(x), 0, 255).astype(np.uint8).reshape(-1)


def check_type2_fallback(pkg, lib):
    cu8 = type2_capture()
    want = oracle_lines(cu8, "-v")
    assert len(want) >= 6
    got, st = run_lines(pkg, lib, cu8, "-v")
    assert got == want
    assert st.rl_fallbacks >= 1, "the capture is meant to exercise the monolithic fallback"
    got, st = run_lines(pkg, lib, cu8, "-v", max_batch_mib=1, pushes=[1 << 18] * 6)
    assert got == want and st.rl_fallbacks >= 1
    # forcing the monolithic lanes everywhere gives the same lines
    got, st = run_lines(pkg, lib, cu8, "-v", reserved=(C.c_uint32 * 2)(1, 0))
    assert got == want and st.rl_fallbacks == 0


def check_overflow_degrades(pkg, lib):
    """More access-code matches in a batch than the device tables hold (a jammer repeating the sync
    word, ...) costs lines, not the stream: no error, wmb_stats.overflow_batches counts the batches, what is printed is
    a subsequence of the reference's lines, and the same context decodes everything again once the flood is over.
    The tables are shrunk through the test knob opts.reserved[1] >> 8 so that an ordinary capture overflows them."""
    cu8 = np.ascontiguousarray(np.tile(load_fixture("synth_mixed_1m6.cu8"), 3))
    want = oracle_lines(cu8, "-v")
    for cap in (4, 6):
        with pkg.WmbusB200("-v", lib=lib, max_batch_mib=4, reserved=(C.c_uint32 * 2)(0, cap << 8)) as ctx:
            got = [orc.blank_ts(l) for l in ctx.process(cu8.ctypes.data, len(cu8), flush=True)]
            st = ctx.stats()
            assert st.overflow_batches >= 1, "the capture is meant to overflow a %d-entry table" % cap
            assert len(got) < len(want)
            it = iter(want)
            assert all(any(l == w for w in it) for l in got), "lines after an overflow must still be reference lines, in order"
    # a table that is large enough again: nothing is lost, nothing is counted
    with pkg.WmbusB200("-v", lib=lib, max_batch_mib=4, reserved=(C.c_uint32 * 2)(0, 4096 << 8)) as ctx:
        got = [orc.blank_ts(l) for l in ctx.process(cu8.ctypes.data, len(cu8), flush=True)]
        assert got == want and ctx.stats().overflow_batches == 0


def carriers_capture(n_bytes=3 << 20, fs=1.6e6):
    """Five emitters on five carriers of one 1.6 MS/s capture (offsets from the centre, kHz): T1 at +325 and -150,
    C1 at +575, S1 at -325 and +100."""
    import importlib
    synth = importlib.import_module("rtl-wmbus_b200.synth")
    E = synth.Emitter
    em = [E("T1", 0x71200023, amp=60.0, offset_hz=325e3 + 6e3, l_field=0x29, period_s=0.11, start_s=0.004, seed=31),
          E("T1", 0x64700082, amp=55.0, offset_hz=-150e3 - 4e3, l_field=0x19, period_s=0.13, start_s=0.021, seed=32),
          E("C1A", 0x20338739, amp=55.0, offset_hz=575e3 + 3e3, l_field=0x19, period_s=0.12, start_s=0.040, seed=33),
          E("S1", 0x19131290, amp=60.0, offset_hz=-325e3 + 2e3, l_field=0x19, period_s=0.17, start_s=0.010, seed=34),
          E("S1", 0x02717473, amp=55.0, offset_hz=100e3 - 3e3, l_field=0x2E, period_s=0.19, start_s=0.060, seed=35)]
    cap, plan = synth.synth_capture(n_bytes, fs=fs, emitters=em, seed=0xB2000061, noise_sigma=4.0)
    return np.ascontiguousarray(cap.numpy()), [(325, "T"), (-150, "T"), (575, "T"), (-325, "S"), (100, "S")]


def check_carriers(pkg, lib, to_device=None):
    """the -s mixer with the carrier as a parameter.  (1) {+13, -13} through the parameter is the
    reference's -s, bit for bit (the goldens of -s come from the reference binary); (2) every carrier of a five-carrier
    capture decodes to the lines of the oracle's restatement with the same parameters, and each carrier yields its
    emitter's telegrams."""
    import importlib
    shard = importlib.import_module("rtl-wmbus_b200.shard")
    cu8s = load_fixture("synth_mixed_2m4_shift.cu8")
    want = oracle_lines(cu8s, "-v -d 3 -s")
    got, _ = run_lines(pkg, lib, cu8s, "-v -d 3", simultaneous=2, carrier_25khz=(C.c_int32 * 2)(13, -13))
    assert got == want and len(want) > 3

    cu8, carriers = carriers_capture()
    dev = to_device(cu8) if to_device else None
    if dev is not None:
        run = lambda ctx: ctx.process_device(dev.data_ptr(), len(cu8), flush=True)
    else:
        run = lambda ctx: ctx.process(cu8.ctypes.data, len(cu8), flush=True)
    got = shard.decode_carriers(lambda f, **kw: pkg.WmbusB200(f, lib=lib, **kw), run, carriers, "-v")
    idents = {(325, "T"): "71200023", (-150, "T"): "64700082", (575, "T"): "20338739", (-325, "S"): "19131290", (100, "S"): "02717473"}
    for (t_off, s_off) in shard.plan_carriers(carriers):
        o = orc.opts_from_flags("-v")
        o.simultaneous = 2
        o.carrier_25khz[0] = 0 if t_off is None else t_off // 25
        o.carrier_25khz[1] = 0 if s_off is None else s_off // 25
        o.t1c1_enabled = int(t_off is not None); o.s1_enabled = int(s_off is not None)
        ref = [orc.blank_ts(l) for l in orc.run_lines(cu8, o)]
        mine = []
        if t_off is not None: mine += got[(t_off, "T")]
        if s_off is not None: mine += got[(s_off, "S")]
        assert sorted(orc.blank_ts(l) for l in mine) == sorted(ref)
    for key, ident in idents.items():
        ok = [l for l in got[key] if l.split(";")[2] == "1" and ident in l]
        assert len(ok) >= 4, (key, len(got[key]), len(ok))


def check_cw_interferer(pkg, lib):
    """A clean carrier next to the channel (here: a CW tone 310 kHz off, through the box filter's side lobe) parks the
    clock filter near a fixed point between telegrams, where speculative lanes do not re-join the true trajectory: a
    third of the lanes is refuted and re-run (segment pass + fix-up block).  The lines must not notice."""
    import importlib
    synth = importlib.import_module("rtl-wmbus_b200.synth")
    E = synth.Emitter
    em = [E("T1", 0x71200023, amp=60.0, offset_hz=6e3, l_field=0x29, period_s=0.11, start_s=0.004, seed=31),
          E("S1", 0x19131290, amp=50.0, offset_hz=2e3, l_field=0x19, period_s=0.17, start_s=0.050, seed=34)]
    cap, _ = synth.synth_capture(2 << 20, emitters=em, seed=0xB2000071, noise_sigma=2.0)
    x = cap.numpy().astype(np.float64).reshape(-1, 2)
    tone = 30.0 * np.exp(2j * np.pi * 310e3 / 1.6e6 * np.arange(len(x)))
    x[:, 0] += tone.real; x[:, 1] += tone.imag
    cu8 = np.ascontiguousarray(np.clip(np.round(x), 0, 255).astype(np.uint8).reshape(-1))
    want = oracle_lines(cu8, "-v")
    assert len(want) >= 12
    for tuning in (dict(chunk_samples=4096), dict(chunk_samples=1024, max_batch_mib=1), dict()):
        got, st = run_lines(pkg, lib, cu8, "-v", **tuning)
        assert got == want
        if tuning:
            assert st.lanes_rerun > 50, "the capture is meant to refute lanes (%d of %d)" % (st.lanes_rerun, st.lanes_run)


def check_prefilter(pkg, lib):
    """the reference's dormant pre-decimation low-passes as the front end -- opts.prefilter = 1 the 23-tap
    float FIR, 2 the float polyphase filter of ppf.h, 3 / 4 their 24.8 fixed-point twins (rtl_wmbus.c:197-333).  The
    oracle's versions are pinned against the reference's own functions (tests/test_oracle.py); here the product against
    the oracle: every dphi / rssi sample, and the lines, without and with the mixer, through several batches."""
    import importlib
    synth = importlib.import_module("rtl-wmbus_b200.synth")
    cu8 = load_fixture("synth_mixed_1m6.cu8")
    cap, _ = synth.synth_capture(1 << 21, fs=1.6e6, emitters=synth.default_emitters("mixed"), seed=0xB2000081, center_shift_hz=325e3)
    shifted = np.ascontiguousarray(cap.numpy())
    for mode in (1, 2, 3, 4):
        for flags in ("", "-a", "-o") if mode == 1 else ("", "-a -p S"):
            check_stages(pkg, lib, cu8[:1 << 19], flags, orc_extra=dict(prefilter=mode), prefilter=mode)
        check_stages(pkg, lib, shifted[:1 << 19], "-s", orc_extra=dict(prefilter=mode), prefilter=mode)
        for data, flags in ((cu8, "-v"), (shifted, "-v -s")):
            o = orc.opts_from_flags(flags)
            o.prefilter = mode
            want = [orc.blank_ts(l) for l in orc.run_lines(data, o)]
            assert len(want) >= 8
            if mode == 1 or flags == "-v":
                got, _ = run_lines(pkg, lib, data, flags, prefilter=mode)
                assert got == want
            got, _ = run_lines(pkg, lib, data, flags, prefilter=mode, max_batch_mib=1, pushes=[4096 * 2 * 7, 1 << 19, 12345])
            assert got == want
    # only defined at 1.6 MS/s, and there are four of them
    ctx = C.c_void_p()
    o = pkg.opts_from_flags(lib, "-d 3", prefilter=1)
    assert lib.wmb_create(C.byref(o), 0, C.byref(ctx)) == -1 and b"decimation" in lib.wmb_last_error()
    o = pkg.opts_from_flags(lib, "", prefilter=5)
    assert lib.wmb_create(C.byref(o), 0, C.byref(ctx)) == -1 and b"prefilter" in lib.wmb_last_error()


def check_lane_event_overflow(pkg, lib):
    """Found by tests/tools/fuzz_hostsim.py (seed 22, case 525): 0.8 MS/s, an in-channel CW tone under three emitters.  The
    T1/C1 run-length tracker's bit length collapses to a fraction of a sample and single edges emit tens of thousands of
    bits (193 k in one 1024-sample stretch): more than a lane's event buffer holds.  That used to end the stream
    (WMB_E_OVERFLOW); now the lane keeps what fits, the batch is counted in wmb_stats.overflow_batches and the lines are
    still the reference's -- there is no telegram in such a stretch."""
    import fuzz_cases
    c = fuzz_cases.case(22, 525)
    assert c["flags"] == "-v -d 1" and c["n"] == 827392
    cu8 = fuzz_cases.build_capture(c)
    want = oracle_lines(cu8, c["flags"])
    assert len(want) == 7
    seen = 0
    for tuning, pushes in ((dict(), None), (dict(chunk_samples=8192), None), (c["tuning"], c["pushes"]), (dict(max_batch_mib=1, chunk_samples=4096), None)):
        got, st = run_lines(pkg, lib, cu8, c["flags"], pushes=pushes, **tuning)
        assert got == want, tuning
        seen += st.overflow_batches
    assert seen >= 2, "the capture is meant to overflow a lane's event buffer"
    # seed 33, case 1190 (2.4 MS/s, -d 3, a CW tone 506 kHz off): the same regime for so long that a 2 MiB batch wrote
    # more events than its ring held.  The rings of the run-length streams are now sized for what the lanes of a batch can
    # emit (batches up to 128 MiB), so nothing is lost; with the ring sized by the large-batch rule (test knob) the batch
    # loses that stream's candidates -- and only that stream's: every missing line is a run-length line -- and the
    # stream goes on.
    c = fuzz_cases.case(33, 1190)
    assert c["flags"] == "-d 3" and c["tuning"] == dict(max_batch_mib=2)
    cu8 = fuzz_cases.build_capture(c)
    want = oracle_lines(cu8, c["flags"])
    got, st = run_lines(pkg, lib, cu8, c["flags"], pushes=c["pushes"], **c["tuning"])
    assert got == want and st.overflow_batches == 0
    got, st = run_lines(pkg, lib, cu8, c["flags"], pushes=c["pushes"], reserved=(C.c_uint32 * 2)(0, 2), **c["tuning"])
    it = iter(want)
    assert st.overflow_batches >= 1 and len(want) - 2 <= len(got) <= len(want) and all(any(l == w for w in it) for l in got)
    got, st = run_lines(pkg, lib, cu8, c["flags"])
    assert got == want and st.overflow_batches == 0
    # time-chunk fuzzer, seed 777, case 322 (3.2 MS/s, -d 4, noise sigma 1, a CW tone 614 kHz off): 1 / 2 / 4 MiB batches
    # used to overrun the T1/C1 run-length ring, and the overrun used to cost the batch's time2 candidates as well (15 of
    # 259 lines).  With the knob: only run-length lines are missing; without: none.
    c = fuzz_cases.time_chunk_case(777, 322)
    assert c["flags"] == "-v -d 4" and c["n"] == 23887872
    cu8 = fuzz_cases.build_capture(c)
    want = oracle_lines(cu8, c["flags"])
    assert len(want) == 259
    got, st = run_lines(pkg, lib, cu8, c["flags"], max_batch_mib=4)           # (4 MiB batches lost the most: 17 lines)
    assert got == want and st.overflow_batches == 0
    got, st = run_lines(pkg, lib, cu8, c["flags"], max_batch_mib=1, reserved=(C.c_uint32 * 2)(0, 2))
    assert st.overflow_batches >= 1
    missing = list(want)
    for l in got:
        missing.remove(l)                                        # (raises if a line was invented)
    assert all(l.startswith("rla;") for l in missing), missing[:3]


def check_sample_index_wrap(pkg, lib):
    """The device keeps 40 bits of the decimated sample index in its bit events (15.9 days of streaming at 800 kS/s).
    A stream positioned just below 2^40 must decode the telegrams that span the wrap exactly like a fresh stream:
    same lines, same order, in small pushes (candidates stay pending across the wrap) and in one."""
    cu8 = load_fixture("synth_mixed_1m6.cu8")
    want = oracle_lines(cu8, "-v")
    data = np.ascontiguousarray(cu8, np.uint8)
    m_total = len(data) // 4                                  # decimated samples of the capture (d = 2)
    for back in (2048 * 20, (m_total // 2) // 2048 * 2048):   # the wrap falls early in / in the middle of the capture
        first_iq = ((1 << 40) - back) * 2
        for step in (len(data), 1 << 17):
            with pkg.WmbusB200("-v", lib=lib, max_batch_mib=1) as ctx:
                ctx.seek(first_iq)
                lines = []
                for off in range(0, len(data), step):
                    n = min(step, len(data) - off)
                    lines += ctx.process(data.ctypes.data + off, n, flush=False, timestamp_mode=2)
                lines += ctx.process(0, 0, flush=True, timestamp_mode=2)
            shard = importlib.import_module("rtl-wmbus_b200.shard")
            keys = [shard.line_key(l) for l in lines]
            assert keys == sorted(keys) and keys[0][0] > (1 << 40) - back and keys[-1][0] > (1 << 40)
            assert [shard.blank_position(l) for l in lines] == want, (back, step)


def check_bitsync_stages(pkg, lib, cu8, flags):
    """Every intermediate of the bit-sync stage, one by one, against the oracle (which is pinned stage by stage to
    the reference's own functions, tests/test_oracle.py): slicer bits behind the optional DC block (a6/a7), clock signs
    of the band-pass (a9), lock-stencil strobes (a10), and per (chain, algorithm) the exact sequence of decoder calls
    -- sample, bit, access-code flag, run-length reset, (unsigned)rssi (a11, a12, a13)."""
    o = orc.opts_from_flags(flags)
    gran = 4096 * max(1, o.decimation)
    data = np.ascontiguousarray(cu8[:len(cu8) // gran * gran], np.uint8)
    counts = {}
    with pkg.WmbusB200(flags, lib=lib, reserved=(C.c_uint32 * 2)(0, 1)) as ctx:
        ctx.process(data.ctypes.data, len(data), flush=True)
        assert ctx.stats().batches == 1
        for chain in (0, 1):
            if (chain == 0 and not o.t1c1_enabled) or (chain == 1 and not o.s1_enabled):
                continue
            st = orc.stages(data, o, chain)
            M = st["M"]
            for which, name in ((0, "bit"), (2, "clk"), (1, "strobe")):
                if which != 0 and not o.t2_enabled:
                    continue
                got = ctx.debug_bits(chain, which, M)
                bad = np.nonzero(got != st[name])[0]
                assert len(bad) == 0, (flags, chain, name, len(bad), bad[:5])
            for algo in (0, 1):
                if (algo == 0 and not o.rla_enabled) or (algo == 1 and not o.t2_enabled):
                    continue
                want = orc.events(st, chain, algo)
                got = ctx.debug_events(chain, algo)
                assert len(got["m"]) == len(want), (flags, chain, algo, len(got["m"]), len(want))
                for f in ("m", "bit", "sync", "rssi") + (("reset",) if algo == 0 else ()):
                    bad = np.nonzero(got[f] != want[f].astype(np.uint64))[0]
                    assert len(bad) == 0, (flags, chain, algo, f, len(bad), bad[:5], got[f][bad[:5]], want[f][bad[:5]])
                counts[(chain, algo)] = (len(want), int(want["sync"].sum()), int(want["reset"].sum()))
    return counts


def check_device_push_ragged(pkg, lib, to_device=None):
    """wmb_push_device with a capture whose length is a multiple of 4096 but not of 4096 * d: the trailing partial
    granule waits for the flush, exactly like the reference's last whole 4096-byte items (rtl_wmbus.c:1301-1308)."""
    cu8 = load_fixture("synth_mixed_2m4_shift.cu8")
    n = len(cu8) // 4096 * 4096
    if n % (4096 * 3) == 0:
        n -= 4096
    data = np.ascontiguousarray(cu8[:n])
    want = oracle_lines(data, "-v -d 3 -s")
    keep = to_device(data) if to_device else data           # GPU: a device copy; CPU simulation: the host array
    ptr = keep.data_ptr() if to_device else data.ctypes.data
    with pkg.WmbusB200("-v -d 3 -s", lib=lib) as ctx:
        ctx.push_device(ptr, n)
        ctx.poll_flush()
        got = ctx.take_lines()
    assert got == want and len(want) > 3
