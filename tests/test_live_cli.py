"""`-m gpu`: the drop-in host program on a LIVE stream: `rtl_sdr ... | rtl_wmbus_b200` is a pipe that
delivers 3.2 MB/s, not a file.  The tests feed the CLI through a pipe at the real sample rate and in bursts, compare
the lines with the offline result, bound how long a telegram waits for its line, and fire the `-f` watchdog the way the
reference's would fire (rtl_wmbus.c:71-78, :1238-1246, :1300-1302: a whole 4096-byte item must arrive within 2 s)."""
import importlib
import os
import subprocess
import threading
import time

import numpy as np
import pytest

import orc
import pipeline_checks as pc
from conftest import load_fixture

gpu = pytest.mark.gpu


def _exe(pkg):
    return os.path.join(os.path.dirname(pkg.library_path()), "rtl_wmbus_b200")


def _sim_exe(hostsim_lib):
    """the same host program linked against the CPU simulation of the library (tests/hostsim: test infrastructure)"""
    from conftest import HOSTSIM_SO
    return os.path.join(os.path.dirname(HOSTSIM_SO), "rtl_wmbus_hostsim")


def _run_fed(exe, flags, feeder, timeout=60):
    """Start the CLI, run feeder(stdin) in this thread, collect (line, arrival time) pairs on a reader thread."""
    p = subprocess.Popen([exe] + flags.split(), stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.PIPE, bufsize=0)
    got = []

    def reader():
        for raw in iter(p.stdout.readline, b""):
            got.append((raw.decode().rstrip("\n"), time.perf_counter()))
    th = threading.Thread(target=reader, daemon=True)
    th.start()
    try:
        feeder(p.stdin)
    finally:
        try:
            p.stdin.close()
        except BrokenPipeError:
            pass
    rc = p.wait(timeout=timeout)
    th.join(timeout=5)
    return rc, got, p.stderr.read().decode()


@gpu
def test_real_time_pipe_same_lines_and_bounded_latency(pkg, gpu_lib):
    """1.6 MS/s real time, 16 KiB writes: identical lines, and a telegram's line is out at most ~0.3 s after its last
    bit went into the pipe (100 ms hand-over cadence + one device pass)."""
    exe = _exe(pkg)
    cu8 = np.ascontiguousarray(np.tile(load_fixture("synth_mixed_1m6.cu8"), 3))          # 1.4 s of signal
    want = pc.oracle_lines(cu8, "-v")
    with pkg.WmbusB200("-v", lib=gpu_lib) as ctx:
        stamped = ctx.process(cu8.ctypes.data, len(cu8), flush=True, timestamp_mode=2)
    shard = importlib.import_module("rtl-wmbus_b200.shard")
    end_m = [shard.line_key(l)[0] for l in stamped]                    # decimated sample at which each line is due
    assert [shard.blank_position(l) for l in stamped] == want
    chunk, rate = 16384, 3.2e6                                          # bytes, bytes per second
    t0 = [0.0]

    def feeder(w):
        subprocess.run([exe, "-V"], capture_output=True)               # the image is paged in before the clock starts
        time.sleep(1.0)                                                 # CUDA context of the child
        t0[0] = time.perf_counter()
        for i, off in enumerate(range(0, len(cu8), chunk)):
            due = t0[0] + off / rate
            d = due - time.perf_counter()
            if d > 0:
                time.sleep(d)
            w.write(cu8[off:off + chunk].tobytes())
    rc, got, err = _run_fed(exe, "-v", feeder)
    assert rc == 0, err
    assert [orc.blank_ts(l) for l, _ in got] == want
    lat = sorted(t - (t0[0] + 4 * m / rate) for (_, t), m in zip(got, end_m))      # 4 input bytes per decimated sample
    p99 = lat[min(len(lat) - 1, int(0.99 * len(lat)))]
    print(f"[live] {len(lat)} lines, latency median {1e3 * lat[len(lat) // 2]:.0f} ms, p99 {1e3 * p99:.0f} ms, max {1e3 * lat[-1]:.0f} ms")
    assert lat[0] > -0.01, "a line cannot precede its telegram"
    # 100 ms hand-over + one device pass (2 ms); the bound leaves room for the first pass of a cold process (lazy module
    # loading) and for a busy host
    assert lat[len(lat) // 2] < 0.15 and p99 < 0.5


def check_bursty_pipe(exe):
    cu8 = load_fixture("synth_mixed_2m4_shift.cu8")
    want = pc.oracle_lines(cu8, "-d 3 -s -o")

    def feeder(w):
        cuts = [0, 5000, 5001, 300000, 300000 + 4096 * 3, 900001, len(cu8)]
        for a, b in zip(cuts, cuts[1:]):
            w.write(cu8[a:b].tobytes())
            time.sleep(0.25)
    rc, got, err = _run_fed(exe, "-d 3 -s -o", feeder)
    assert rc == 0, err
    assert [orc.blank_ts(l) for l, _ in got] == want and len(want) > 3


def check_flow_watchdog(exe, mode):
    """-f: the input stops (or trickles below one 4096-byte item per 2 s) without end of file -> the reference's
    message on stderr and EXIT_FAILURE; lines decoded before that were printed."""
    cu8 = load_fixture("synth_mixed_1m6.cu8")
    t = {}

    def feeder(w):
        w.write(cu8[:1 << 20].tobytes())
        t["fed"] = time.perf_counter()
        try:
            if mode == "stall":
                time.sleep(4.0)
            else:
                for _ in range(16):                                     # 100 bytes every 0.25 s: 1.6 KB in 4 s
                    time.sleep(0.25)
                    w.write(bytes(100))
        except BrokenPipeError:
            pass
    t0 = time.perf_counter()
    rc, got, err = _run_fed(exe, "-f -v", feeder)
    assert rc == 1
    assert "rtl_wmbus: monitoring flow" in err and "rtl_wmbus: exiting since incoming data stopped flowing!" in err
    assert len(got) > 0
    assert 1.5 < got[-1][1] - t0 + 10 and time.perf_counter() - t["fed"] < 6.0


def check_watchdog_quiet(exe):
    cu8 = load_fixture("synth_mixed_1m6.cu8")

    def feeder(w):
        for off in range(0, len(cu8), 1 << 16):
            w.write(cu8[off:off + (1 << 16)].tobytes())
            time.sleep(0.02)
    rc, got, err = _run_fed(exe, "-f", feeder)
    assert rc == 0 and "stopped flowing" not in err
    assert [orc.blank_ts(l) for l, _ in got] == pc.oracle_lines(cu8, "")


@gpu
def test_bursty_pipe_same_lines(pkg, gpu_lib):
    check_bursty_pipe(_exe(pkg))


@gpu
@pytest.mark.parametrize("mode", ["stall", "trickle"])
def test_flow_watchdog(pkg, gpu_lib, mode):
    check_flow_watchdog(_exe(pkg), mode)


@gpu
def test_flow_watchdog_quiet_on_a_healthy_stream(pkg, gpu_lib):
    check_watchdog_quiet(_exe(pkg))


# ---- the same host program on the CPU simulation (`not gpu`): the pipe handling and the watchdog are host code ----

def test_bursty_pipe_same_lines_cpu_build(hostsim_lib):
    check_bursty_pipe(_sim_exe(hostsim_lib))


@pytest.mark.parametrize("mode", ["stall", "trickle"])
def test_flow_watchdog_cpu_build(hostsim_lib, mode):
    check_flow_watchdog(_sim_exe(hostsim_lib), mode)


def test_flow_watchdog_quiet_cpu_build(hostsim_lib):
    check_watchdog_quiet(_sim_exe(hostsim_lib))


# ---- the process contract of SURVEY.md 8(b), argv by argv against the reference binary itself ----

ARGVS = ["-h", "-V", "-x", "-d", "-p", "-p X", "-p s", "-p t", "-r 1", "-r 0", "-t 2", "-t 0", "-d abc", "-d 0", "-d 1", "-d 9 -s",
         "-r 0 -t 0", "-p S -p T", "-o -a -v -s", "", "-v", "-o", "-a -v", "-v -r 0 -p S", "-f", "-f -v -p S", "stray", "-v stray -o", "-d 3 -V", "-V -x", "-x -V"]


def _normalise(text, exe):
    """program name (getopt's and the usage text's argv[0]) and the version block are the builds' own"""
    out = []
    for l in text.decode().replace(exe, "EXE").split("\n"):
        if l.startswith("rtl_wmbus: ") and "monitoring flow" not in l and "flow stopped" not in l.lower():
            l = "rtl_wmbus: VERSION"
        out.append(l)
    return out


def check_argv_matrix(exe, real_stdin, skip=()):
    ref = orc.ref_binary()
    cu8 = bytes(load_fixture("synth_mixed_1m6.cu8")) if real_stdin else b""
    for args in ARGVS:
        if args in skip:
            continue
        a = subprocess.run([ref] + args.split(), input=cu8, capture_output=True, timeout=60)
        b = subprocess.run([exe] + args.split(), input=cu8, capture_output=True, timeout=60)
        assert a.returncode == b.returncode, (args, a.returncode, b.returncode, b.stderr)
        assert _normalise(a.stderr, ref) == _normalise(b.stderr, exe), (args, a.stderr, b.stderr)
        la, lb = _normalise(a.stdout, ref), _normalise(b.stdout, exe)
        if a.returncode == 0 and la[0] == "rtl_wmbus: VERSION":     # `-V`: name line + the build's commit / ABI line (rtl_wmbus.c:886-890)
            assert lb[0] == la[0] and len(lb) == len(la)
            continue
        blank = lambda ls: [orc.blank_ts(l) if l.count(";") >= 7 else l for l in ls]
        assert blank(la) == blank(lb), (args, la[:3], lb[:3])


@pytest.mark.skipif(orc.ref_binary() is None, reason="compiled reference (oracle/_ref) not present")
def test_argv_matrix_matches_the_reference_binary_cpu_build(hostsim_lib):
    """getopt string, usage text, exit codes, `-V`, stray operands, repeated options, and the decoded lines of a short
    capture for every accepted flag set: the reference binary and the drop-in host program side by side
    (rtl_wmbus.c:869-967)."""
    check_argv_matrix(_sim_exe(hostsim_lib), real_stdin=True)
