"""Full-size parity helper -- TEST INFRASTRUCTURE ONLY.

Runs the UNMODIFIED reference binary (oracle/_ref/rtl_wmbus) over a large capture on all host cores by cutting the
capture into time chunks (chunk starts on multiples of 4096*d input bytes, which keeps the decimation
phase `(k+1) % d` (rtl_wmbus.c:1350-1352) and the -s mixer phase `(13 k) mod 32d` (:1001-1003) global).

The reference prints no sample positions, so ownership is taken from what the program itself guarantees: it is a
causal, deterministic loop, hence its output for the bytes [h, b) is a PREFIX of its output for [h, e), b < e.
Chunk g = [b_g, b_g+1) is therefore decoded twice from the same halo start h_g = b_g - halo:

    A = reference([h_g, b_g+1))      B = reference([h_g, b_g))        owned_g = A[len(B):]

i.e. exactly the lines the sequential run prints while it consumes the chunk (a telegram belongs to the chunk in
which its last bit arrives), given that the halo is long enough for the cold-started filters, clock recovery,
run-length tracker and busy decoders to have re-joined the sequential trajectory (<= 20 k decimated
samples for the float state, one maximum telegram (~170 k samples for S1) for the decoders; the default halo is 2^20
decimated samples).  Concatenating owned_0, owned_1, ... gives the sequential run's lines IN ORDER.
tests/test_ref_chunked.py checks the helper against whole-capture runs of the same binary.
"""
import concurrent.futures
import os
import shutil
import subprocess
import tempfile

import orc


def host_cpus():
    """CPUs this process may really use: min(affinity, cgroup quota, cpu_count)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    quota = None
    try:                                                        # cgroup v2
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            quota = float(q) / float(p)
    except (OSError, ValueError):
        try:                                                    # cgroup v1
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            p = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0 and p > 0:
                quota = q / p
        except (OSError, ValueError):
            pass
    if quota is not None:
        n = min(n, max(1, int(quota + 0.5)))
    return max(1, n)


def scratch_dir(need_bytes):
    """/dev/shm when it has room (page-cache speed for P concurrent readers), else the temp dir."""
    for d in ("/dev/shm", tempfile.gettempdir()):
        try:
            if os.path.isdir(d) and shutil.disk_usage(d).free > need_bytes + (64 << 20):
                return d
        except OSError:
            continue
    return tempfile.gettempdir()


def _ref_on_range(exe, flags, path, lo, hi):
    """Lines (timestamps blanked) the reference prints for the byte range [lo, hi) of the file."""
    if hi <= lo:
        return []
    cmd = f"tail -c +{lo + 1} '{path}' | head -c {hi - lo} | '{exe}' {flags}"
    out = subprocess.run(["bash", "-c", cmd], capture_output=True, check=True).stdout.decode()
    return [orc.blank_ts(l) for l in out.split("\n") if l]


def chunk_plan(nbytes, d, chunks, halo_m):
    """[(halo_start, chunk_start, chunk_end)] in bytes; chunk boundaries are multiples of 4096*d."""
    d = max(1, d)
    gran = 4096 * d
    usable = nbytes // 4096 * 4096                  # the reference drops a trailing partial item (rtl_wmbus.c:1301-1308)
    halo = (2 * d * halo_m + gran - 1) // gran * gran
    chunks = max(1, min(chunks, usable // gran or 1))
    bounds = [min(usable, (usable * g // chunks) // gran * gran) for g in range(chunks)] + [usable]
    return [(max(0, bounds[g] - halo), bounds[g], bounds[g + 1]) for g in range(chunks) if bounds[g + 1] > bounds[g]]


def ref_lines_chunked(path, nbytes, flags, d=2, procs=None, chunks=None, halo_m=1 << 20):
    """The sequential reference run's lines for the first nbytes of `path`, computed on `procs` cores."""
    exe = orc.ref_binary()
    assert exe, "oracle/_ref/rtl_wmbus missing (built by oracle/Makefile where /root/reference exists)"
    procs = procs or host_cpus()
    plan = chunk_plan(nbytes, d, chunks or procs, halo_m)

    def work(item):
        h, b, e = item
        a_lines = _ref_on_range(exe, flags, path, h, e)
        if b == h:
            return a_lines
        b_lines = _ref_on_range(exe, flags, path, h, b)
        assert a_lines[:len(b_lines)] == b_lines, "the reference is causal: the halo run must be a prefix"
        return a_lines[len(b_lines):]

    with concurrent.futures.ThreadPoolExecutor(max_workers=procs) as ex:
        parts = list(ex.map(work, plan))
    return [l for part in parts for l in part]


def ref_lines_whole(path, nbytes, flags):
    return _ref_on_range(orc.ref_binary(), flags, path, 0, nbytes)
