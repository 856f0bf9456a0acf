"""Time-chunk sharding under random conditions, on the CPU build: random captures and flags
(tests/fuzz_cases.py), 2-4 chunks, halos of 2^16..2^18 decimated samples; the merged lines must be the oracle's, in
order; a chunk whose boundary digest differs from its left neighbour's repeats with a 4x longer halo.
    python tests/tools/fuzz_time_chunks.py [seconds] [seed]"""
import importlib, sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import fuzz_cases, pipeline_checks as pc
from conftest import HOSTSIM_SO
pkg = importlib.import_module("rtl-wmbus_b200"); shard = importlib.import_module("rtl-wmbus_b200.shard")
lib = pkg.load_library(HOSTSIM_SO)
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 600.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 4242)
t0 = time.time(); k = 0; total_retries = 0
rng_b = np.random.default_rng((int(sys.argv[2]) if len(sys.argv) > 2 else 4242) + 7000003)     # batch size per case (its own generator)
while time.time() - t0 < budget:
    k += 1
    c = fuzz_cases.draw_time_chunk_case(rng)
    if c is None:
        continue
    cu8 = fuzz_cases.build_capture(c)
    flags = c["flags"]
    want = pc.oracle_lines(cu8, flags)
    world, halo = c["world"], c["halo"]
    overflow = 0
    mib = int(rng_b.choice([1, 1, 2, 8, 64]))
    geom = {k_: v for k_, v in c["tuning"].items() if k_ in ("chunk_samples", "warmup_samples")}      # lane geometry of the case
    got, ends, retries, ok = [], [], 0, True
    for rank in range(world):
        h = halo
        while True:
            with pkg.WmbusB200(flags, lib=lib, max_batch_mib=mib, **geom) as ctx:
                lines, ds, de, start = shard.decode_time_chunk(ctx, lambda lo, hi: ctx.push(cu8.ctypes.data + lo, hi - lo),
                                                               len(cu8), c["d"], rank, world, h)
                overflow += ctx.stats().overflow_batches
            if rank == 0 or start == 0 or ds == ends[rank - 1]:
                break
            h *= 4; retries += 1
            if h > (1 << 24):
                ok = False
                break
        ends.append(de); got.append(lines)
    total_retries += retries
    merged = shard.merge_lines(got)
    same = ok and merged == want
    lost = 0
    if ok and not same and overflow:                          # a device table was full: lines may be missing, none may be invented
        it = iter(want)
        same = all(any(l == w for w in it) for l in merged)
        lost = len(want) - len(merged)
    print("case %d %s flags=%r world=%d halo=%d mib=%d geom=%r lines=%d retries=%d overflow_batches=%d lost=%d" % (k, "ok" if same else "MISMATCH", flags, world, halo, mib, geom, len(want), retries, overflow, lost), flush=True)
    if not same:
        sys.exit(1)
print("done", k, "cases,", total_retries, "halo retries")
