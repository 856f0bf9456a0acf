"""Multi-GPU sharding rules of the path.

* Independent captures (BASELINE config 5): one capture per rank, no data-path collective; the only exchange is
  the all-reduce of the packet counters.  Used by bench.py (NCCL) and by the world_size-2 gloo test on CPU.
* Time chunks of ONE capture (config 4): rank g owns the decimated samples [S_g, S_g+1), warm-starts a
  halo earlier and proves that it re-joined the sequential run by comparing `wmb_boundary_state()` with its left
  neighbour's (an all-gather of two digests); a rank whose halo was too short repeats with a longer one."""
from __future__ import annotations

import torch
import torch.distributed as dist

COUNTER_FIELDS = ("lines", "crc_ok", "t1", "c1", "s1")


def capture_seed(config_index: int, rank: int) -> int:
    """Seed of rank `rank`'s capture (0xB200_0000 + config*16 + rank)."""
    return 0xB2000000 + 16 * config_index + rank


def count_lines(lines) -> torch.Tensor:
    """Packet counters of one rank from its datagram lines (with or without the -v prefix)."""
    c = dict.fromkeys(COUNTER_FIELDS, 0)
    for l in lines:
        f = l.split(";")
        if f[0] in ("rla", "t2a"):
            f = f[1:]
        c["lines"] += 1
        c["crc_ok"] += f[1] == "1"
        c[f[0].lower()] += 1
    return torch.tensor([c[k] for k in COUNTER_FIELDS], dtype=torch.int64)


def reduce_counts(counts: torch.Tensor, device=None) -> dict:
    """Sum the counters over all ranks (NCCL over NVLink on GPUs, gloo in the CPU test)."""
    t = counts.to(device) if device is not None else counts.clone()
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return dict(zip(COUNTER_FIELDS, (int(v) for v in t.cpu())))


# ---- time chunks of one capture ---------------------------------------------------------------------

# right halo, decimated samples.  The run-length tracker of the S1 chain accepts up to just under 36 samples per chip
# (rtl_wmbus.c:659; nominal 24.4), so the longest telegram it can deliver -- (16 * 290 + 17) chips -- spans up to
# 4657 * 36 = 167,652 samples; T1/C1: (12 * 290 + 13) chips * 8 * 1.25 < 35,000.
MAX_TELEGRAM_M = 1 << 18


def line_key(line: str):
    """(end sample, stream priority) of a line taken with timestamp_mode=2 ("@<sample>.<prio>" in the TIMESTAMP
    column): the position at which the reference prints it (rtl_wmbus.c:1354-1355)."""
    f = line.split(";")
    ts = f[4 if f[0] in ("rla", "t2a") else 3]
    a, b = ts[1:].split(".")
    return int(a), int(b)


def blank_position(line: str) -> str:
    f = line.split(";")
    f[4 if f[0] in ("rla", "t2a") else 3] = "TS"
    return ";".join(f)


def merge_lines(parts):
    """Lines of several time chunks (each taken with timestamp_mode=2) -> the sequential run's print order, with the
    TIMESTAMP column blanked.  Within one chunk the order is already right; across chunks a telegram that started in
    chunk g may finish after one that started in chunk g+1, so the merge is by print position (stable)."""
    flat = [l for part in parts for l in part]
    flat.sort(key=line_key)
    return [blank_position(l) for l in flat]


def chunk_bounds(n_bytes: int, d: int, world: int):
    """IQ-sample boundaries [k_0 .. k_world] of `world` time chunks, multiples of the batch granule."""
    gran = 2048 * d                                   # IQ samples per 4096*d input bytes
    n_iq = (n_bytes // 2) // gran * gran
    return [min(n_iq, (n_iq * g // world) // gran * gran) for g in range(world)] + [n_iq]


def decode_time_chunk(ctx, push, n_bytes: int, d: int, rank: int, world: int, halo_m: int = 1 << 18):
    """Decode rank `rank`'s chunk of a capture of n_bytes cu8 bytes.  `push(byte_lo, byte_hi)` feeds that byte
    range of the capture to ctx (host or device memory: the caller's business).
    Returns (lines, digest_start, digest_end, halo_start_iq): digest_start is None for a chunk that starts at 0.
    The lines carry their print position in the TIMESTAMP column (timestamp_mode 2) for merge_lines()."""
    import hashlib
    k = chunk_bounds(n_bytes, d, world)
    lo, hi = k[rank], k[rank + 1]
    gran = 2048 * d
    start = max(0, lo - (halo_m * d + gran - 1) // gran * gran)
    ctx.seek(start)
    ctx.set_line_window(lo // d, hi // d if rank + 1 < world else (1 << 63))
    lines = []
    dig_start = None
    if start < lo:
        push(2 * start, 2 * lo)
        lines += ctx.take_lines(2)
    if lo > 0:
        dig_start = hashlib.sha256(ctx.boundary_state()).digest()
    push(2 * lo, 2 * hi)
    lines += ctx.take_lines(2)
    dig_end = hashlib.sha256(ctx.boundary_state()).digest()
    if rank + 1 < world:                              # finish the telegrams that started in the chunk
        step = (MAX_TELEGRAM_M * d + gran - 1) // gran * gran
        tail = min(k[world], hi + step)
        push(2 * hi, 2 * tail)
        # MAX_TELEGRAM_M bounds a telegram whose samples carry edges.  One that runs into a gap in the input (dead air:
        # the run-length tracker emits nothing until the next edge, then all the missing bits at once) ends arbitrarily
        # late: go on while a telegram matched in the chunk is still in flight
        while tail < k[world] and ctx.pending_before(hi // d) > 0:
            nxt = min(k[world], tail + step)
            push(2 * tail, 2 * nxt)
            tail = nxt
        ctx.poll_flush()
    else:
        if n_bytes > 2 * hi:
            push(2 * hi, n_bytes)                     # the ragged end of the capture (the reference drops a short item)
        ctx.poll_flush()
    lines += ctx.take_lines(2)
    return lines, dig_start, dig_end, start


def decode_time_sharded(ctx, push, n_bytes: int, d: int, halo_m: int = 1 << 18):
    """All ranks: decode one capture in time chunks, exact by construction (see module docstring).
    Returns (my_lines, rounds)."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    rounds = 0
    lines = None
    redo = True
    while True:
        rounds += 1
        if redo:
            lines, ds, de, start = decode_time_chunk(ctx, push, n_bytes, d, rank, world, halo_m)
        mine = torch.zeros(65, dtype=torch.uint8)
        mine[:32] = torch.frombuffer(bytearray(ds or bytes(32)), dtype=torch.uint8)
        mine[32:64] = torch.frombuffer(bytearray(de), dtype=torch.uint8)
        mine[64] = 1 if (ds is None or start == 0) else 0          # started from the true beginning: exact
        if world > 1:
            backend = dist.get_backend()
            dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
            allv = [torch.zeros(65, dtype=torch.uint8, device=dev) for _ in range(world)]
            dist.all_gather(allv, mine.to(dev))
            allv = [v.cpu() for v in allv]
        else:
            allv = [mine]
        bad = [g for g in range(1, world)
               if not allv[g][64] and not torch.equal(allv[g][:32], allv[g - 1][32:64])]
        if not bad:
            return lines, rounds
        redo = rank in bad
        if redo:
            halo_m *= 4                                 # too short: the neighbour's state was not reached yet


# ---- many carriers in one capture --------------------------------------------------

def plan_carriers(carriers):
    """carriers: [(offset_khz, "T" | "S"), ...] -- the carriers of one capture, as offsets from its centre frequency
    (multiples of 25 kHz, the grid of the reference's mixer table, rtl_wmbus.c:974-993) and the chain that listens to
    each: "T" = the T1/C1 chain, "S" = the S1 chain.  A context has one chain of each kind (the reference's -s is one
    context with {+325 "T", -325 "S"}), so the carriers are dealt out two per context.
    Returns [(t_offset_khz | None, s_offset_khz | None), ...]."""
    for off, kind in carriers:
        if kind not in ("T", "S"):
            raise ValueError(f"carrier kind {kind!r}: 'T' (T1/C1 chain) or 'S' (S1 chain)")
        if off % 25:
            raise ValueError(f"carrier offset {off} kHz is not on the 25 kHz grid")
    ts = [off for off, kind in carriers if kind == "T"]
    ss = [off for off, kind in carriers if kind == "S"]
    n = max(len(ts), len(ss))
    return [(ts[i] if i < len(ts) else None, ss[i] if i < len(ss) else None) for i in range(n)]


def decode_carriers(make_ctx, run, carriers, flags: str = ""):
    """Decode every carrier of one capture: one context per (T, S) pair of plan_carriers(), all over the same input.
    make_ctx(flags, **opts) -> a WmbusB200; run(ctx) -> its lines for the whole capture (e.g.
    `lambda ctx: ctx.process_device(ptr, n, flush=True)` -- the capture stays where it is, every context reads it).
    Returns {(offset_khz, kind): [lines in print order]}."""
    import ctypes as C
    out = {}
    for t_off, s_off in plan_carriers(carriers):
        carr = (C.c_int32 * 2)(0 if t_off is None else t_off // 25, 0 if s_off is None else s_off // 25)
        opts = dict(simultaneous=2, carrier_25khz=carr, t1c1_enabled=int(t_off is not None), s1_enabled=int(s_off is not None))
        with make_ctx(flags, **opts) as ctx:
            lines = run(ctx)
        if t_off is not None:
            out[(t_off, "T")] = []
        if s_off is not None:
            out[(s_off, "S")] = []
        for l in lines:
            f = l.split(";")
            mode = f[1] if f[0] in ("rla", "t2a") else f[0]
            out[(s_off, "S") if mode == "S1" else (t_off, "T")].append(l)
    return out
