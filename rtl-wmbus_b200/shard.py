"""Multi-GPU sharding rule of the path: one independent capture per rank, no data-path
collective; the only exchange is the all-reduce of the packet counters.  Used by bench.py (NCCL) and by the
world_size-2 gloo test on CPU."""
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
