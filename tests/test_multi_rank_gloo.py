"""`not gpu`: the N>1 path of bench.py on CPU -- two ranks (gloo), one independent capture each, packet
counters all-reduced.  Each rank decodes through the C ABI of the CPU-simulation build (host logic only;
test infrastructure) and the reduced totals must equal the oracle's totals for both captures."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import HOSTSIM_SO, ROOT


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module("rtl-wmbus_b200")
    synth = importlib.import_module("rtl-wmbus_b200.synth")
    shard = importlib.import_module("rtl-wmbus_b200.shard")
    import orc
    lib = pkg.load_library(HOSTSIM_SO)
    buf, _ = synth.synth_capture(1 << 20, emitters=synth.default_emitters("mixed"), seed=shard.capture_seed(5, rank))
    cu8 = buf.numpy()
    with pkg.WmbusB200("-v", lib=lib) as ctx:
        lines = ctx.process(cu8.ctypes.data, len(cu8), flush=True)
    mine = shard.count_lines(lines)
    want = shard.count_lines(orc.run_lines(cu8, orc.opts_from_flags("-v")))
    total = shard.reduce_counts(mine)
    total_want = shard.reduce_counts(want)
    if rank == 0:
        out.put((total, total_want, int(mine[0])))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_all_reduce_packet_counters(hostsim_lib):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    total, total_want, rank0_lines = q.get(timeout=300)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert total == total_want
    assert total["lines"] > rank0_lines > 0, "the total must include the other rank's capture"
