"""Time-chunk sharding of one capture: every chunk is decoded by its own context
from a warm-up halo, the boundary states must chain, and the union of the chunks' lines must be exactly the
sequential run's lines.  `not gpu`: through the C ABI of the CPU-simulation build (host logic + kernels' phase
functions; test infrastructure), single process and two ranks over gloo."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import pipeline_checks as pc
from conftest import HOSTSIM_SO, ROOT


def _capture(nbytes, seed=0xB2000047, emitters="mixed"):
    synth = importlib.import_module("rtl-wmbus_b200.synth")
    buf, _ = synth.synth_capture(nbytes, emitters=synth.default_emitters(emitters), seed=seed)
    return np.ascontiguousarray(buf.numpy())


def _pusher(ctx, cu8):
    return lambda lo, hi: ctx.push(cu8.ctypes.data + lo, hi - lo)


def check_time_chunks(pkg, lib, cu8, flags, world, halo_m, d=2):
    shard = importlib.import_module("rtl-wmbus_b200.shard")
    want = pc.oracle_lines(cu8, flags)                        # the CPU oracle, not the library itself
    got, ends, retries = [], [], 0
    for rank in range(world):
        h = halo_m
        while True:
            with pkg.WmbusB200(flags, lib=lib, max_batch_mib=1) as ctx:
                lines, ds, de, start = shard.decode_time_chunk(ctx, _pusher(ctx, cu8), len(cu8), d, rank, world, h)
            if rank == 0 or start == 0 or ds == ends[rank - 1]:
                break
            h *= 4
            retries += 1
        ends.append(de)
        got.append(lines)
    flat = shard.merge_lines(got)                             # the sequential run's print order
    assert flat == want
    assert len(want) > 10
    assert all(len(part) > 0 for part in got)
    return retries


def test_time_chunks_match_the_sequential_run(hostsim_lib, pkg):
    cu8 = _capture(6 << 20)
    retries = check_time_chunks(pkg, hostsim_lib, cu8, "-v", world=3, halo_m=1 << 18)
    assert retries == 0


def test_time_chunks_with_mixer_decimation_and_dc_block(hostsim_lib, pkg):
    """-d 3 -s -o: the mixer LUT phase and the decimation phase must stay global after wmb_seek."""
    cu8 = np.fromfile(os.path.join(ROOT, "tests", "golden", "synth_mixed_2m4_shift.cu8"), np.uint8)
    check_time_chunks(pkg, hostsim_lib, cu8, "-v -d 3 -s -o", world=2, halo_m=1 << 16, d=3)


def test_short_halo_is_detected_and_repeated(hostsim_lib, pkg):
    cu8 = _capture(4 << 20, seed=0xB2000048)
    retries = check_time_chunks(pkg, hostsim_lib, cu8, "-v", world=2, halo_m=1 << 10)
    assert retries >= 1, "a 1024-sample halo cannot re-join the clock-recovery filters"


def test_telegram_that_ends_after_a_gap_in_the_input(hostsim_lib, pkg):
    """Found by tests/tools/fuzz_time_chunks.py (seed 882, case 126): a T1 telegram starts in chunk 0, the input goes dead
    (constant samples: no edges, so the run-length tracker delivers no bits) for 456 k decimated samples, and the first
    edge after the gap delivers the rest of the telegram at once -- 291 k samples after the chunk's end, beyond a right
    halo of one maximum telegram.  The worker goes on pushing while wmb_pending_before() reports a telegram matched in
    its chunk in flight; the line (CRC and 3-out-of-6 failed, printed all the same by the reference) comes from chunk 0."""
    import fuzz_cases
    c = fuzz_cases.time_chunk_case(882, 126)
    assert c["flags"] == "-v" and c["world"] == 4 and c["dead"] == 578583
    cu8 = fuzz_cases.build_capture(c)
    check_time_chunks(pkg, hostsim_lib, cu8, c["flags"], world=4, halo_m=1 << 18)
    with pkg.WmbusB200("-v", lib=hostsim_lib, max_batch_mib=1) as ctx:      # the new entry point by itself
        assert ctx.pending_before(1 << 62) == 0
        ctx.push(cu8.ctypes.data, 2 * 2 * 454656)                            # up to chunk 0's end: the telegram is in flight
        assert ctx.pending_before(454656) >= 1 and ctx.pending_before(1000) == 0


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module("rtl-wmbus_b200")
    shard = importlib.import_module("rtl-wmbus_b200.shard")
    lib = pkg.load_library(HOSTSIM_SO)
    cu8 = _capture(4 << 20, seed=0xB2000049)
    with pkg.WmbusB200("-v", lib=lib, max_batch_mib=1) as ctx:
        lines, rounds = shard.decode_time_sharded(ctx, _pusher(ctx, cu8), len(cu8), 2, halo_m=1 << 12)
    counts = shard.reduce_counts(shard.count_lines(lines))
    gathered = [None] * world
    dist.all_gather_object(gathered, lines)
    if rank == 0:
        import pipeline_checks as pc
        out.put((shard.merge_lines(gathered), [len(g) for g in gathered], pc.oracle_lines(cu8, "-v"), counts, rounds))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_time_sharded_gloo(hostsim_lib):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29950 + os.getpid() % 40
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    flat, per_rank, want, counts, rounds = q.get(timeout=600)
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    assert flat == want and len(want) > 10                    # merged in print order == the oracle's sequential run
    assert counts["lines"] == len(want)
    assert rounds >= 2, "the 4096-sample halo must have been rejected once"
    assert all(n > 0 for n in per_rank)


@pytest.mark.gpu
def test_time_chunks_on_the_gpu(gpu_lib, pkg):
    """Same property through the CUDA library: 64 MiB, three chunks, host pushes (H2D path) and device pushes."""
    synth = importlib.import_module("rtl-wmbus_b200.synth")
    shard = importlib.import_module("rtl-wmbus_b200.shard")
    n = 64 << 20
    cap, _ = synth.synth_capture(n, emitters=synth.default_emitters("mixed"), seed=0xB200004A, device="cuda")
    cu8 = np.ascontiguousarray(cap.cpu().numpy())
    assert check_time_chunks(pkg, gpu_lib, cu8, "-v", world=3, halo_m=1 << 18) == 0
    want = pc.oracle_lines(cu8, "")
    got = []
    ends = []
    for rank in range(4):
        with pkg.WmbusB200("", lib=gpu_lib, max_batch_mib=8) as ctx:
            push = lambda lo, hi, c=ctx: c.push_device(cap.data_ptr() + lo, hi - lo)
            lines, ds, de, start = shard.decode_time_chunk(ctx, push, n, 2, rank, 4)
        assert rank == 0 or ds == ends[-1]
        ends.append(de)
        got.append(lines)
    assert shard.merge_lines(got) == want and len(want) > 100
