"""Where a time-chunk rank's step goes (single GPU, no collectives): rank 1 of 4 on the 4 GiB -d 3 capture.
   python tools/time_shard_breakdown.py"""
import importlib, sys, time, hashlib
sys.path.insert(0, '.')
import torch
pkg = importlib.import_module("rtl-wmbus_b200"); synth = importlib.import_module("rtl-wmbus_b200.synth")
shard = importlib.import_module("rtl-wmbus_b200.shard")
lib = pkg.load_library()
n = 4 << 30; d = 3; world = 4; rank = 1
cap, plan = synth.synth_capture(n, fs=2.4e6, emitters=synth.default_emitters("mixed"), seed=shard.capture_seed(4, 0), device="cuda")
torch.cuda.synchronize()
ctx = pkg.WmbusB200("-d 3", lib=lib, max_batch_mib=1024)
k = shard.chunk_bounds(n, d, world); lo, hi = k[rank], k[rank + 1]; gran = 2048 * d
halo = (1 << 18); start = max(0, lo - (halo * d + gran - 1) // gran * gran)
tail = min(k[world], hi + (shard.MAX_TELEGRAM_M * d + gran - 1) // gran * gran)
push = lambda a, b: ctx.push_device(cap.data_ptr() + a, b - a)
for it in range(4):
    torch.cuda.synchronize(); t = [time.perf_counter()]
    ctx.seek(start); ctx.set_line_window(lo // d, hi // d); t.append(time.perf_counter())
    push(2 * start, 2 * lo); l = ctx.take_lines(2); t.append(time.perf_counter())
    ds = hashlib.sha256(ctx.boundary_state()).digest(); t.append(time.perf_counter())
    push(2 * lo, 2 * hi); l += ctx.take_lines(2); t.append(time.perf_counter())
    de = hashlib.sha256(ctx.boundary_state()).digest(); t.append(time.perf_counter())
    push(2 * hi, 2 * tail); ctx.poll_flush(); l += ctx.take_lines(2); t.append(time.perf_counter())
    names = ["seek", "halo push", "state@lo", "chunk push", "state@hi", "tail push+flush"]
    print("  ".join("%s %.2f" % (nm, (b - a) * 1e3) for nm, a, b in zip(names, t, t[1:])), " total %.2f ms, %d lines" % ((t[-1] - t[0]) * 1e3, len(l)), flush=True)
