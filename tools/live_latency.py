"""Live-stream cadence: push a capture in small pieces, as the CLI does with `rtl_sdr |` every 100 ms,
and time each wmb_push + wmb_take_lines.  python tools/live_latency.py   (GPU box)"""
import importlib, sys, time
sys.path.insert(0, '.')
import numpy as np
import torch
pkg = importlib.import_module("rtl-wmbus_b200"); synth = importlib.import_module("rtl-wmbus_b200.synth")
lib = pkg.load_library()
n = 64 << 20
cap, plan = synth.synth_capture(n, emitters=synth.default_emitters("mixed"), seed=0xB2000051, device="cuda")
host = torch.empty(n, dtype=torch.uint8, pin_memory=True); host.copy_(cap); torch.cuda.synchronize()
with pkg.WmbusB200("", lib=lib) as ctx:
    want = ctx.process(host.data_ptr(), n, flush=True)
for ms in (100, 50, 20, 10):
    piece = int(1.6e6 * 2 * ms / 1000) // 8192 * 8192          # bytes per hand-over at 1.6 MS/s
    with pkg.WmbusB200("", lib=lib, max_batch_mib=64) as ctx:
        lat, lines = [], []
        for off in range(0, n, piece):
            t0 = time.perf_counter()
            ctx.push(host.data_ptr() + off, min(piece, n - off))
            lines += ctx.take_lines()
            lat.append((time.perf_counter() - t0) * 1e3)
        ctx.poll_flush(); lines += ctx.take_lines()
    lat = np.array(lat[3:])
    print("hand-over every %3d ms (%7d bytes): push+lines latency median %.2f ms, p99 %.2f ms, max %.2f ms; GPU busy %.1f %% of real time; lines %s" % (
        ms, piece, np.median(lat), np.percentile(lat, 99), lat.max(), 100 * np.median(lat) / ms, "identical" if lines == want else "DIFFERENT"))
