"""A few 100 ms hand-overs (both chains, host pushes) for ncu launch lists: python tools/small_push.py [piece_bytes]"""
import importlib, sys, time
sys.path.insert(0, '.')
import torch
pkg = importlib.import_module("rtl-wmbus_b200"); synth = importlib.import_module("rtl-wmbus_b200.synth")
lib = pkg.load_library()
piece = int(sys.argv[1]) if len(sys.argv) > 1 else 319488
n = piece * 12
cap, plan = synth.synth_capture((n + 4095) // 4096 * 4096, emitters=synth.default_emitters("mixed"), seed=0xB2000051, device="cuda")
host = torch.empty(cap.numel(), dtype=torch.uint8, pin_memory=True); host.copy_(cap); torch.cuda.synchronize()
with pkg.WmbusB200("", lib=lib, max_batch_mib=64) as ctx:
    for k in range(12):
        t0 = time.perf_counter()
        ctx.push(host.data_ptr() + k * piece, piece)
        lines = ctx.take_lines()
        print("push %d: %.3f ms, %d lines" % (k, (time.perf_counter() - t0) * 1e3, len(lines)), file=sys.stderr)
