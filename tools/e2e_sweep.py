import importlib, sys, time, json
sys.path.insert(0, '.')
import torch
pkg = importlib.import_module("rtl-wmbus_b200"); synth = importlib.import_module("rtl-wmbus_b200.synth")
lib = pkg.load_library()
n = 1 << 30
cap, plan = synth.synth_capture(n, emitters=synth.default_emitters("t1x2"), seed=0xB2000020, device="cuda")
host = torch.empty(n, dtype=torch.uint8, pin_memory=True); host.copy_(cap); torch.cuda.synchronize()
dst = torch.empty_like(cap)
for _ in range(2):
    t0 = time.perf_counter(); dst.copy_(host, non_blocking=True); torch.cuda.synchronize(); t = time.perf_counter() - t0
print("H2D pinned 1 GiB: %.2f ms = %.1f GB/s" % (t * 1e3, n / t / 1e9))
for mib in (64, 128, 256, 512, 1024):
    ctx = pkg.WmbusB200("-p S", lib=lib, max_batch_mib=mib)
    for _ in range(2):
        ctx.reset(); ctx.process(host.data_ptr(), n, flush=True)
    t0 = time.perf_counter()
    for _ in range(4):
        ctx.reset(); lines = ctx.process(host.data_ptr(), n, flush=True)
    t = (time.perf_counter() - t0) / 4
    st = ctx.stats()
    print("e2e batch %4d MiB: %.2f ms/step = %.0f Msps; host ms per step: batch %.2f gather %.2f decode %.2f; lines %d" % (
        mib, t * 1e3, n / 2 / t / 1e6, st.host_batch_ms / 6, st.host_gather_ms / 6, st.host_decode_ms / 6, len(lines)))
    ctx.close()
