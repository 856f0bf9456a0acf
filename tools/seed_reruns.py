"""Lane re-runs and step time per capture seed (the weak-scaling ranks each decode their own capture):
   python tools/seed_reruns.py [n_seeds]"""
import importlib, sys, time
sys.path.insert(0, '.')
import torch
pkg = importlib.import_module("rtl-wmbus_b200"); synth = importlib.import_module("rtl-wmbus_b200.synth")
shard = importlib.import_module("rtl-wmbus_b200.shard")
lib = pkg.load_library()
n = 1 << 30
for r in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
    cap, plan = synth.synth_capture(n, emitters=synth.default_emitters("t1x2"), seed=shard.capture_seed(2, r), device="cuda")
    torch.cuda.synchronize()
    with pkg.WmbusB200("-p S", lib=lib, max_batch_mib=1024) as ctx:
        best = None
        for i in range(4):
            ctx.reset(); torch.cuda.synchronize()
            r0 = ctx.stats().lanes_rerun
            t0 = time.perf_counter()
            txt = ctx.process_device(cap.data_ptr(), n, flush=True, raw=True)
            dt = (time.perf_counter() - t0) * 1e3
            st = ctx.stats()
            best = dt if best is None or dt < best else best
        print("seed of rank %d: wall %.3f ms  pass %.3f  re-runs per step %d  lines %d" % (r, best, st.batch_device_ms, st.lanes_rerun - r0, txt.count(b"\n")), flush=True)
    del cap
