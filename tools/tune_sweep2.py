"""Lane-geometry sweep of the bit-sync kernels on device-resident 1 GiB steps (experiments):
   WMBUS_B200_TUNE = t2words:p1chunk:p2records, chunk_samples (clock-recovery lane length)."""
import importlib, os, sys, time
sys.path.insert(0, '.')
import torch
pkg = importlib.import_module("rtl-wmbus_b200"); synth = importlib.import_module("rtl-wmbus_b200.synth")
lib = pkg.load_library()
n = 1 << 30
cap, plan = synth.synth_capture(n, emitters=synth.default_emitters("t1x2"), seed=0xB2000020, device="cuda")
torch.cuda.synchronize()
ref = None
cfgs = [("128:2048:512", 0), ("128:2048:256", 0), ("128:2048:128", 0), ("128:2048:64", 0), ("32:2048:512", 0), ("64:2048:256", 0),
        ("128:1024:256", 0), ("128:4096:256", 0), ("128:2048:512", 22784), ("128:2048:512", 15360), ("128:2048:512", 30464), ("64:1024:128", 22784)]
for tune, chunk in cfgs:
    os.environ["WMBUS_B200_TUNE"] = tune
    kw = dict(chunk_samples=chunk) if chunk else {}
    with pkg.WmbusB200("-p S", lib=lib, max_batch_mib=1024, **kw) as ctx:
        best = None
        for i in range(6):
            ctx.reset()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            lines = ctx.process_device(cap.data_ptr(), n, flush=True)
            dt = (time.perf_counter() - t0) * 1e3
            st = ctx.stats()
            row = (dt, st.batch_device_ms, st.demod_kernel_ms, st.bitsync_kernel_ms)
            best = row if best is None or row[0] < best[0] else best
        ref = ref or lines
        print("tune %-14s chunk %6d : wall %.3f ms  pass %.3f  demod %.3f  bit-sync %.3f  lines %s reruns %d" % (
            tune, chunk, best[0], best[1], best[2], best[3], "same" if lines == ref else "DIFFERENT", st.lanes_rerun), flush=True)
