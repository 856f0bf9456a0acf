"""Device-resident 1 GiB steps under different pipeline settings (experiments):
   WMBUS_B200_PIPE_MIB (batches of a long device push), WMBUS_B200_K1_CTAS (resident demod blocks per SM).
   python tools/pipe_sweep.py [workload]"""
import importlib, os, sys, time
sys.path.insert(0, '.')
import torch
pkg = importlib.import_module("rtl-wmbus_b200"); synth = importlib.import_module("rtl-wmbus_b200.synth")
lib = pkg.load_library()
wl = sys.argv[1] if len(sys.argv) > 1 else "t1x2"
flags = {"t1x2": "-p S", "s1": "-p T", "both": ""}[wl]
n = 1 << 30
cap, plan = synth.synth_capture(n, emitters=synth.default_emitters("mixed" if wl == "both" else wl), seed=0xB2000020, device="cuda")
torch.cuda.synchronize()
ref = None
for pipe, ctas in [(0, 0), (0, 5), (0, 4), (512, 0), (512, 5), (512, 4), (384, 5), (256, 5), (256, 4), (342, 5)]:
    if pipe: os.environ["WMBUS_B200_PIPE_MIB"] = str(pipe)
    else: os.environ.pop("WMBUS_B200_PIPE_MIB", None)
    os.environ["WMBUS_B200_K1_CTAS"] = str(ctas)
    with pkg.WmbusB200(flags, lib=lib, max_batch_mib=1024) as ctx:
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
        print("pipe %4d MiB  k1 ctas %d : wall %.3f ms  pass %.3f  demod %.3f  bit-sync(sum) %.3f  batches %d  lines %s" % (
            pipe, ctas, best[0], best[1], best[2], best[3], st.batches // 6, "same" if lines == ref else "DIFFERENT"), flush=True)
