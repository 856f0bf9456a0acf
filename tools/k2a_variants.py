"""Time device-resident 1 GiB steps with several builds of the library (experiments):
   python tools/k2a_variants.py rtl-wmbus_b200/libwmbus_b200_sk4.so[:chunk_samples] ...   prints demod / bit-sync / pass ms per build"""
import importlib, sys, time
sys.path.insert(0, '.')
import torch
pkg = importlib.import_module("rtl-wmbus_b200"); synth = importlib.import_module("rtl-wmbus_b200.synth")
n = 1 << 30
cap, plan = synth.synth_capture(n, emitters=synth.default_emitters("t1x2"), seed=0xB2000020, device="cuda")
torch.cuda.synchronize()
ref = None
for arg in sys.argv[1:]:
    path, _, chunk = arg.partition(":")                    # path[:chunk_samples]
    lib = pkg.load_library(path)
    kw = dict(chunk_samples=int(chunk)) if chunk else {}
    with pkg.WmbusB200("-p S", lib=lib, max_batch_mib=1024, **kw) as ctx:
        best = None
        for i in range(5):
            ctx.reset()
            t0 = time.perf_counter()
            lines = ctx.process_device(cap.data_ptr(), n, flush=True)
            dt = (time.perf_counter() - t0) * 1e3
            st = ctx.stats()
            row = (st.batch_device_ms, st.demod_kernel_ms, st.bitsync_kernel_ms, dt)
            best = row if best is None or row[0] < best[0] else best
        ref = ref or lines
        print("%-50s pass %.3f ms  demod %.3f  bit-sync %.3f  wall %.3f  lines %s  reruns %d" % (
            arg, best[0], best[1], best[2], best[3], "same" if lines == ref else "DIFFERENT", st.lanes_rerun), flush=True)
