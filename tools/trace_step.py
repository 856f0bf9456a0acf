"""Host wall-clock trace of device-resident steps (WMBUS_B200_TRACE): python tools/trace_step.py [mib] [batch_mib] [workload]"""
import importlib, os, sys, time
sys.path.insert(0, '.')
os.environ["WMBUS_B200_TRACE"] = "1"
import torch
pkg = importlib.import_module("rtl-wmbus_b200"); synth = importlib.import_module("rtl-wmbus_b200.synth")
lib = pkg.load_library()
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
bm = int(sys.argv[2]) if len(sys.argv) > 2 else mib
n = mib << 20
workload = sys.argv[3] if len(sys.argv) > 3 else "t1x2"
flags = {"t1x2": "-p S", "s1": "-p T", "both": ""}[workload]
cap, plan = synth.synth_capture(n, emitters=synth.default_emitters("mixed" if workload == "both" else workload), seed=0xB2000020, device="cuda")
torch.cuda.synchronize()
ctx = pkg.WmbusB200(flags, lib=lib, max_batch_mib=bm)
for i in range(4):
    t0 = time.perf_counter(); ctx.reset(); t1 = time.perf_counter()
    lines = ctx.process_device(cap.data_ptr(), n, flush=True)
    t2 = time.perf_counter()
    print("step %d: reset %.3f ms, process_device %.3f ms, %d lines" % (i, (t1 - t0) * 1e3, (t2 - t1) * 1e3, len(lines)), file=sys.stderr)
