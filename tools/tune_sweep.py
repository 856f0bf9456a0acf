"""Lane-geometry sweep of the bit-sync kernels (WMBUS_B200_TUNE), device-resident 1 GiB capture.
Usage: python tools/tune_sweep.py [workload]      (run on the GPU box; prints one line per setting)"""
import hashlib
import importlib
import os
import sys
import time

sys.path.insert(0, '.')
import torch

pkg = importlib.import_module("rtl-wmbus_b200")
synth = importlib.import_module("rtl-wmbus_b200.synth")
lib = pkg.load_library()
workload = sys.argv[1] if len(sys.argv) > 1 else "t1x2"
flags = {"t1x2": "-p S", "s1": "-p T", "both": ""}[workload]
n = 1 << 30
cap, plan = synth.synth_capture(n, emitters=synth.default_emitters("mixed" if workload == "both" else workload),
                                seed=0xB2000020, device="cuda")
torch.cuda.synchronize()
ref = None
settings = ["128:2048:512", "64:2048:512", "32:2048:512", "128:1024:512", "128:2048:256", "128:2048:128",
            "64:1024:256", "32:1024:128", "32:1024:64", "16:1024:128"]
for tune in settings:
    os.environ["WMBUS_B200_TUNE"] = tune
    ctx = pkg.WmbusB200(flags, lib=lib, max_batch_mib=1024)
    for _ in range(2):
        ctx.reset(); ctx.process_device(cap.data_ptr(), n, flush=True)
    ctx.reset()
    st0 = ctx.stats()
    t0 = time.perf_counter()
    for _ in range(4):
        ctx.reset(); lines = ctx.process_device(cap.data_ptr(), n, flush=True)
    t = (time.perf_counter() - t0) / 4
    st = ctx.stats()
    dig = hashlib.sha256("\n".join(";".join(l.split(";")[:3] + l.split(";")[4:]) for l in lines).encode()).hexdigest()[:12]
    if ref is None:
        ref = dig
    print("tune %-14s step %.2f ms; k1 %.2f k2 %.2f batch-dev %.2f ms; lines %d sha %s %s" % (
        tune, t * 1e3, st.demod_kernel_ms / 6, st.bitsync_kernel_ms / 6, st.batch_device_ms / 6, len(lines), dig,
        "same" if dig == ref else "DIFFERENT"), flush=True)
    ctx.close()
