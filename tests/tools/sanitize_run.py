"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck):
   compute-sanitizer --tool memcheck python tests/tools/sanitize_run.py"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
pkg = importlib.import_module("rtl-wmbus_b200")
import orc
lib = pkg.load_library()
for name, flags, kw in [("synth_mixed_1m6.cu8", "-v", {}), ("synth_mixed_2m4_shift.cu8", "-v -d 3 -s -o", {}),
                        ("synth_mixed_1m6.cu8", "-v", dict(max_batch_mib=1, chunk_samples=2048, warmup_samples=512))]:
    cu8 = np.fromfile(os.path.join(ROOT, "tests", "golden", name), np.uint8)
    with pkg.WmbusB200(flags, lib=lib, **kw) as ctx:
        lines = ctx.process(cu8.ctypes.data, len(cu8), flush=True)
    want = [orc.blank_ts(l) for l in orc.run_lines(cu8, orc.opts_from_flags(flags))]
    assert lines == want, (name, flags)
    print(name, flags, len(lines), "lines ok")
