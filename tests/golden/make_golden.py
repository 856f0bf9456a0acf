#!/usr/bin/env python
"""Regenerates tests/golden/* from the UNMODIFIED reference binary (oracle/_ref/rtl_wmbus, built from
/root/reference by oracle/Makefile).  Run in the authoring container only:

    python tests/golden/make_golden.py

Outputs (committed):
  excerpt_*.cu8        short excerpts of the reference's sample captures around their telegrams, and
                       small synthetic captures carrying S1 / C1-B telegrams (no such fixture exists
                       upstream)
  golden_lines.json    {fixture: {flags: [lines with the TIMESTAMP column blanked]}}
  full_capture_sha.json  sha256 of the blanked output for the four full-size sample captures
"""
import hashlib
import importlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc  # noqa: E402

synth = importlib.import_module("rtl-wmbus_b200.synth")

FLAG_SETS_1M6 = ["", "-v", "-o", "-a", "-a -o -v", "-r 0", "-t 0", "-p S", "-p T", "-v -d 1", "-v -s"]
FLAG_SETS_2M4 = ["-d 3", "-d 3 -s", "-d 3 -s -o", "-d 3 -s -o -a -v", "-d 3 -s -o -p S"]
FULL_FLAGS = ["", "-v", "-o", "-a", "-a -o", "-r 0", "-t 0", "-d 3", "-p T", "-p S", "-s", "-d 1",
              "-d 3 -s -o", "-d 3 -s -o -v", "-d 3 -s"]


def excerpt(path, start_bytes, n_bytes):
    d = np.fromfile(path, np.uint8)
    return d[start_bytes:start_bytes + n_bytes]


def main():
    sdir = os.path.join(ROOT, "oracle", "_ref", "samples")
    fixtures = {}
    # excerpts (4096-aligned) around telegrams of the reference captures
    fixtures["excerpt_samples2_a.cu8"] = (excerpt(f"{sdir}/rtlsdr_868.950M_1M6_samples2.cu8", 0, 1 << 20), FLAG_SETS_1M6)
    fixtures["excerpt_issue47_c1.cu8"] = (excerpt(f"{sdir}/rtlsdr_868.950M_1M6_issue47.cu8", 3 << 20, 3 << 19), FLAG_SETS_1M6)
    fixtures["excerpt_issue48_2m4.cu8"] = (excerpt(f"{sdir}/rtlsdr_868.625M_2M4_issue48.cu8", 0, 204800), FLAG_SETS_2M4)
    # synthetic: every telegram type incl. S1 and C1 frame B
    em = synth.default_emitters("mixed")
    buf, _ = synth.synth_capture(3 << 19, fs=1.6e6, emitters=em, seed=0xB2000001)
    fixtures["synth_mixed_1m6.cu8"] = (buf.numpy(), FLAG_SETS_1M6)
    buf, _ = synth.synth_capture(3 << 19, fs=2.4e6, emitters=em, seed=0xB2000002, center_shift_hz=325e3)
    fixtures["synth_mixed_2m4_shift.cu8"] = (buf.numpy(), FLAG_SETS_2M4)

    lines = {}
    for name, (data, flagsets) in fixtures.items():
        data = np.ascontiguousarray(data)
        assert len(data) % 4096 == 0, name
        data.tofile(os.path.join(HERE, name))
        lines[name] = {fl: orc.ref_lines(data.tobytes(), fl) for fl in flagsets}
        print(name, len(data), {fl: len(v) for fl, v in lines[name].items()})
    json.dump(lines, open(os.path.join(HERE, "golden_lines.json"), "w"), indent=0, sort_keys=True)

    full = {}
    for f in sorted(os.listdir(sdir)):
        data = open(os.path.join(sdir, f), "rb").read()
        full[f] = {}
        for fl in FULL_FLAGS:
            out = orc.ref_lines(data, fl)
            full[f][fl] = {"n": len(out), "sha256": hashlib.sha256("\n".join(out).encode()).hexdigest()}
    json.dump(full, open(os.path.join(HERE, "full_capture_sha.json"), "w"), indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
