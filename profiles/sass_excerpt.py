#!/usr/bin/env python
"""SASS evidence for the two largest kernels, from the in-tree library (no GPU needed):
    python profiles/sass_excerpt.py > profiles/r2_sass_excerpt.txt
* instruction histogram of k1_demod_kernel<1> and k2a2_lanes_kernel<ChainT1C1>
* the TMA / mbarrier instructions of K1 (UBLKCP = cp.async.bulk, SYNCS = mbarrier)
* every FFMA of K1 with the MUFU that precedes it: fused multiply-adds exist ONLY inside the IEEE division
  (MUFU.RCP + 4 FFMA + FMUL) and square-root (MUFU.RSQ + 2 FMUL + 2 FFMA) sequences -- the reference's arithmetic itself
  is never contracted (x86-64 baseline build: every multiply and add rounds separately)
* shared-memory accesses are LDS/STS (no generic LD/ST), the division has no FCHK / slow-path CALL
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "rtl-wmbus_b200", "libwmbus_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs = {}
name = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        funcs[name] = []
    elif name and re.match(r"\s*/\*[0-9a-f]{4,5}\*/", line):
        funcs[name].append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line).rstrip())

def op(l):
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P[0-9T]\s+)?([A-Z0-9_.]+)", l)
    return m.group(1) if m else "?"

for key, title in (("k1_demod_kernelILj1E", "k1_demod_kernel<1> (T1/C1 chain only, the benchmark's kernel)"),
                   ("k2a2_lanes_kernelI9ChainT1C1", "k2a2_lanes_kernel<ChainT1C1> (clock recovery, three threads per lane)")):
    fn = next(n for n in funcs if key in n)
    ins = funcs[fn]
    hist = collections.Counter(op(l).split(".")[0] for l in ins)
    print(f"== {title}: {len(ins)} instructions")
    print("   " + "  ".join(f"{k} {v}" for k, v in hist.most_common(28)))
    if "k1_demod" in fn:
        print("-- TMA bulk copies and mbarrier:")
        for l in ins:
            if re.search(r"UBLKCP|SYNCS|UTMA", l):
                print("   " + l.strip())
        print(f"-- shared memory: LDS {sum('LDS' in op(l) for l in ins)}, STS {sum('STS' in op(l) for l in ins)}, "
              f"generic LD {sum(op(l).startswith('LD.') or op(l) == 'LD' for l in ins)}, generic ST {sum(op(l).startswith('ST.') or op(l) == 'ST' for l in ins)}; "
              f"FCHK {hist['FCHK']}, CALL {hist['CALL']} (sqrt slow path of the d > 3 front end only)")
        print("-- every FFMA with the nearest MUFU before it (division / square-root sequences only):")
        last = None
        seen = collections.Counter()
        for l in ins:
            o = op(l)
            if o.startswith("MUFU"):
                last = o
            if o.startswith("FFMA"):
                seen[last] += 1
        print("   " + ", ".join(f"{v} FFMA after {k}" for k, v in seen.items()))
        i = next(k for k, l in enumerate(ins) if "MUFU.RCP" in l and "|" in l)
        print("-- one division (|y| / |x| of the discriminator), as emitted:")
        for l in ins[i:i + 6]:
            print("   " + l.strip())
    print()
