#!/usr/bin/env python
"""Which kernels of two builds of libwmbus_b200.so are the same machine code?  (No GPU needed.)
    git archive <commit> rtl-wmbus_b200/csrc include | tar -x -C /tmp/old && make -C /tmp/old/rtl-wmbus_b200/csrc ../libwmbus_b200.so
    python profiles/sass_diff.py /tmp/old/rtl-wmbus_b200/libwmbus_b200.so rtl-wmbus_b200/libwmbus_b200.so
Per kernel: `same` when the SASS (addresses and encodings stripped) is identical instruction for instruction, else the
opcodes whose counts differ.  Used to carry a profile (launch list, ncu capture, bench line) taken on one commit over to
a later one: a kernel that is `same` runs exactly as profiled."""
import collections
import re
import subprocess
import sys


def funcs(lib):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    out, name = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            out[name] = []
        elif name and re.match(r"\s*/\*[0-9a-f]{4,5}\*/", line):
            l = re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line).rstrip()
            out[name].append(re.sub(r"^\s*/\*[0-9a-f]+\*/\s*", "", l))
    return out


def op(l):
    m = re.match(r"(?:@!?U?P[0-9T]\s+)?([A-Z0-9_.]+)", l)
    return m.group(1) if m else "?"


def demangle(n):
    s = subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip()
    return re.sub(r"\([^()]*\)$", "", s).replace("void ", "").replace("(unsigned int)", "")


a, b = funcs(sys.argv[1]), funcs(sys.argv[2])
print("%-8s %7s %7s  kernel" % ("", "old", "new"))
for n in sorted(set(a) | set(b), key=demangle):
    if n not in a:
        print("%-8s %7s %7d  %s" % ("new", "-", len(b[n]), demangle(n)))
    elif n not in b:
        print("%-8s %7d %7s  %s" % ("gone", len(a[n]), "-", demangle(n)))
    elif a[n] == b[n]:
        print("%-8s %7d %7d  %s" % ("same", len(a[n]), len(b[n]), demangle(n)))
    else:
        ha, hb = collections.Counter(map(op, a[n])), collections.Counter(map(op, b[n]))
        d = ", ".join("%s %d->%d" % (k, ha[k], hb[k]) for k in sorted(set(ha) | set(hb)) if ha[k] != hb[k])
        print("%-8s %7d %7d  %s\n%26s%s" % ("differs", len(a[n]), len(b[n]), demangle(n), "", d or "operands only (parameter offsets, branch targets)"))
