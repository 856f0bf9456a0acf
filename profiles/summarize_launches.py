"""Per-kernel table of an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py.
bench.py runs the device leg with one 1 GiB batch per step and the end-to-end leg with 256 MiB batches, so every
kernel shows up with two launch sizes: the largest launches are reported as "1 GiB batch", the rest as "256 MiB batch".
usage: summarize_launches.py launches.csv [bench.json] [steps_total=3]"""
import collections
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr = rows[hi]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3        # warm-up + timed steps of each leg
d = collections.defaultdict(list)
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(',', ''))
    v = v / 1e6 if r[ui] == 'ns' else v / 1e3 if r[ui] == 'us' else v
    name = r[ki].split('(')[0]
    if name.startswith('void at::') or 'elementwise' in name:
        continue
    d[name].append(v)
out, tot_big, tot_small = [], 0.0, 0.0
for k, v in d.items():
    per = max(1, len(v) // (5 * steps))                     # launches per batch: 1 big + 4 small batches per step
    big = sorted(v)[-steps * per:]
    small = sorted(v)[:len(v) - steps * per]
    b = sum(big) / steps
    s = sum(small) / (4 * steps) if small else 0.0
    out.append((b, k, per, s))
    tot_big += b
    tot_small += s
print(f"{'kernel':46s} {'launches':>8s} {'1 GiB batch':>12s} {'share':>7s} {'256 MiB batch':>14s}")
for b, k, per, s in sorted(out, reverse=True):
    print(f"{k:46s} {per:8d} {b:9.3f} ms {b / tot_big:7.1%} {s:11.3f} ms")
print(f"{'sum of kernels':46s} {'':8s} {tot_big:9.3f} ms {'':7s} {tot_small:11.3f} ms")
if len(sys.argv) > 2:
    j = json.load(open(sys.argv[2]))
    print("bench line of the same build: value", j['value'], "ms/step", j['ms_per_step'], "e2e", j['e2e']['value'], j['e2e']['ms_per_step'],
          "k1", j['roofline']['k1_demod_ms'], "k2", j['roofline']['k2_bitsync_ms'], "frac", j['roofline']['frac'], j['lanes'], j['packets'])
