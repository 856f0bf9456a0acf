import csv, collections, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i,r in enumerate(rows) if r and r[0]=='ID')
hdr = rows[hdr_i]; data = rows[hdr_i+1:]
ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in data:
    if len(r) <= vi: continue
    v = float(r[vi].replace(',',''))
    if r[ui]=='ns': v/=1e6
    elif r[ui]=='us': v/=1e3
    name = r[ki].split('(')[0]
    if name.startswith('void at::'): continue
    a = agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v
tot = sum(a[1] for a in agg.values())
for k,a in sorted(agg.items(), key=lambda x:-x[1][1])[:14]: print(f"{k:45s} n={a[0]:4d} per={a[1]/a[0]:8.3f} ms share={a[1]/tot:6.1%}")
if len(sys.argv) > 2:
    j = json.load(open(sys.argv[2]))
    print("value", j['value'], "ms/step", j['ms_per_step'], "e2e", j['e2e']['value'], j['e2e']['ms_per_step'], "k1", j['roofline']['k1_demod_ms'], "k2", j['roofline']['k2_bitsync_ms'], "frac", j['roofline']['frac'], j['lanes'], j['packets'])
