#!/usr/bin/env python
"""Derive the numbers bench.py reports from an ncu capture -- nothing typed in by hand.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv
    python profiles/ncu_extract.py raw.csv profiles/r2_ncu.json [profiles/r2_ncu_full_summary.txt]

Input: the `--page raw --csv` dump of an `ncu --set full --clock-control none` capture of ONE 1 GiB device-resident
step (tools/trace_step.py).  Output JSON: per kernel (launches of the same kernel are summed -- a step launches some
kernels more than once) time, DRAM bytes, issue-active %, warps-active %, instruction count and the top stall reasons,
plus `traffic_bytes_per_step` = sum over the captured kernels of dram__bytes_read.sum + dram__bytes_write.sum.
bench.py copies `traffic_bytes_per_step` into roofline.traffic and the per-kernel issue/DRAM percentages into
roofline.ncu.  The optional third argument writes the human-readable excerpt kept under profiles/.
"""
import collections
import csv
import json
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
STALL = "smsp__average_warps_issue_stalled_"
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def main():
    rows = list(csv.reader(open(sys.argv[1], newline="")))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units = rows[hi], rows[hi + 1]
    col = {n: i for i, n in enumerate(hdr)}
    ki = col["Kernel Name"]
    per = collections.OrderedDict()
    for r in rows[hi + 2:]:
        if len(r) <= ki:
            continue
        name = r[ki].split("(")[0].replace("void ", "")
        k = per.setdefault(name, {"launches": 0, "ms": 0.0, "dram_read": 0.0, "dram_write": 0.0, "inst": 0.0, "rows": []})
        k["launches"] += 1
        k["rows"].append(r)
        g = lambda m: (num(r[col[m]]) or 0.0) * UNIT.get(units[col[m]], 1.0) if m in col else 0.0
        k["ms"] += g("gpu__time_duration.sum")
        k["dram_read"] += g("dram__bytes_read.sum")
        k["dram_write"] += g("dram__bytes_write.sum")
        k["inst"] += g("smsp__inst_executed.sum")
    out = {"source": sys.argv[1], "kernels": {}, "traffic_bytes_per_step": 0}
    text = []
    for name, k in per.items():
        big = max(k["rows"], key=lambda r: num(r[col["gpu__time_duration.sum"]]) or 0.0)      # the longest launch speaks for the kernel
        pct = lambda m: num(big[col[m]]) if m in col else None
        stalls = sorted(((num(big[i]) or 0.0, hdr[i][len(STALL):].replace("_per_issue_active.ratio", ""))
                         for i in range(len(hdr)) if hdr[i].startswith(STALL) and hdr[i].endswith("_per_issue_active.ratio")),
                        reverse=True)[:5]
        out["kernels"][name] = {
            "launches": k["launches"], "ms": round(k["ms"], 4),
            "dram_read_bytes": int(k["dram_read"]), "dram_write_bytes": int(k["dram_write"]),
            "inst_executed": int(k["inst"]),
            "issue_active_pct": pct("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "dram_pct": pct("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "warps_active_pct": pct("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "threads_per_inst": pct("smsp__thread_inst_executed_per_inst_executed.ratio"),
            "registers": pct("launch__registers_per_thread"),
            "stalls_per_issue": {n: round(v, 2) for v, n in stalls},
        }
        out["traffic_bytes_per_step"] += int(k["dram_read"] + k["dram_write"])
        text.append(f"Kernel Name  {name}   ({k['launches']} launch(es), {k['ms']:.4f} ms, "
                    f"dram {k['dram_read'] / 1e9:.3f} GB read + {k['dram_write'] / 1e9:.3f} GB written)")
        for m in WANT:
            if m in col:
                text.append(f"  {m:68s} {big[col[m]]} {units[col[m]]}")
        for v, n in stalls:
            text.append(f"     stall {v:6.2f} {n}")
        text.append("")
    out["total_ms"] = round(sum(k["ms"] for k in per.values()), 4)
    json.dump(out, open(sys.argv[2], "w"), indent=1)
    if len(sys.argv) > 3:
        open(sys.argv[3], "w").write("\n".join(text) + "\n")
    print(f"{len(per)} kernels, {out['total_ms']} ms, traffic {out['traffic_bytes_per_step'] / 1e9:.3f} GB per step -> {sys.argv[2]}")


if __name__ == "__main__":
    main()
