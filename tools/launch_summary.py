#!/usr/bin/env python
"""Summarise an `ncu --csv --metrics gpu__time_duration.sum,...` launch list: per-kernel totals over the
second half of the launches (the timed step) and, with -v, every launch."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
per = defaultdict(dict)
for r in data:
    d = dict(zip(hdr, r))
    k = int(d["ID"])
    per[k]["name"] = d["Kernel Name"].split("(")[0].replace("cdl::", "")
    per[k]["grid"] = d["Grid Size"]
    per[k][d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
ids = sorted(per)
ids = ids[len(ids) // 2:]
agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
tot = 0.0
for i in ids:
    p = per[i]
    t = p["gpu__time_duration.sum"] / 1e6
    a = agg[p["name"]]
    a[0] += 1
    a[1] += t
    a[2] += t * p.get("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", 0)
    a[3] += t * p.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0)
    tot += t
print(f"{'kernel':40s} {'n':>4s} {'ms':>9s} {'share':>6s} {'fmaheavy%':>9s} {'lanes':>6s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:40]:40s} {v[0]:4d} {v[1]:9.3f} {v[1] / tot:6.3f} {v[2] / max(v[1], 1e-9):9.1f} {v[3] / max(v[1], 1e-9):6.1f}")
print(f"total {tot:.3f} ms over {len(ids)} launches")
if "-v" in sys.argv:
    for i in ids:
        p = per[i]
        print(i, p["name"], p["grid"], round(p["gpu__time_duration.sum"] / 1e6, 3), "lanes",
              round(p.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0), 1), "fmah",
              round(p.get("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", 0), 1), "warps",
              round(p.get("sm__warps_active.avg.per_cycle_active", 0), 1))
