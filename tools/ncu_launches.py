#!/usr/bin/env python
"""Print every launch of an `ncu --csv --metrics gpu__time_duration.sum,...` log: time, fmaheavy %, lanes, DRAM bytes.
Usage: python tools/ncu_launches.py LOG.csv [first_id]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
per = defaultdict(dict)
for r in rows[hi + 1:]:
    d = dict(zip(hdr, r))
    k = int(d["ID"])
    per[k]["name"] = d["Kernel Name"].split("(")[0].replace("cdl::", "")
    per[k][d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
tot = 0.0
for k in sorted(per):
    if k < first:
        continue
    p = per[k]
    ms = p["gpu__time_duration.sum"] / 1e6
    tot += ms
    print("%4d %-34s %8.3f ms  fmaheavy %5.1f %%  lanes %4.1f  dram rd %8.1f MB wr %8.1f MB" % (
        k, p["name"][:34], ms, p.get("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", 0),
        p.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0), p.get("dram__bytes_read.sum", 0) / 1e6,
        p.get("dram__bytes_write.sum", 0) / 1e6))
print("total %.3f ms" % tot)
