#!/usr/bin/env python
"""profiles/traffic.json from an ncu launch list with dram__bytes_read.sum + dram__bytes_write.sum:
DRAM bytes per launch of each kernel class of bench.py (a class launch = one engine stage, e.g. the
four kernels of one MSM chain).  Usage: python tools/traffic.py launches.csv BATCH"""
import csv
import json
import os
import sys
from collections import defaultdict

path, batch = sys.argv[1], int(sys.argv[2])
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
CLASS = {"k_msm_recode": "msm_small", "k_msm_warp": "msm_small", "k_msm_warp_gmem": "msm_small", "k_msm_chunk_sum": "msm_small",
         "k_msm_combine_tp": "msm_small", "k_msm_combine_quad": "msm_small", "k_msm_small": "msm_small",
         "k_msm_fixed": "msm_small",
         "k_elem_ops": "elem_scalar_mul", "k_elem_ops_quad": "elem_scalar_mul",
         "k_decompress_idx": "decompress", "k_compress_idx": "compress", "k_jac_to_affine": "compress"}
STAGE_END = {"msm_small": ("k_msm_combine_tp", "k_msm_combine_quad", "k_msm_small"),
             "elem_scalar_mul": ("k_elem_ops", "k_elem_ops_quad"),
             "decompress": ("k_decompress_idx",), "compress": ("k_compress_idx", "k_jac_to_affine")}
bytes_ = defaultdict(float)
stages = defaultdict(int)
unit_scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
seen = set()
for r in data:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"].split("(")[0].replace("cdl::", "").replace("void ", "")
    cls = CLASS.get(name)
    if not cls:
        continue
    if name in STAGE_END[cls] and d["ID"] not in seen:
        seen.add(d["ID"])
        stages[cls] += 1
    if "dram__bytes" not in d["Metric Name"]:
        continue
    bytes_[cls] += float(d["Metric Value"].replace(",", "")) * unit_scale.get(d["Metric Unit"], 1)
out = {"batch": batch, "source": os.path.basename(path),
       "what": "dram__bytes_read.sum + dram__bytes_write.sum per class launch (ncu, lanes = 1)"}
for cls in bytes_:
    out[cls] = bytes_[cls] / max(1, stages[cls])
    out[cls + "_launches"] = stages[cls]
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
