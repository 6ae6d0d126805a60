import importlib, json, sys
sys.path.insert(0, '.')
pkg = importlib.import_module("go-curdleproofs_b200")
ctx = pkg.Context(0)
peak = ctx.int_peak(1, 4000)[0] / 300
for kind in (2, 3, 4):
    for bps, tpb in ((1, 32), (1, 128), (1, 256), (2, 256), (3, 256), (4, 256), (6, 256), (8, 256)):
        ops, ms = ctx.int_peak_cfg(kind, 1000, bps, tpb)
        print(json.dumps({"kind": kind, "warps_per_sm": bps * tpb // 32, "gmodmul_s": round(ops / 1e9, 2), "frac": round(ops / peak, 3), "ms": round(ms, 3)}))
