import importlib, json, sys
sys.path.insert(0, '.')
pkg = importlib.import_module("go-curdleproofs_b200")
ctx = pkg.Context(0)
info = ctx.device_info()
res = {"device": info}
for kind, name, iters in ((0, "imad32", 4000), (1, "imad_wide", 4000), (2, "modmul_chain", 2000)):
    ops, ms = ctx.int_peak(kind, iters)
    res[name] = {"ops_per_s": ops, "ms": ms}
print(json.dumps(res))
