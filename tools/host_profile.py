#!/usr/bin/env python
"""Wall-clock split of batched proving / validation between host work and GPU stages (CDL_PROFILE=1
prints one line per lane engine when the context closes).  Usage:
  [taskset -c 0-3] python tools/host_profile.py [B] [lanes] [passes]"""
import importlib
import os
import resource
import sys
import time

os.environ["CDL_PROFILE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 8
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 3
pkg = importlib.import_module("go-curdleproofs_b200")
ctx = pkg.Context(0)   # generation (its profile lines print first)
ctx.set_lanes(lanes)
ctxv = pkg.Context(0)  # validation
ctxv.set_lanes(lanes)
crs = ctx.generate_crs(bench.ELL, pkg.Rand(0))
crsv = ctxv.generate_crs(bench.ELL, pkg.Rand(0))
sets = bench.make_trackers(ctx, pkg, bench.ELL, [1000 + i for i in range(4)])
pre = b"".join(sets[i % 4] for i in range(B))
rg = [[pkg.Rand((p << 20) | i) for i in range(B)] for p in range(passes + 1)]
rv = [[pkg.Rand((1 << 30) | (p << 20) | i) for i in range(B)] for p in range(passes + 1)]
def cpu_s():
    r = resource.getrusage(resource.RUSAGE_SELF)
    return r.ru_utime + r.ru_stime


tg, tv, cg, cv = [], [], [], []
for p in range(passes + 1):
    c0 = cpu_s()
    t0 = time.perf_counter()
    post, proofs, st = ctx.whisk_generate_shuffle_proof_batch(crs, pre, rg[p])
    t1 = time.perf_counter()
    c1 = cpu_s()
    ok, vs = ctxv.whisk_is_valid_shuffle_proof_batch(crsv, pre, post, proofs, rv[p])
    t2 = time.perf_counter()
    c2 = cpu_s()
    assert st == [0] * B and ok == [1] * B
    if p:
        tg.append(t1 - t0)
        tv.append(t2 - t1)
        cg.append(c1 - c0)
        cv.append(c2 - c1)
print(f"host CPU (user + sys, all threads) per proof: generate {1e3 * sum(cg) / len(cg) / B:.3f} ms, "
      f"validate {1e3 * sum(cv) / len(cv) / B:.3f} ms")
print(f"cores {len(os.sched_getaffinity(0))} B {B} lanes {lanes}: generate {B / (sum(tg) / len(tg)):.0f} proofs/s, "
      f"validate {B / (sum(tv) / len(tv)):.0f} /s, round trip {B / ((sum(tg) + sum(tv)) / len(tg)):.0f} /s")
sys.stdout.flush()
sys.stderr.write("--- generation engines\n")
ctx.close()
sys.stderr.write("--- validation engines\n")
ctxv.close()
