#!/usr/bin/env python
"""Single-proof latency through the C ABI (BASELINE.json configs 1-3): curdleproof.Prove and
Verify for shuffled_elements = 60 / 124 / 508, one proof at a time (B = 1), wall clock."""
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    pkg = importlib.import_module("go-curdleproofs_b200")
    ctx = pkg.Context(0)
    ells = [int(x) for x in sys.argv[1:]] or [60, 124, 508]
    for ell in ells:
        r = pkg.Rand(0)
        crs = ctx.generate_crs(ell, r)
        k = r.get_fr()
        Rs = ctx.rand_get_g1_affines(r, ell)
        Ss = ctx.rand_get_g1_affines(r, ell)
        perm = pkg.Rand(42).generate_permutation(ell)
        Ts, Us, M, rs_m = ctx.shuffle_permute_commit(crs, Rs, Ss, perm, k, r)
        best_p, best_v = None, None
        for it in range(4):
            t0 = time.perf_counter()
            proof = ctx.prove(crs, Rs, Ss, Ts, Us, M, perm, k, rs_m, pkg.Rand(42))
            t1 = time.perf_counter()
            ok = ctx.verify(crs, proof, Rs, Ss, Ts, Us, M, pkg.Rand(43))
            t2 = time.perf_counter()
            assert ok
            if it:
                best_p = t1 - t0 if best_p is None else min(best_p, t1 - t0)
                best_v = t2 - t1 if best_v is None else min(best_v, t2 - t1)
        print(json.dumps({"shuffled_elements": ell, "prove_ms": round(best_p * 1e3, 2), "verify_ms": round(best_v * 1e3, 2),
                          "proof_bytes": len(proof)}), flush=True)


if __name__ == "__main__":
    main()
