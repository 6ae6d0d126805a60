#!/usr/bin/env python
"""Standalone G1 MSM sweep (BASELINE.json config 5) on one GPU: N = 2^lo .. 2^hi,
device-resident inputs, CUDA-event kernel time, algorithmic modmul/s against the
integer-pipe peak.  Usage: python tools/msm_sweep.py [lo hi reps] [--c C | --c C1,C2 | --scan-c] [--ba R1,R2,..]
(--ba: batch-affine rounds to try, -1 = chosen by bucket load)"""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def alg_modmul(n):
    best = None
    for c in range(1, 25):
        W = -(-256 // c)
        v = 6 * n * W + 28 * (1 << (c - 1)) * W + 9 * c * (W - 1) + 14 * W
        best = v if best is None or v < best else best
    return best


def main():
    args = [a for i, a in enumerate(sys.argv[1:], 1) if not a.startswith("--") and sys.argv[i - 1] not in ("--c", "--ba")]
    lo, hi, reps = (int(args[0]), int(args[1]), int(args[2])) if len(args) >= 3 else (10, 22, 3)
    cs = [int(x) for x in sys.argv[sys.argv.index("--c") + 1].split(",")] if "--c" in sys.argv else [0]
    bas = [int(x) for x in sys.argv[sys.argv.index("--ba") + 1].split(",")] if "--ba" in sys.argv else [-1]
    if "--scan-c" in sys.argv:
        cs = None
    pkg = importlib.import_module("go-curdleproofs_b200")
    enc = importlib.import_module("go-curdleproofs_b200.encoding")
    aff_enc = enc.aff_enc

    class b:  # generator only
        G1_GEN = enc.G1_GEN

    ctx = pkg.Context(0)
    nmax = 1 << hi
    r5, r6 = pkg.Rand(5), pkg.Rand(6)
    dp = ctx.dev_buffer(96 * nmax)
    da = ctx.dev_buffer(32 * nmax)
    ds = ctx.dev_buffer(32 * nmax)
    gen = aff_enc(b.G1_GEN)
    chunk = 1 << 16
    for o in range(0, nmax, chunk):
        m = min(chunk, nmax - o)
        dp.upload(gen * m, 96 * o)
        da.upload(r5.get_frs(m), 32 * o)
        ds.upload(r6.get_frs(m), 32 * o)
    ctx.g1_scalar_mul_affine_device(dp, da, nmax, False, dp)
    peak, _ = ctx.int_peak(1, 4000)
    peak_mm = peak / 300.0
    for lg in range(lo, hi + 1):
        n = 1 << lg
        for c in (cs if cs is not None else range(max(6, lg // 2 + 1), min(18, lg) + 1)):
            ctx.set_msm_window(c)
            ref = None
            for ba in bas:
                ctx.set_msm_batch_affine(ba)
                best = None
                for _ in range(reps + 1):
                    out, ms = ctx.g1_msm_device(dp, ds, n)
                    best = ms if best is None or ms < best else best
                ref = out if ref is None else ref
                mm = alg_modmul(n)
                print(json.dumps({"log2n": lg, "c": c, "ba": ba, "ms": round(best, 4),
                                  "mpoints_per_s": round(n / best / 1e3, 2), "gmodmul_per_s": round(mm / best / 1e6, 3),
                                  "frac_of_int_peak": round(mm / (best * 1e-3) / peak_mm, 4),
                                  "same_point": out == ref}), flush=True)
    ctx.set_msm_window(0)
    ctx.set_msm_batch_affine(-1)


if __name__ == "__main__":
    main()
