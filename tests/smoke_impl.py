"""__graft_entry__.smoke(): one small invocation of the hot path on cuda:0 — an MSM, a base fold and
a Whisk shuffle-proof round trip (ell = 4), a Pippenger MSM with and without batch-affine buckets, a batch of
proofs through the throughput kernels with and without the CRS fixed-base tables — checked against the CPU
oracle and the committed golden fixture (the oracle is the checker only)."""
import importlib
import json
import os

from oracle import bls12381 as b
from oracle.rand import Rand
from util import aff_enc, affs_dec, affs_enc, fr_enc, frs_enc, jac_dec

HERE = os.path.dirname(os.path.abspath(__file__))


def run():
    pkg = importlib.import_module("go-curdleproofs_b200")
    ctx = pkg.Context(0)
    r = Rand(1)
    pts = r.get_g1_affines(16)
    ks = r.get_frs(16)
    got = jac_dec(ctx.g1_msm(affs_enc(pts), frs_enc(ks)))
    assert got == b.g1_msm(pts, ks), "MSM mismatch vs oracle"
    x = r.get_fr()
    got = affs_dec(ctx.g1_fold(affs_enc(pts[:8]), affs_enc(pts[8:]), fr_enc(x)))
    assert got == [b.g1_add(l, b.g1_mul(q, x)) for l, q in zip(pts[:8], pts[8:])], "fold mismatch vs oracle"
    # whisk/whisk_test.go:36-56 at ell = 4 against the golden fixture generated from the oracle
    with open(os.path.join(HERE, "golden", "whisk_ell4.json")) as fh:
        g = json.load(fh)
    ell = 4
    rand = pkg.Rand(0)
    crs = ctx.generate_crs(ell, rand)
    kk, rr = [], []
    for _ in range(ell):
        kk.append(rand.get_fr())
        rr.append(rand.get_fr())
    rG = ctx.g1_scalar_mul_affine(aff_enc(b.G1_GEN) * ell, b"".join(rr), broadcast=False)
    krG = ctx.g1_scalar_mul_affine(rG, b"".join(kk), broadcast=False)
    e1, e2 = ctx.g1_compress(rG), ctx.g1_compress(krG)
    pre = b"".join(e1[48 * i:48 * i + 48] + e2[48 * i:48 * i + 48] for i in range(ell))
    post, proof = ctx.whisk_generate_shuffle_proof(crs, pre, rand)
    assert post.hex() == g["post_trackers"] and proof.hex() == g["proof"], "Whisk proof bytes differ from the fixture"
    assert ctx.whisk_is_valid_shuffle_proof(crs, pre, post, proof, rand) is True
    # a 2 000-term MSM through the Pippenger chain: (sum a_i s_i) * G
    n = 2000
    a, s = Rand(5).get_frs(n), Rand(6).get_frs(n)
    big = ctx.g1_scalar_mul_affine(aff_enc(b.G1_GEN) * n, frs_enc(a), broadcast=False)
    want = b.g1_mul(b.G1_GEN, sum(u * v for u, v in zip(a, s)) % b.R)
    assert jac_dec(ctx.g1_msm(big, frs_enc(s))) == want, "large MSM mismatch vs oracle"
    # the same MSM with batch-affine bucket accumulation forced (narrow windows, three pair-sum rounds)
    ctx.set_msm_window(7)
    ctx.set_msm_batch_affine(3)
    try:
        assert jac_dec(ctx.g1_msm(big, frs_enc(s))) == want, "batch-affine MSM mismatch vs oracle"
    finally:
        ctx.set_msm_window(0)
        ctx.set_msm_batch_affine(-1)
    # a batch of Whisk proofs through the throughput kernels, with and without the CRS fixed-base tables:
    # both runs give the same bytes, every proof validates
    B = 128
    outs = []
    for min_b in (1, 0):
        ctx.set_fixed_base_min_batch(min_b)
        rb = [pkg.Rand(200 + i) for i in range(B)]
        outs.append(ctx.whisk_generate_shuffle_proof_batch(crs, pre * B, rb))
    ctx.set_fixed_base_min_batch(1)
    assert bytes(outs[0][0]) == bytes(outs[1][0]) and bytes(outs[0][1]) == bytes(outs[1][1]), "fixed-base tables changed bytes"
    assert outs[0][2] == [0] * B
    ok, st = ctx.whisk_is_valid_shuffle_proof_batch(crs, pre * B, outs[0][0], outs[0][1], [pkg.Rand(300 + i) for i in range(B)])
    ctx.set_fixed_base_min_batch(-1)
    assert ok == [1] * B and st == [0] * B, "batched validation failed"
    print("smoke ok:", ctx.device_info())
    ctx.close()
