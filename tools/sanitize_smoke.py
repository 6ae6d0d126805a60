#!/usr/bin/env python
"""Small end-to-end workload for `compute-sanitizer --tool memcheck`: a 3 000-term large MSM (every
kernel of the Pippenger chain, including a degenerate all-equal-scalar input for the slice path) and
a 100-instance Whisk round trip on ell = 12 (throughput MSM chain with the bucket scratch, lanes)."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("go-curdleproofs_b200")
enc = importlib.import_module("go-curdleproofs_b200.encoding")
aff_enc, fr_enc = enc.aff_enc, enc.fr_enc


class b:  # generator only
    G1_GEN = enc.G1_GEN


ctx = pkg.Context(0)
n = 3000
r = pkg.Rand(5)
pts = ctx.g1_scalar_mul_affine(aff_enc(b.G1_GEN) * n, r.get_frs(n), broadcast=False)
out1 = ctx.g1_msm(pts, pkg.Rand(6).get_frs(n))
out2 = ctx.g1_msm(pts, fr_enc(123456789) * n)
ell, B = 12, 100
crs = ctx.generate_crs(ell, pkg.Rand(0))
rr = pkg.Rand(1000)
rG = ctx.g1_scalar_mul_affine(aff_enc(b.G1_GEN) * ell, rr.get_frs(ell), broadcast=False)
krG = ctx.g1_scalar_mul_affine(rG, rr.get_frs(ell), broadcast=False)
e1, e2 = ctx.g1_compress(rG), ctx.g1_compress(krG)
pre = b"".join(e1[48 * j:48 * j + 48] + e2[48 * j:48 * j + 48] for j in range(ell)) * B
post, proofs, status = ctx.whisk_generate_shuffle_proof_batch(crs, pre, [pkg.Rand(3000 + i) for i in range(B)])
ok, st = ctx.whisk_is_valid_shuffle_proof_batch(crs, pre, post, proofs, [pkg.Rand(2000 + i) for i in range(B)])
assert status == [0] * B and ok == [1] * B and st == [0] * B
print("sanitize smoke ok")
# one proof at a time: the quad (warp-cooperative) kernels, the small-MSM CTA kernel and the few-task Horner
post1, proof1 = ctx.whisk_generate_shuffle_proof(crs, pre[:ell * 96], pkg.Rand(7))
assert ctx.whisk_is_valid_shuffle_proof(crs, pre[:ell * 96], post1, proof1, pkg.Rand(8)) is True
# device-resident pool MSM and the persistent elementwise kernel above its quad threshold
pool = ctx.dev_buffer(96 * (n + 2))
pool.upload(pts + bytes(192))
aff, enc48 = ctx.g1_msm_batch_device(pool, list(range(64)) * 2, pkg.Rand(9).get_frs(128), [0, 64, 128], out_slot=[n, n + 1])
big = 9000
pts2 = ctx.g1_scalar_mul_affine((pts * 3)[:96 * big], pkg.Rand(10).get_frs(big), broadcast=False)
assert len(pts2) == 96 * big and len(aff) == 192 and len(enc48) == 96
print("sanitize smoke 2 ok")
