"""__graft_entry__.smoke(): one small hot-path invocation on cuda:0, checked
against the CPU oracle (checker only)."""
import importlib

from oracle import bls12381 as b
from oracle.rand import Rand
from util import affs_dec, affs_enc, fr_enc, frs_enc, jac_dec


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
    print("smoke ok:", ctx.device_info())
    ctx.close()
