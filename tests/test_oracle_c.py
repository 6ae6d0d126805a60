"""The two CPU restatements (pure-Python big-int, plain C 6x64 Montgomery) must
agree; the C one then serves as the fast checker / CPU baseline."""
import hashlib
import random

import pytest

from oracle import bls12381 as b
from oracle import merlin
from oracle import protocol as P
from oracle import whisk as W
from oracle.cbackend import CBackend
from oracle.rand import Rand


@pytest.fixture(scope="module")
def cb():
    return CBackend(accelerate_keccak=False)


def test_c_mul_fold_msm_match_python(cb):
    r = Rand(5)
    pts = r.get_g1_affines(24) + [None]
    ks = r.get_frs(22) + [0, 1, b.R - 1]
    assert cb.mul_batch(pts, ks) == b.g1_mul_batch(pts, ks)
    x = r.get_fr()
    L, Rr = pts[:12], pts[12:24]
    assert cb.fold(L, Rr, x) == [b.g1_add(l, b.g1_mul(q, x)) for l, q in zip(L, Rr)]
    for n in (1, 2, 7, 9, 25):
        assert cb.msm(pts[:n], ks[:n]) == b.g1_msm_naive(pts[:n], ks[:n])
    assert cb.msm([], []) is None
    assert cb.sum(pts) == b.g1_sum(pts)
    # larger, threaded: algebraic check  sum s_i (a_i G) = (sum a_i s_i) G
    random.seed(1)
    n = 600
    a = [random.randrange(b.R) for _ in range(n)]
    s = [random.randrange(b.R) for _ in range(n)]
    bases = cb.mul_batch([b.G1_GEN] * n, a)
    want = b.g1_mul(b.G1_GEN, sum(x * y for x, y in zip(a, s)) % b.R)
    assert cb.msm(bases, s) == want


def test_c_codec_matches_python(cb):
    r = Rand(6)
    pts = r.get_g1_affines(10) + [None]
    enc = cb.compress(pts)
    assert enc == b"".join(b.g1_compress(p) for p in pts)
    dec, st = cb.decompress(enc)
    assert dec == pts and not any(st)
    dec, st = cb.decompress(bytes([0x9F]) + b"\xff" * 47 + bytes([0xC0]) + bytes(46) + b"\x01")
    assert st == [2, 5]


def test_c_keccak_matches_python(cb):
    for seed in range(4):
        st = bytearray(hashlib.shake_256(bytes([seed])).digest(200))
        a = bytearray(st)
        bb = bytearray(st)
        merlin.keccak_f1600(a)
        cb.keccak(bb)
        assert a == bb


def test_whisk_proof_bytes_identical_across_backends(cb):
    ell = 4

    def run():
        rand = Rand(0)
        crs = P.generate_crs(ell, rand)
        pre = W.generate_shuffle_trackers(rand, ell)
        post, pb = W.generate_whisk_shuffle_proof(crs, pre, rand, ell=ell)
        ok = W.is_valid_whisk_shuffle_proof(crs, pre, post, pb, rand)
        return post, pb, ok

    P.set_backend(P.PyBackend())
    ref = run()
    try:
        P.set_backend(cb)
        got = run()
    finally:
        P.set_backend(P.PyBackend())
    assert got == ref and ref[2] is True
