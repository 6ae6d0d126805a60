"""CPU validation of the product's field / curve source (csrc/*.cuh) against
the oracle.  The carry-chain primitives are emulated bit-exactly on the host
(csrc/carry.cuh), so this runs the same Montgomery / group-law code the GPU
runs, without a GPU.  The harness is test-only."""
import ctypes
import os
import random
import subprocess

import pytest

from oracle import bls12381 as b
from util import RP, RP_INV, RR_INV, aff_dec, aff_enc

P, R = b.P, b.R
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hc(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hc") / "libhostcheck.so")
    src = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
    # CDL_HOSTCHECK_FLAGS="-fsanitize=undefined -fno-sanitize-recover=all -static-libubsan" runs the tier under UBSan
    extra = os.environ.get("CDL_HOSTCHECK_FLAGS", "").split()
    subprocess.run(["g++", "-O2", "-std=c++17", *extra, "-x", "c++", "-shared", "-fPIC", "-o", out, src], check=True)
    return ctypes.CDLL(out)


def pack(vals, nb):
    return b"".join(v.to_bytes(nb, "little") for v in vals)


def unpack(buf, nb):
    return [int.from_bytes(buf[i:i + nb], "little") for i in range(0, len(buf), nb)]


def run2(fn, A, B, nb=48):
    out = ctypes.create_string_buffer(nb * len(A))
    fn(pack(A, nb), pack(B, nb), out, len(A))
    return unpack(out.raw, nb)


def run1(fn, A, nb=48):
    out = ctypes.create_string_buffer(nb * len(A))
    fn(pack(A, nb), out, len(A))
    return unpack(out.raw, nb)


def tom(x):
    return x * RP % P


def test_fp_arithmetic(hc):
    random.seed(1)
    edge = [0, 1, P - 1, P - 2, RP, RP * RP % P, 2**380, P - 3, (P - 1) // 2, (P + 1) // 2]
    vals = edge + [random.randrange(P) for _ in range(300)]
    A = [random.choice(vals) for _ in range(1500)]
    B = [random.choice(vals) for _ in range(1500)]
    assert run2(hc.hc_fp_mul, [tom(a) for a in A], [tom(x) for x in B]) == [tom(a * x % P) for a, x in zip(A, B)]
    assert run2(hc.hc_fp_mul, A, B) == [a * x * RP_INV % P for a, x in zip(A, B)]
    assert run2(hc.hc_fp_add, A, B) == [(a + x) % P for a, x in zip(A, B)]
    assert run2(hc.hc_fp_sub, A, B) == [(a - x) % P for a, x in zip(A, B)]
    assert run1(hc.hc_fp_neg, A) == [(-a) % P for a in A]
    assert run1(hc.hc_fp_from_mont, A) == [a * RP_INV % P for a in A]
    assert run1(hc.hc_fp_to_mont, A) == [tom(a) for a in A]
    S = edge + [1 << k for k in (1, 31, 32, 63, 64, 380)] + A[:200]  # inversion: every edge value incl. 0
    assert run1(hc.hc_fp_inv, [tom(a) for a in S]) == [tom(pow(a, P - 2, P)) for a in S]
    # the approximate-GCD inversion against the limb-wide binary Euclid it replaced, on raw residues of every size
    T = S + [random.randrange(1 << k) for k in (8, 31, 32, 33, 62, 63, 64, 65, 96, 127, 190, 255, 320, 379) for _ in range(40)]
    T = [t % P for t in T]
    assert run1(hc.hc_fp_inv, T) == run1(hc.hc_fp_inv_eea, T) == [pow(t * RP_INV % P, P - 2, P) * RP % P for t in T]
    ok = (ctypes.c_int * len(S))()
    out = ctypes.create_string_buffer(48 * len(S))
    hc.hc_fp_sqrt(pack([tom(a * a % P) for a in S], 48), out, ok, len(S))
    for a, s, o in zip(S, unpack(out.raw, 48), ok):
        assert o == 1 and s * RP_INV % P in (a % P, (P - a) % P)
    hc.hc_fp_sqrt(pack([tom(a) for a in S], 48), out, ok, len(S))
    assert list(ok) == [int(b.fp_sqrt(a) is not None) for a in S]
    lx = (ctypes.c_int * len(A))()
    hc.hc_fp_lex(pack([tom(a) for a in A], 48), lx, len(A))
    assert list(lx) == [int(a > (P - 1) // 2) for a in A]


def test_fp_fr_dedicated_squaring(hc):
    """Mont::sqr (off-diagonal + diagonal products, then a pure reduction) against a*a on values that
    stress every carry path: all-ones limbs, single-limb values, values next to p and to 2^381."""
    random.seed(11)
    ones = [(1 << (32 * k)) - 1 for k in range(1, 13)]
    single = [0xffffffff << (32 * k) for k in range(12)]
    alt = [int("ffffffff00000000" * 6, 16) % P, int("00000000ffffffff" * 6, 16), int("80000000" * 12, 16) % P,
           int("7fffffff" * 12, 16) % P, int("ffffffff" * 12, 16) % P]
    edge = [0, 1, 2, P - 1, P - 2, (P - 1) // 2, 2**380, 2**381 - 1 - (2**381 - 1 >= P) * (2**381 - P)]
    vals = [v % P for v in ones + single + alt + edge] + [random.randrange(P) for _ in range(3000)]
    assert run1(hc.hc_fp_sqr, vals) == [a * a * RP_INV % P for a in vals]
    assert run1(hc.hc_fp_sqr, vals) == run2(hc.hc_fp_mul, vals, vals)
    rones = [((1 << (32 * k)) - 1) % R for k in range(1, 9)]
    rvals = rones + [0, 1, R - 1, R - 2, (R - 1) // 2] + [random.randrange(R) for _ in range(3000)]
    assert run1(hc.hc_fr_sqr, rvals, 32) == [a * a * RR_INV % R for a in rvals]


def test_fr_arithmetic(hc):
    random.seed(2)
    FA = [random.randrange(R) for _ in range(300)] + [0, 1, R - 1]
    FB = [random.randrange(R) for _ in range(303)]
    assert run2(hc.hc_fr_mul, FA, FB, 32) == [a * x * RR_INV % R for a, x in zip(FA, FB)]
    assert run1(hc.hc_fr_from_mont, FA, 32) == [a * RR_INV % R for a in FA]


def _binop(hc, op, As, Bs):
    out = ctypes.create_string_buffer(96 * len(As))
    hc.hc_g1_binop(op, b"".join(aff_enc(p) for p in As), b"".join(aff_enc(p) for p in Bs), out, len(As))
    return [aff_dec(out.raw[i * 96:(i + 1) * 96]) for i in range(len(As))]


def test_group_law_all_cases(hc):
    random.seed(3)
    pts = [b.g1_mul(b.G1_GEN, random.randrange(R)) for _ in range(24)]
    As = pts[:12] + [None, pts[0], pts[1], None, pts[2]]
    Bs = pts[12:] + [pts[3], None, b.g1_neg(pts[1]), None, pts[2]]
    d = lambda p: b.g1_add(p, p)  # noqa: E731
    exp = {0: lambda a, c: b.g1_add(d(a), c), 1: lambda a, c: b.g1_add(d(a), d(c)),
           2: lambda a, c: b.g1_add(d(a), c), 3: lambda a, c: b.g1_add(d(a), d(c)),
           4: lambda a, c: d(a), 5: lambda a, c: d(a),
           6: b.g1_add, 7: b.g1_add, 8: b.g1_add, 9: b.g1_add}
    for op in range(10):
        assert _binop(hc, op, As, Bs) == [exp[op](a, c) for a, c in zip(As, Bs)], op
    a = pts[5]
    assert _binop(hc, 0, [a, a], [d(a), b.g1_neg(d(a))]) == [d(d(a)), None]
    assert _binop(hc, 2, [a, a], [d(a), b.g1_neg(d(a))]) == [d(d(a)), None]
    assert _binop(hc, 1, [a, a], [a, b.g1_neg(a)]) == [d(d(a)), None]
    assert _binop(hc, 3, [a, a], [a, b.g1_neg(a)]) == [d(d(a)), None]


def test_scalar_mul(hc):
    random.seed(4)
    pts = [b.g1_mul(b.G1_GEN, random.randrange(R)) for _ in range(6)]
    ks = [0, 1, 2, 8, 9, 15, 16, R - 1, R - 2, R - 3, int("8" * 64, 16) % R] + [random.randrange(R) for _ in range(10)]
    ps = [random.choice(pts) for _ in ks]
    ps[3] = None
    out = ctypes.create_string_buffer(96 * len(ks))
    hc.hc_g1_scalar_mul(b"".join(aff_enc(p) for p in ps), pack(ks, 32), out, len(ks))
    got = [aff_dec(out.raw[i * 96:(i + 1) * 96]) for i in range(len(ks))]
    assert got == [b.g1_mul(p, k) for p, k in zip(ps, ks)]


def test_glv_decomposition_and_scalar_mul(hc):
    random.seed(7)
    lam = 0xAC45A4010001A40200000000FFFFFFFF
    ks = [0, 1, 2, lam, lam - 1, lam + 1, (R - 1) // 2, (R + 1) // 2, R - 1, R - 2, lam * lam % R] + \
         [random.randrange(R) for _ in range(300)]
    n = len(ks)
    k1 = (ctypes.c_uint32 * (4 * n))()
    k2 = (ctypes.c_uint32 * (4 * n))()
    n1 = (ctypes.c_int * n)()
    n2 = (ctypes.c_int * n)()
    hc.hc_glv_decompose(pack(ks, 32), k1, k2, n1, n2, n)
    for i, k in enumerate(ks):
        a = sum(k1[4 * i + j] << (32 * j) for j in range(4))
        c = sum(k2[4 * i + j] << (32 * j) for j in range(4))
        assert a < 2**127 and c < 2**127
        val = ((-a if n1[i] else a) + (-c if n2[i] else c) * lam) % R
        assert val == k, i
    pts = [b.g1_mul(b.G1_GEN, random.randrange(R)) for _ in range(4)]
    sel = ks[:11] + ks[11:31]
    ps = [random.choice(pts) for _ in sel]
    ps[4] = None
    out = ctypes.create_string_buffer(96 * len(sel))
    hc.hc_g1_scalar_mul_glv(b"".join(aff_enc(p) for p in ps), pack(sel, 32), out, len(sel))
    got = [aff_dec(out.raw[i * 96:(i + 1) * 96]) for i in range(len(sel))]
    assert got == [b.g1_mul(p, k) for p, k in zip(ps, sel)]


def test_endomorphism_subgroup_check(hc):
    random.seed(8)
    z = 0xD201000000010000
    h = (z + 1) ** 2 // 3
    good = [b.g1_mul(b.G1_GEN, random.randrange(1, R)) for _ in range(3)] + [None]
    bad = []
    while len(bad) < 4:
        x = random.randrange(P)
        y = b.fp_sqrt((x**3 + 4) % P)
        if y is None:
            continue
        Q = (x, y)
        bad.append(Q)                                              # generic curve point
        C = b._from_jac(b.g1_mul_jac_raw(Q, R))                    # cofactor-subgroup point
        if C is not None:
            bad.append(C)
        S = b._from_jac(b.g1_mul_jac_raw(Q, R * (h // 3)))         # order-3 point
        if S is not None:
            bad.append(S)
            bad.append(b.g1_add(good[0], S))                       # G1 point + small-order point
    pts = good + bad
    out = (ctypes.c_int * len(pts))()
    hc.hc_g1_in_subgroup(b"".join(aff_enc(p) for p in pts), out, len(pts))
    assert list(out) == [1] * len(good) + [0] * len(bad)
    assert [int(b.g1_in_subgroup(p)) for p in pts] == list(out)


def test_batch_affine_run_all_cases(hc):
    """One batch-affine run (csrc/batch_affine.cuh: shared inversion by Montgomery's trick) over pairs that
    include every exceptional case — doubling, P - P, infinity on either or both sides — in the middle of
    ordinary pairs, against the oracle's affine addition."""
    random.seed(17)
    pts = [b.g1_mul(b.G1_GEN, random.randrange(R)) for _ in range(40)]
    As = pts[:10] + [None, pts[0], pts[1], None, pts[2]] + pts[10:20] + [pts[5], None]
    Bs = pts[20:30] + [pts[3], None, b.g1_neg(pts[1]), None, pts[2]] + pts[30:40] + [b.g1_neg(pts[5]), None]
    for lo, hi in ((0, len(As)), (10, 15), (12, 13), (14, 15), (3, 4)):
        A, Bv = As[lo:hi], Bs[lo:hi]
        out = ctypes.create_string_buffer(96 * len(A))
        pre = ctypes.create_string_buffer(48 * len(A))
        hc.hc_batch_affine(b"".join(aff_enc(p) for p in A), b"".join(aff_enc(p) for p in Bv), out, pre, len(A))
        got = [aff_dec(out.raw[i * 96:(i + 1) * 96]) for i in range(len(A))]
        assert got == [b.g1_add(x, y) for x, y in zip(A, Bv)]


def test_fixed_base_table_lookup(hc):
    """k * P from a fixed-base table (csrc/fixed_base.cuh: signed 12-bit digits, 22 look-ups, no doublings)
    against the oracle, on scalars that stress the digit recoding: all-ones windows (carries ripple through
    every window), digits exactly 2048 / 2049, r - 1, single bits at window boundaries, zero."""
    random.seed(23)
    pt = b.g1_mul(b.G1_GEN, 0x1234567)
    ks = [0, 1, 2047, 2048, 2049, 4095, 4096, 4097, (1 << 12) - 1, (1 << 252) - 1, R - 1, R - 2, (R - 1) // 2,
          int("800" * 21, 16), int("801" * 21, 16), int("7ff" * 21, 16), int("fff" * 21, 16) % R, 1 << 252,
          (1 << 254) + (1 << 11), (7 << 252) | ((1 << 252) - 1) if ((7 << 252) | ((1 << 252) - 1)) < R else R - 3]
    ks += [1 << (12 * w) for w in range(22) if (1 << (12 * w)) < R] + [random.randrange(R) for _ in range(40)]
    negs = [i % 3 == 0 for i in range(len(ks))]
    out = ctypes.create_string_buffer(96 * len(ks))
    kb = b"".join(k.to_bytes(32, "little") for k in ks)
    hc.hc_fixed_base(aff_enc(pt), kb, (ctypes.c_int * len(ks))(*[int(x) for x in negs]), out, len(ks))
    got = [aff_dec(out.raw[i * 96:(i + 1) * 96]) for i in range(len(ks))]
    want = [b.g1_mul(pt, (R - k) % R if ng else k) for k, ng in zip(ks, negs)]
    assert got == want
