"""GPU parity tests of the large-MSM (Pippenger) path through the C ABI
(BASELINE.json config 5).  Oracle: for points P_i = a_i*G the exact result is
(sum a_i*s_i mod r)*G — O(N) Fr work on the CPU oracle for any N (SURVEY.md
§8d) — plus direct comparison with the oracle's own MSM at small sizes."""
import random

import pytest

from oracle import bls12381 as b
from oracle.rand import Rand
from util import R, aff_enc, affs_dec, affs_enc, fr_enc, frs_enc, jac_dec

pytestmark = pytest.mark.gpu


def make_points(ctx, a):
    """P_i = a_i * G on the GPU (cdl_g1_scalar_mul_affine is oracle-checked in test_gpu_primitives)."""
    return ctx.g1_scalar_mul_affine(aff_enc(b.G1_GEN) * len(a), frs_enc(a), broadcast=False)


def expected(a, s):
    return b.g1_mul(b.G1_GEN, sum(x * y for x, y in zip(a, s)) % R)


@pytest.fixture(scope="module")
def sweep(ctx):
    n = 1 << 16
    a = Rand(5).get_frs(n)
    s = Rand(6).get_frs(n)
    return a, s, make_points(ctx, a)


@pytest.mark.parametrize("n", [1025, 1500, 2048, 4099, 1 << 14, 1 << 16])
def test_big_msm_sizes(ctx, sweep, n):
    a, s, pts = sweep
    got = jac_dec(ctx.g1_msm(pts[: 96 * n], frs_enc(s[:n])))
    assert got == expected(a[:n], s[:n])


def test_big_msm_matches_oracle_msm(ctx):
    r = Rand(31)
    base = r.get_g1_affines(40)
    n = 1300
    ps = [base[i % 40] for i in range(n)]
    ks = r.get_frs(n)
    from oracle.cbackend import CBackend
    want = CBackend(accelerate_keccak=False).msm(ps, ks)
    out = ctx.g1_msm(affs_enc(ps), frs_enc(ks))
    assert jac_dec(out) == want
    # normalised representative: Z == 1 in Montgomery form
    assert out[96:144] == (pow(2, 384, b.P)).to_bytes(48, "little")


@pytest.mark.parametrize("c", [2, 3, 5, 7, 8, 9, 11, 13, 15, 16, 17])
def test_big_msm_every_window_width(ctx, sweep, c):
    a, s, pts = sweep
    n = 3000
    ctx.set_msm_window(c)
    try:
        got = jac_dec(ctx.g1_msm(pts[: 96 * n], frs_enc(s[:n])))
    finally:
        ctx.set_msm_window(0)
    assert got == expected(a[:n], s[:n])


@pytest.mark.parametrize("c,rounds", [(8, 1), (8, 2), (8, 3), (8, 5), (8, 8), (11, 4), (5, 6), (13, 1), (16, 2)])
def test_big_msm_batch_affine_rounds(ctx, sweep, c, rounds):
    """Batch-affine bucket accumulation (cdl_set_msm_batch_affine): any number of pair-sum rounds, more
    than the buckets hold included, gives the same point as the extended-Jacobian buckets alone."""
    a, s, pts = sweep
    n = 1 << 14
    ctx.set_msm_window(c)
    try:
        ctx.set_msm_batch_affine(0)
        plain = ctx.g1_msm(pts[: 96 * n], frs_enc(s[:n]))
        ctx.set_msm_batch_affine(rounds)
        got = ctx.g1_msm(pts[: 96 * n], frs_enc(s[:n]))
        odd = jac_dec(ctx.g1_msm(pts[: 96 * 4099], frs_enc(s[:4099])))
    finally:
        ctx.set_msm_window(0)
        ctx.set_msm_batch_affine(-1)
    assert got == plain and jac_dec(got) == expected(a[:n], s[:n])
    assert odd == expected(a[:4099], s[:4099])


@pytest.mark.parametrize("c,rounds", [(0, -1), (9, 3), (12, 2), (7, 7)])
def test_big_msm_adversarial(ctx, sweep, c, rounds):
    ctx.set_msm_window(c)
    ctx.set_msm_batch_affine(rounds)
    try:
        _adversarial(ctx, sweep)
    finally:
        ctx.set_msm_window(0)
        ctx.set_msm_batch_affine(-1)


def _adversarial(ctx, sweep):
    a, s, pts = sweep
    n = 1 << 14
    random.seed(9)
    A = a[:n]
    P = pts[: 96 * n]
    # all-equal scalars: every point of a window lands in one bucket (large-bucket slice path)
    k = random.randrange(R)
    assert jac_dec(ctx.g1_msm(P, fr_enc(k) * n)) == expected(A, [k] * n)
    # scalars below 2^9 (common/util.go:68-75 shape), zero scalars among them
    tiny = [random.randrange(512) for _ in range(n)]
    assert jac_dec(ctx.g1_msm(P, frs_enc(tiny))) == expected(A, tiny)
    # extreme scalars
    ext = [random.choice([0, 1, R - 1, R - 2, 2**254, (R - 1) // 2, (R + 1) // 2, 2**255 % R]) for _ in range(n)]
    assert jac_dec(ctx.g1_msm(P, frs_enc(ext))) == expected(A, ext)
    # 1 % infinity bases
    S = s[:n]
    Pb = bytearray(P)
    A2 = list(A)
    for i in random.sample(range(n), n // 100):
        Pb[96 * i: 96 * i + 96] = bytes(96)
        A2[i] = 0
    assert jac_dec(ctx.g1_msm(bytes(Pb), frs_enc(S))) == expected(A2, S)
    # duplicated points with equal scalars (P + P inside a bucket) and +-P pairs (cancel)
    half = n // 2
    dup = P[: 96 * half] + P[: 96 * half]
    sd = S[:half] + S[:half]
    assert jac_dec(ctx.g1_msm(dup, frs_enc(sd))) == expected(A[:half] + A[:half], sd)
    neg = affs_enc([b.g1_neg(p) for p in affs_dec(P[: 96 * 600])])
    pm = P[: 96 * 600] + neg + P[96 * 600: 96 * 1200]
    spm = S[:600] + S[:600] + S[600:1200]
    assert jac_dec(ctx.g1_msm(pm, frs_enc(spm))) == expected(A[600:1200], S[600:1200])
    # everything cancels: infinity, gnark's (1, 1, 0)
    out = ctx.g1_msm(P[: 96 * 600] + neg + P[: 96 * 600] + neg, frs_enc(S[:600] * 4))
    assert jac_dec(out) is None and out[96:144] == bytes(48)
    # all scalars zero
    assert jac_dec(ctx.g1_msm(P, bytes(32 * n))) is None


@pytest.mark.parametrize("parts", [2, 3, 8, 40])
def test_window_partition_partials_add_up(ctx, sweep, parts):
    """Multi-GPU decomposition on one GPU: the shifted partial sums of all window
    parts add up to the MSM (what cdl_g1_msm_sharded all-gathers)."""
    a, s, pts = sweep
    n = 5000
    dp = ctx.dev_buffer(96 * n)
    ds = ctx.dev_buffer(32 * n)
    dp.upload(pts[: 96 * n])
    ds.upload(frs_enc(s[:n]))
    full, ms = ctx.g1_msm_device(dp, ds, n)
    assert jac_dec(full) == expected(a[:n], s[:n]) and ms > 0
    acc = None
    for p in range(parts):
        part, _ = ctx.g1_msm_device(dp, ds, n, part_index=p, part_count=parts, normalize=False)
        acc = b.g1_add(acc, jac_dec(part))
    assert acc == jac_dec(full)
    # world == 1 sharded form needs no communicator
    one, _ = ctx.g1_msm_sharded_device(dp, ds, n)
    assert one == full
    dp.close()
    ds.close()


def test_device_scalar_mul_and_sum(ctx, sweep):
    a, s, pts = sweep
    n = 2000
    dp = ctx.dev_buffer(96 * n)
    ds = ctx.dev_buffer(32 * n)
    dp.upload(aff_enc(b.G1_GEN) * n)
    ds.upload(frs_enc(a[:n]))
    ctx.g1_scalar_mul_affine_device(dp, ds, n, False, dp)
    assert dp.download() == pts[: 96 * n]
    # device-resident fold L[i] += x * R[i]: (a_i + x * a_{i+1000}) * G
    x = 0x1234567890abcdef1234567890abcdef % R
    dl = ctx.dev_buffer(96 * 1000)
    dr = ctx.dev_buffer(96 * 1000)
    dx = ctx.dev_buffer(32)
    dl.upload(pts[: 96 * 1000])
    dr.upload(pts[96 * 1000: 96 * 2000])
    dx.upload(fr_enc(x))
    ctx.g1_fold_device(dl, dr, dx, 1000)
    want = ctx.g1_scalar_mul_affine(aff_enc(b.G1_GEN) * 1000, frs_enc([(a[i] + x * a[i + 1000]) % R for i in range(1000)]),
                                    broadcast=False)
    assert dl.download() == want
    for d in (dl, dr, dx):
        d.close()
    # Gsum-style sum of many points (crs.go:41-48) goes through the same path
    from util import aff_dec
    assert aff_dec(ctx.g1_sum_affine(pts[: 96 * n])) == b.g1_mul(b.G1_GEN, sum(a[:n]) % R)
    dp.close()
    ds.close()


def test_big_msm_2_20_property(ctx, sweep):
    """Full-size check by linearity: MSM(P, s) for N = 2^20 built from 16 rotations of the
    2^16 base vector equals (sum a_i * s_i) * G."""
    a, s, pts = sweep
    reps = 16
    n = len(a) * reps
    r = Rand(77)
    mults = r.get_frs(reps)
    sc = []
    for m in mults:
        sc.extend(x * m % R for x in s)
    tot = sum(x * y for x, y in zip(a, s)) % R
    want = b.g1_mul(b.G1_GEN, tot * sum(mults) % R)
    for rounds in (-1, 0, 3):  # chosen by bucket load / extended-Jacobian buckets only / three batch-affine rounds
        ctx.set_msm_batch_affine(rounds)
        try:
            got = jac_dec(ctx.g1_msm(pts * reps, frs_enc(sc)))
        finally:
            ctx.set_msm_batch_affine(-1)
        assert got == want
