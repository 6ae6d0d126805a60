"""Shapes of the reference's own tests run on the pure-Python
oracle at a small size: completeness, the four soundness mutations, encode /
decode stability, accumulator-level soundness."""
import pytest

from oracle import bls12381 as bls
from oracle import protocol as P
from oracle import whisk as W
from oracle.rand import Rand

ELL = 4  # n = 8


@pytest.fixture(scope="module")
def inst():
    rand = Rand(0)
    crs = P.generate_crs(ELL, rand)
    k = rand.get_fr()
    Rs = rand.get_g1_affines(ELL)
    Ss = rand.get_g1_affines(ELL)
    perm = Rand(42).generate_permutation(ELL)
    Ts, Us, M, rs_m = P.shuffle_permute_commit(crs.Gs, crs.Hs, Rs, Ss, perm, k, rand)
    proof = P.prove(crs, Rs, Ss, Ts, Us, M, perm, k, rs_m, Rand(42))
    return dict(crs=crs, k=k, Rs=Rs, Ss=Ss, Ts=Ts, Us=Us, M=M, perm=perm, rs_m=rs_m, proof=proof)


def test_completeness(inst):  # curdleproof_test.go:16-46
    i = inst
    assert P.verify(i["proof"], i["crs"], i["Rs"], i["Ss"], i["Ts"], i["Us"], i["M"], Rand(43)) is True


def test_soundness_mutations(inst):  # curdleproof_test.go:48-167
    i = inst
    crs, proof = i["crs"], i["proof"]
    assert P.verify(proof, crs, i["Ss"], i["Rs"], i["Ts"], i["Us"], i["M"], Rand(43)) is False
    p2 = Rand(5).generate_permutation(ELL)
    assert P.verify(proof, crs, i["Rs"], i["Ss"], P.permute(i["Ts"], p2), P.permute(i["Us"], p2), i["M"], Rand(43)) is False
    assert P.verify(proof, crs, i["Rs"], i["Ss"], i["Ts"], i["Us"], bls.g1_mul(i["M"], i["k"]), Rand(43)) is False
    k2 = Rand(9).get_fr()
    Ts2 = [bls.g1_mul(t, k2) for t in i["Ts"]]
    Us2 = [bls.g1_mul(u, k2) for u in i["Us"]]
    assert P.verify(proof, crs, i["Rs"], i["Ss"], Ts2, Us2, i["M"], Rand(43)) is False


def test_encode_decode_stable(inst):  # curdleproof_test.go:169-181
    b = inst["proof"].serialize()
    assert len(b) == 48 * (18 + 10 * 3) + 32 * 7 + 4 * 10  # 18+10m points, 7 scalars, 10 length prefixes
    assert P.Proof.deserialize(b).serialize() == b


def test_randomizer_zero_is_error(inst):
    i = inst
    with pytest.raises(P.ProofError):
        P.verify(i["proof"], i["crs"], i["Rs"], i["Ss"], [None] + i["Ts"][1:], i["Us"], i["M"], Rand(43))


def test_whisk_roundtrip_small():  # whisk_test.go:36-56 at ell = 4
    rand = Rand(0)
    crs = P.generate_crs(ELL, rand)
    pre = W.generate_shuffle_trackers(rand, ELL)
    post, pb = W.generate_whisk_shuffle_proof(crs, pre, rand, ell=ELL)
    assert len(pb) == W.WHISK_SHUFFLE_PROOF_SIZE
    assert W.is_valid_whisk_shuffle_proof(crs, pre, post, pb, rand) is True
    bad = bytearray(pb)
    bad[-100] ^= 1  # inside the scalar x at the end of the used region? (padding stays ignored)
    # flipping a padding byte must not matter; flipping a used byte must reject or error
    assert W.is_valid_whisk_shuffle_proof(crs, pre, post, bytes(pb[:-1]) + b"\x01", Rand(1)) is True
    used = len(pb.rstrip(b"\0"))
    bad = bytearray(pb)
    bad[used - 1] ^= 1
    try:
        ok = W.is_valid_whisk_shuffle_proof(crs, pre, post, bytes(bad), Rand(1))
    except (W.WhiskError, P.ProofError):
        ok = False
    assert ok is False


def test_tracker_proof_roundtrip():  # whisk_test.go:13-34
    rand = Rand(0)
    k = rand.get_fr()
    r = rand.get_fr()
    rG = bls.g1_mul(bls.G1_GEN, r)
    tracker = W.new_tracker(rG, bls.g1_mul(rG, k))
    k_comm = bls.g1_compress(bls.g1_mul(bls.G1_GEN, k))
    pr = W.generate_whisk_tracker_proof(tracker, k, rand)
    assert len(pr) == W.TRACKER_PROOF_SIZE
    assert W.is_valid_whisk_tracker_proof(tracker, k_comm, pr) is True
    assert W.is_valid_whisk_tracker_proof(tracker, bls.g1_compress(bls.G1_GEN), pr) is False


def test_msm_accumulator():  # msmaccumulator_test.go:12-50
    rand = Rand(0)
    acc = P.MsmAccumulator()
    for n in (1, 2, 3):
        pts = rand.get_g1_affines(n)
        xs = rand.get_frs(n)
        acc.accumulate_check(bls.g1_msm(pts, xs), xs, pts, rand)
    assert acc.verify() is True
    pts = rand.get_g1_affines(2)
    xs = rand.get_frs(2)
    acc.accumulate_check(bls.g1_msm(pts, xs), [xs[0], xs[1] + 1], pts, rand)
    assert acc.verify() is False


def test_lazy_folding_identities():
    """The algebra behind the prover's lazy schedule (csrc/engine.cu, `lazy`): two folds of a base
    vector (innerproductargument.go:155-166, samemultiscalarargument.go:129-135: V <- V_L + x V_R) are the linear
    combination V2[i] = V[i] + x1 V[i + n/2] + x2 V[i + n/4] + x1 x2 V[i + 3n/4]; the second round's MSM over the
    once-folded bases equals an MSM with two terms per base on the unfolded ones; and G' = scale * C folds the same
    way with the scale factors carried in the scalars."""
    rand = Rand(11)
    n, h, q = 16, 8, 4
    C = rand.get_g1_affines(n)
    x1, x2 = rand.get_fr(), rand.get_fr()
    scale = rand.get_frs(n)
    a = rand.get_frs(h)

    def fold(V, x):
        m = len(V) // 2
        return [bls.g1_add(V[i], bls.g1_mul(V[m + i], x)) for i in range(m)]

    R = bls.R
    V1 = fold(C, x1)
    V2 = fold(V1, x2)
    for i in range(q):
        want = bls.g1_msm_naive([C[i], C[i + h], C[i + q], C[i + 3 * q]], [1, x1, x2, x1 * x2 % R])
        assert bls.g1_eq(V2[i], want)
    # second-round MSM <a, V1> on pairs of unfolded bases
    pts, scs = [], []
    for e in range(h):
        pts += [C[e], C[e + h]]
        scs += [a[e], a[e] * x1 % R]
    assert bls.g1_eq(bls.g1_msm_naive(V1, a), bls.g1_msm_naive(pts, scs))
    # G'[i] = scale[i] C[i]: folds with the inverse challenges, scale factors multiplied into the scalars
    y1, y2 = bls.fr_inv(x1), bls.fr_inv(x2)
    Gp2 = fold(fold([bls.g1_mul(C[i], scale[i]) for i in range(n)], y1), y2)
    for i in range(q):
        want = bls.g1_msm_naive([C[i], C[i + h], C[i + q], C[i + 3 * q]],
                                [scale[i], y1 * scale[i + h] % R, y2 * scale[i + q] % R, y1 * y2 % R * scale[i + 3 * q] % R])
        assert bls.g1_eq(Gp2[i], want)
