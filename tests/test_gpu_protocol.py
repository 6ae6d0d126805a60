"""GPU parity of the protocol-level API against the golden fixtures (byte
identical proofs, identical verdicts) and the reference's own test shapes
(SURVEY.md §4): completeness, the four soundness mutations, encode/decode,
Whisk round trip, malformed inputs."""
import hashlib
import json
import os

import pytest

from oracle import bls12381 as b
from util import aff_dec, aff_enc, affs_dec, affs_enc, fr_dec, fr_enc, jac_dec, jac_enc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    with open(os.path.join(GOLD, name)) as fh:
        return json.load(fh)


def crs_enc(ctx, crs):
    return ctx.g1_compress(crs.export())


@pytest.mark.parametrize("ell", [4, 12, 124])
def test_whisk_roundtrip_matches_golden(ctx, pkg, ell):
    """whisk/whisk_test.go:36-56 with one continuing Rand(0): CRS, trackers,
    GenerateWhiskShuffleProof, IsValidWhiskShuffleProof."""
    g = gold(f"whisk_ell{ell}.json")
    rand = pkg.Rand(0)
    crs = ctx.generate_crs(ell, rand)
    assert crs_enc(ctx, crs).hex() == g["crs"]
    # generateShuffleTrackers: per tracker k then r; (r*G, k*r*G)
    ks, rs = [], []
    for _ in range(ell):
        ks.append(rand.get_fr())
        rs.append(rand.get_fr())
    gen = aff_enc(b.G1_GEN)
    rG = ctx.g1_scalar_mul_affine(gen * ell, b"".join(rs), broadcast=False)
    krG = ctx.g1_scalar_mul_affine(rG, b"".join(ks), broadcast=False)
    e1, e2 = ctx.g1_compress(rG), ctx.g1_compress(krG)
    pre = b"".join(e1[48 * i:48 * i + 48] + e2[48 * i:48 * i + 48] for i in range(ell))
    assert pre.hex() == g["pre_trackers"]
    post, proof = ctx.whisk_generate_shuffle_proof(crs, pre, rand)
    assert post.hex() == g["post_trackers"]
    assert proof.hex() == g["proof"]
    assert ctx.whisk_is_valid_shuffle_proof(crs, pre, post, proof, rand) is g["valid"] is True
    assert "%064x" % fr_dec(rand.get_fr()) == g["next_fr_after_roundtrip"]


def _setup(ctx, pkg, ell):
    """curdleproof_test.go:239-274 with perm = Rand(42).GeneratePermutation."""
    rand = pkg.Rand(0)
    crs = ctx.generate_crs(ell, rand)
    perm = pkg.Rand(42).generate_permutation(ell)
    k = rand.get_fr()
    Rs = ctx.rand_get_g1_affines(rand, ell)
    Ss = ctx.rand_get_g1_affines(rand, ell)
    Ts, Us, M, rs_m = ctx.shuffle_permute_commit(crs, Rs, Ss, perm, k, rand)
    return crs, Rs, Ss, Ts, Us, M, perm, k, rs_m


@pytest.mark.parametrize("ell", [12, 60, 508])
def test_prove_verify_matches_golden(ctx, pkg, ell):
    g = gold(f"prove_ell{ell}.json")
    crs, Rs, Ss, Ts, Us, M, perm, k, rs_m = _setup(ctx, pkg, ell)
    assert hashlib.sha256(crs_enc(ctx, crs)).hexdigest() == g["crs_sha256"]
    assert perm == g["perm"]
    assert "%064x" % fr_dec(k) == g["k"]
    assert ["%064x" % fr_dec(rs_m[i:i + 32]) for i in range(0, 128, 32)] == g["rs_m"]
    M_aff = aff_enc(jac_dec(M))
    inst = ctx.g1_compress(Rs + Ss + Ts + Us + M_aff)
    assert hashlib.sha256(inst).hexdigest() == g["instance_sha256"]
    proof = ctx.prove(crs, Rs, Ss, Ts, Us, M, perm, k, rs_m, pkg.Rand(42))
    assert proof.hex() == g["proof"]
    assert ctx.verify(crs, proof, Rs, Ss, Ts, Us, M, pkg.Rand(43)) is True
    # soundness mutations (curdleproof_test.go:48-167): verdict false, no error
    mut = g["mutations"]
    assert ctx.verify(crs, proof, Ss, Rs, Ts, Us, M, pkg.Rand(43)) is mut["swap_Rs_Ss"] is False
    p2 = pkg.Rand(5).generate_permutation(ell)
    perm_pts = lambda pts: b"".join(pts[96 * j:96 * j + 96] for j in p2)  # noqa: E731
    assert ctx.verify(crs, proof, Rs, Ss, perm_pts(Ts), perm_pts(Us), M, pkg.Rand(43)) is mut["repermute_Ts_Us"] is False
    kM = ctx.g1_scalar_mul_affine(M_aff, k, broadcast=True)
    assert ctx.verify(crs, proof, Rs, Ss, Ts, Us, jac_enc(aff_dec(kM)), pkg.Rand(43)) is mut["M_times_k"] is False
    k2 = pkg.Rand(9).get_fr()
    Ts2 = ctx.g1_scalar_mul_affine(Ts, k2, broadcast=True)
    Us2 = ctx.g1_scalar_mul_affine(Us, k2, broadcast=True)
    assert ctx.verify(crs, proof, Rs, Ss, Ts2, Us2, M, pkg.Rand(43)) is mut["rescale_Ts_Us"] is False


def test_verify_errors_and_malformed(ctx, pkg):
    ell = 12
    crs, Rs, Ss, Ts, Us, M, perm, k, rs_m = _setup(ctx, pkg, ell)
    proof = ctx.prove(crs, Rs, Ss, Ts, Us, M, perm, k, rs_m, pkg.Rand(42))
    # "randomizer is zero": Ts[0] at infinity -> (false, err)
    with pytest.raises(pkg.CdlError) as ei:
        ctx.verify(crs, proof, Rs, Ss, bytes(96) + Ts[96:], Us, M, pkg.Rand(43))
    assert ei.value.code == -5 and "randomizer is zero" in ei.value.msg
    # truncated proof / trailing garbage flag / non-canonical scalar -> decode error
    with pytest.raises(pkg.CdlError) as ei:
        ctx.verify(crs, proof[:-1], Rs, Ss, Ts, Us, M, pkg.Rand(43))
    assert ei.value.code == -4
    bad = bytearray(proof)
    bad[0] |= 0xE0
    with pytest.raises(pkg.CdlError) as ei:
        ctx.verify(crs, bytes(bad), Rs, Ss, Ts, Us, M, pkg.Rand(43))
    assert ei.value.code == -4
    bad = bytearray(proof)
    bad[-32:] = b"\xff" * 32  # x >= r
    with pytest.raises(pkg.CdlError) as ei:
        ctx.verify(crs, bytes(bad), Rs, Ss, Ts, Us, M, pkg.Rand(43))
    assert ei.value.code == -4
    # a flipped scalar bit keeps the encoding valid: verdict false, no error
    bad = bytearray(proof)
    bad[-1] ^= 1
    assert ctx.verify(crs, bytes(bad), Rs, Ss, Ts, Us, M, pkg.Rand(43)) is False
    # a wrong round count is an error (the reference would index out of range)
    m = 4
    off = 48 * 12  # A T1 T2 U1 U2 R S | B | C | B_c B_d  = 11 points, + Rp (32 B) -> first length prefix
    off = 48 * 9 + 32 + 48 * 2
    assert int.from_bytes(proof[off:off + 4], "big") == m
    bad = proof[:off] + (m - 1).to_bytes(4, "big") + proof[off + 4 + 48:]  # drop one L_C
    with pytest.raises(pkg.CdlError):
        ctx.verify(crs, bad, Rs, Ss, Ts, Us, M, pkg.Rand(43))


def test_whisk_batch_matches_single_and_oracle_verdicts(ctx, pkg):
    """Config-4 shape at small size: independent instances in lock step give the
    same bytes / verdicts as one-at-a-time calls; mutated proofs are rejected."""
    ell, B = 12, 5
    crs = ctx.generate_crs(ell, pkg.Rand(0))
    gen = aff_enc(b.G1_GEN)
    pres = []
    for i in range(B):
        r = pkg.Rand(1000 + i)
        ks, rs = [], []
        for _ in range(ell):
            ks.append(r.get_fr())
            rs.append(r.get_fr())
        rG = ctx.g1_scalar_mul_affine(gen * ell, b"".join(rs), broadcast=False)
        krG = ctx.g1_scalar_mul_affine(rG, b"".join(ks), broadcast=False)
        e1, e2 = ctx.g1_compress(rG), ctx.g1_compress(krG)
        pres.append(b"".join(e1[48 * j:48 * j + 48] + e2[48 * j:48 * j + 48] for j in range(ell)))
    single = [ctx.whisk_generate_shuffle_proof(crs, pres[i], pkg.Rand(3000 + i)) for i in range(B)]
    post, proofs, status = ctx.whisk_generate_shuffle_proof_batch(crs, b"".join(pres), [pkg.Rand(3000 + i) for i in range(B)])
    assert status == [0] * B
    for i in range(B):
        assert post[i * ell * 96:(i + 1) * ell * 96] == single[i][0]
        assert proofs[i * 4576:(i + 1) * 4576] == single[i][1]
    # mutate: 1 swap pre/post trackers, 2 flip a used byte of a scalar, 3 corrupt a point, 4 Ts[0] = infinity
    posts = [bytearray(post[i * ell * 96:(i + 1) * ell * 96]) for i in range(B)]
    prfs = [bytearray(proofs[i * 4576:(i + 1) * 4576]) for i in range(B)]
    pre_m = [bytearray(p) for p in pres]
    pre_m[1], posts[1] = posts[1], pre_m[1]
    used = len(bytes(prfs[2]).rstrip(b"\0"))
    prfs[2][used - 1] ^= 1
    prfs[3][48 + 5] ^= 0x55  # inside A's x coordinate: almost surely off-curve or another point
    posts[4][0:48] = bytes([0xC0]) + bytes(47)
    ok, st = ctx.whisk_is_valid_shuffle_proof_batch(crs, b"".join(bytes(p) for p in pre_m), b"".join(bytes(p) for p in posts),
                                                    b"".join(bytes(p) for p in prfs), [pkg.Rand(2000 + i) for i in range(B)])
    assert ok[0] == 1 and st[0] == 0
    assert ok[1] == 0 and st[1] == 0
    assert ok[2] == 0 and st[2] == 0
    assert ok[3] == 0  # decode error or plain reject, never accept
    assert ok[4] == 0 and st[4] == -5  # randomizer is zero
    # the same verdicts from the CPU oracle
    from oracle import protocol as P, whisk as W
    from oracle.cbackend import CBackend
    from oracle.rand import Rand as ORand
    P.set_backend(CBackend())
    try:
        pts = [b.g1_decompress(bytes(crs_enc(ctx, crs))[48 * j:48 * j + 48]) for j in range(ell + 9)]
        ocrs = P.CRS(pts[:ell], pts[ell:ell + 4], pts[ell + 4], pts[ell + 5], pts[ell + 6], pts[ell + 7], pts[ell + 8])
        for i in range(B):
            pre_t = [(bytes(pre_m[i][96 * j:96 * j + 48]), bytes(pre_m[i][96 * j + 48:96 * j + 96])) for j in range(ell)]
            post_t = [(bytes(posts[i][96 * j:96 * j + 48]), bytes(posts[i][96 * j + 48:96 * j + 96])) for j in range(ell)]
            try:
                want = W.is_valid_whisk_shuffle_proof(ocrs, pre_t, post_t, bytes(prfs[i]), ORand(2000 + i))
                err = False
            except (W.WhiskError, P.ProofError):
                want, err = False, True
            assert bool(ok[i]) == want and (st[i] != 0) == err, i
    finally:
        P.set_backend(P.PyBackend())


def test_lanes_give_the_same_bytes_and_verdicts(ctx, pkg):
    """A batch cut into concurrent lanes (own stream / engine / host thread each) returns exactly
    what one lane returns; every 8th instance is mutated (config 4 of BASELINE.json)."""
    ell, B = 12, 100
    crs = ctx.generate_crs(ell, pkg.Rand(0))
    gen = aff_enc(b.G1_GEN)
    pres = []
    for i in range(4):
        r = pkg.Rand(1000 + i)
        ks, rs = [], []
        for _ in range(ell):
            ks.append(r.get_fr())
            rs.append(r.get_fr())
        rG = ctx.g1_scalar_mul_affine(gen * ell, b"".join(rs), broadcast=False)
        krG = ctx.g1_scalar_mul_affine(rG, b"".join(ks), broadcast=False)
        e1, e2 = ctx.g1_compress(rG), ctx.g1_compress(krG)
        pres.append(b"".join(e1[48 * j:48 * j + 48] + e2[48 * j:48 * j + 48] for j in range(ell)))
    pre = b"".join(pres[i % 4] for i in range(B))
    out = {}
    try:
        for lanes in (1, 3):
            ctx.set_lanes(lanes)
            post, proofs, status = ctx.whisk_generate_shuffle_proof_batch(crs, pre, [pkg.Rand(3000 + i) for i in range(B)])
            assert status == [0] * B
            tb = ell * 96
            pre_m, post_m = bytearray(pre), bytearray(post)
            for i in range(0, B, 8):  # mutation: swap the instance's pre and post trackers
                pre_m[i * tb:(i + 1) * tb], post_m[i * tb:(i + 1) * tb] = post[i * tb:(i + 1) * tb], pre[i * tb:(i + 1) * tb]
            ok, st = ctx.whisk_is_valid_shuffle_proof_batch(crs, bytes(pre_m), bytes(post_m), proofs,
                                                            [pkg.Rand(2000 + i) for i in range(B)])
            assert st == [0] * B
            assert ok == [0 if i % 8 == 0 else 1 for i in range(B)]
            out[lanes] = (post, proofs, ok)
    finally:
        ctx.set_lanes(4)
    assert out[1] == out[3]


def test_whisk_tracker_proofs_match_oracle(ctx, pkg):
    """whisk.GenerateWhiskTrackerProof / IsValidWhiskTrackerProof (whisk/whisk.go:116-176), batched:
    byte-identical proofs and identical verdicts / errors against the oracle restatement."""
    from oracle import protocol as P, whisk as W
    from oracle.cbackend import CBackend
    from oracle.rand import Rand as ORand
    from util import fr_enc

    B = 9
    P.set_backend(CBackend())
    try:
        ks, trackers, kcomms, oproofs = [], [], [], []
        for i in range(B):
            r = ORand(500 + i)
            k = r.get_fr()
            rr = r.get_fr()
            rG = b.g1_mul(b.G1_GEN, rr)
            tr = W.new_tracker(rG, b.g1_mul(rG, k))
            ks.append(k)
            trackers.append(tr)
            kcomms.append(b.g1_compress(b.g1_mul(b.G1_GEN, k)))
            oproofs.append(W.generate_whisk_tracker_proof(tr, k, ORand(900 + i)))
        tbytes = b"".join(t[0] + t[1] for t in trackers)
        proofs, status = ctx.whisk_generate_tracker_proof_batch(tbytes, b"".join(fr_enc(k) for k in ks),
                                                                [pkg.Rand(900 + i) for i in range(B)])
        assert status == [0] * B
        assert [proofs[128 * i:128 * i + 128] for i in range(B)] == oproofs
        # validation with mutations: 1 wrong commitment, 2 flipped scalar bit, 3 corrupted A, 4 tracker of another k,
        # 5 non-canonical scalar, 6 undecodable tracker point
        pm = [bytearray(proofs[128 * i:128 * i + 128]) for i in range(B)]
        km = [bytearray(x) for x in kcomms]
        tm = [bytearray(tbytes[96 * i:96 * i + 96]) for i in range(B)]
        km[1] = bytearray(kcomms[2])
        pm[2][127] ^= 1
        pm[3][5] ^= 0x10
        tm[4] = bytearray(tbytes[96 * 5:96 * 5 + 96])
        pm[5][96:128] = (b.R + 5).to_bytes(32, "big")
        tm[6][0:48] = bytes([0x9F]) + b"\xff" * 47
        ok, st = ctx.whisk_is_valid_tracker_proof_batch(b"".join(bytes(t) for t in tm), b"".join(bytes(x) for x in km),
                                                        b"".join(bytes(p) for p in pm))
        for i in range(B):
            tr = (bytes(tm[i][:48]), bytes(tm[i][48:]))
            try:
                want, err = W.is_valid_whisk_tracker_proof(tr, bytes(km[i]), bytes(pm[i])), False
            except (W.WhiskError, b.DecodeError):
                want, err = False, True
            assert bool(ok[i]) == want and (st[i] != 0) == err, (i, ok[i], st[i], want, err)
        assert ok[0] == 1 and ok[7] == 1 and ok[1] == 0 and ok[2] == 0 and ok[4] == 0
    finally:
        P.set_backend(P.PyBackend())
