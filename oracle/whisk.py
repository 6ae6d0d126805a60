"""CPU restatement of the Whisk wrapper (/root/reference/whisk/whisk.go,
whisk/types.go).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Trackers are (rG_bytes, krG_bytes) pairs of 48-byte compressed points
(whisk/types.go:74-95); a shuffle proof is a 4576-byte array holding
``M | curdleproof.Proof`` and zero padding (whisk/types.go:53-72).
"""
from __future__ import annotations

from . import bls12381 as bls
from . import protocol as proto
from .merlin import Transcript
from .rand import Rand

G1POINT_SIZE = 48  # whisk/types.go:15
N = 128  # :17
ELL = N - proto.N_BLINDERS  # :18
TRACKER_PROOF_SIZE = 128  # :20
WHISK_SHUFFLE_PROOF_SIZE = 4576  # :21

R = bls.R


class WhiskError(Exception):
    pass


def new_tracker(rG, krG):  # types.go:79-84
    return (bls.g1_compress(rG), bls.g1_compress(krG))


def tracker_points(tracker):  # types.go:86-95
    try:
        return bls.g1_decompress(tracker[0]), bls.g1_decompress(tracker[1])
    except bls.DecodeError as e:
        raise WhiskError(f"failed to set point: {e}")


def serialize_shuffle_proof(M, proof: proto.Proof, size: int = WHISK_SHUFFLE_PROOF_SIZE) -> bytes:  # types.go:53-72
    e = bls.Encoder()
    e.point(M)
    proof.serialize_into(e)
    b = e.bytes()
    if len(b) > size:
        # Go's copy() would silently truncate; no reference config reaches this.
        raise WhiskError("proof larger than WHISK_SHUFFLE_PROOF_SIZE")
    return b + bytes(size - len(b))


def deserialize_shuffle_proof(buf: bytes):  # types.go:39-51
    d = bls.Decoder(buf)
    M = d.point()
    proof = proto.Proof.from_reader(d)
    return M, proof


def generate_whisk_shuffle_proof(crs: proto.CRS, pre_trackers, rand: Rand, ell: int = ELL,
                                 proof_size: int = WHISK_SHUFFLE_PROOF_SIZE):
    """whisk.GenerateWhiskShuffleProof (whisk.go:63-114).  ``ell``/``proof_size``
    generalise the compile-time constants for the small test shapes."""
    perm = rand.generate_permutation(ell)
    k = rand.get_fr()
    try:
        flat = bls.g1_decompress_many([x for t in pre_trackers for x in t])
    except bls.DecodeError as e:
        raise WhiskError(f"getting points: {e}")
    Rs, Ss = flat[0::2], flat[1::2]
    Ts, Us, M, rs_m = proto.shuffle_permute_commit(crs.Gs, crs.Hs, Rs, Ss, perm, k, rand)
    proof = proto.prove(crs, Rs, Ss, Ts, Us, M, perm, k, rs_m, rand)
    proof_bytes = serialize_shuffle_proof(M, proof, proof_size)
    post = [new_tracker(Ts[i], Us[i]) for i in range(len(pre_trackers))]
    return post, proof_bytes


def is_valid_whisk_shuffle_proof(crs: proto.CRS, pre, post, proof_bytes: bytes, rand: Rand) -> bool:
    """whisk.IsValidWhiskShuffleProof (whisk.go:20-61).  Returns the verdict;
    raises WhiskError / ProofError where the reference returns an error."""
    if len(pre) != len(post):
        raise WhiskError("pre and post shuffle trackers must be the same length")
    try:
        M, proof = deserialize_shuffle_proof(proof_bytes)
    except bls.DecodeError as e:
        raise WhiskError(f"decoding proof: {e}")
    try:
        flat = bls.g1_decompress_many([x for i in range(len(pre)) for x in (pre[i][0], pre[i][1], post[i][0], post[i][1])])
    except bls.DecodeError as e:
        raise WhiskError(f"getting shuffle points: {e}")
    Rs, Ss, Ts, Us = flat[0::4], flat[1::4], flat[2::4], flat[3::4]
    return proto.verify(proof, crs, Rs, Ss, Ts, Us, M, rand)


# ---------------------------------------------------------------------------
# tracker opening proofs (whisk.go:116-176) — constant work, host-side
# ---------------------------------------------------------------------------
def generate_whisk_tracker_proof(tracker, k: int, rand: Rand) -> bytes:  # whisk.go:149-176
    be = proto.get_backend()
    rG, krG = tracker_points(tracker)
    kG = be.mul(bls.G1_GEN, k)
    blinder = rand.get_fr()
    A = be.mul(bls.G1_GEN, blinder)
    B = be.mul(rG, blinder)
    tr = Transcript(b"whisk_opening_proof")
    tr.append_points(b"tracker_opening_proof", kG, bls.G1_GEN, krG, rG, A, B)
    ch = tr.get_and_append_challenge(b"tracker_opening_proof_challenge")
    s = (blinder - ch * k) % R
    e = bls.Encoder()
    e.point(A)
    e.point(B)
    e.scalar(s)
    return e.bytes()


def is_valid_whisk_tracker_proof(tracker, k_comm: bytes, proof_bytes: bytes) -> bool:  # whisk.go:116-147
    be = proto.get_backend()
    try:
        d = bls.Decoder(proof_bytes)
        A = d.point()
        B = d.point()
        s = d.scalar()
        kG = bls.g1_decompress(k_comm)
    except bls.DecodeError as e:
        raise WhiskError(f"decoding proof: {e}")
    rG, krG = tracker_points(tracker)
    tr = Transcript(b"whisk_opening_proof")
    tr.append_points(b"tracker_opening_proof", kG, bls.G1_GEN, krG, rG, A, B)
    ch = tr.get_and_append_challenge(b"tracker_opening_proof_challenge")
    A_prime = be.add(be.mul(bls.G1_GEN, s), be.mul(kG, ch))
    B_prime = be.add(be.mul(rG, s), be.mul(krG, ch))
    return A_prime == A and B_prime == B


def generate_shuffle_trackers(rand: Rand, n: int):
    """whisk_test.go generateShuffleTrackers/generateTracker: per tracker draw
    k then r; tracker = (r·G, k·r·G)."""
    be = proto.get_backend()
    ks, rs = [], []
    for _ in range(n):
        ks.append(rand.get_fr())
        rs.append(rand.get_fr())
    rGs = be.mul_batch([bls.G1_GEN] * n, rs)
    krGs = be.mul_batch(rGs, ks)
    return [new_tracker(rGs[i], krGs[i]) for i in range(n)]
