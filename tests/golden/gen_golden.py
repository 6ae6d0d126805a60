"""Generates the golden fixtures under tests/golden/ from the CPU oracle.

The reference ships no golden vectors and cannot be run here
(no Go toolchain), so these pin the *oracle's* outputs — itself pinned to public
KATs and cross-checked between its pure-Python and C restatements — for the
reference's own test flows:
  whisk_ell{4,12,124}.json : whisk/whisk_test.go:36-56 (one continuing Rand(0))
  prove_ell{12,60,508}.json: curdleproof_test.go:16-46 / 239-274 with the
      permutation drawn from Rand(42).GeneratePermutation (Go's math/rand
      shuffle cannot be reproduced without Go — SURVEY.md §8c-iii).
Usage:  python tests/golden/gen_golden.py [--pure-python-max-ell 12]
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import bls12381 as bls  # noqa: E402
from oracle import protocol as P  # noqa: E402
from oracle import whisk as W  # noqa: E402
from oracle.cbackend import CBackend  # noqa: E402
from oracle.rand import Rand  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def enc_pts(pts) -> str:
    return b"".join(bls.g1_compress(p) for p in pts).hex()


def crs_points(crs):
    return list(crs.Gs) + list(crs.Hs) + [crs.H, crs.Gt, crs.Gu, crs.Gsum, crs.Hsum]


def whisk_fixture(ell: int, backend):
    P.set_backend(backend)
    rand = Rand(0, backend=backend if isinstance(backend, CBackend) else None)
    crs = P.generate_crs(ell, rand)
    pre = W.generate_shuffle_trackers(rand, ell)
    post, proof = W.generate_whisk_shuffle_proof(crs, pre, rand, ell=ell)
    ok = W.is_valid_whisk_shuffle_proof(crs, pre, post, proof, rand)
    tail = rand.get_fr()  # pins the RNG position after the round trip
    return {
        "flow": "whisk/whisk_test.go:36-56 TestWhiskShuffleProof, Rand(0) continuing",
        "ell": ell,
        "crs": enc_pts(crs_points(crs)),
        "pre_trackers": b"".join(a + b for a, b in pre).hex(),
        "post_trackers": b"".join(a + b for a, b in post).hex(),
        "proof": proof.hex(),
        "proof_used_bytes": len(proof.rstrip(b"\0")),
        "valid": ok,
        "next_fr_after_roundtrip": "%064x" % tail,
    }


def prove_fixture(ell: int, backend):
    P.set_backend(backend)
    rand = Rand(0, backend=backend if isinstance(backend, CBackend) else None)
    crs = P.generate_crs(ell, rand)
    perm = Rand(42).generate_permutation(ell)
    k = rand.get_fr()
    Rs = rand.get_g1_affines(ell)
    Ss = rand.get_g1_affines(ell)
    Ts, Us, M, rs_m = P.shuffle_permute_commit(crs.Gs, crs.Hs, Rs, Ss, perm, k, rand)
    proof = P.prove(crs, Rs, Ss, Ts, Us, M, perm, k, rs_m, Rand(42))
    pb = proof.serialize()
    ok = P.verify(P.Proof.deserialize(pb), crs, Rs, Ss, Ts, Us, M, Rand(43))
    # the four soundness mutations of curdleproof_test.go:48-167
    p2 = Rand(5).generate_permutation(ell)
    k2 = Rand(9).get_fr()
    be = P.get_backend()
    mut = {
        "swap_Rs_Ss": P.verify(proof, crs, Ss, Rs, Ts, Us, M, Rand(43)),
        "repermute_Ts_Us": P.verify(proof, crs, Rs, Ss, P.permute(Ts, p2), P.permute(Us, p2), M, Rand(43)),
        "M_times_k": P.verify(proof, crs, Rs, Ss, Ts, Us, be.mul(M, k), Rand(43)),
        "rescale_Ts_Us": P.verify(proof, crs, Rs, Ss, be.mul_batch(Ts, [k2] * ell), be.mul_batch(Us, [k2] * ell), M, Rand(43)),
    }
    inputs = enc_pts(Rs) + enc_pts(Ss) + enc_pts(Ts) + enc_pts(Us) + enc_pts([M])
    return {
        "flow": "curdleproof_test.go:16-46 (setup :239-274; perm = Rand(42).GeneratePermutation)",
        "ell": ell,
        "crs_sha256": hashlib.sha256(bytes.fromhex(enc_pts(crs_points(crs)))).hexdigest(),
        "instance_sha256": hashlib.sha256(bytes.fromhex(inputs)).hexdigest(),
        "M": enc_pts([M]),
        "rs_m": ["%064x" % x for x in rs_m],
        "k": "%064x" % k,
        "perm": perm,
        "proof": pb.hex(),
        "valid": ok,
        "mutations": mut,
    }


def main():
    cb = CBackend()
    py = P.PyBackend()
    jobs = [("whisk", 4), ("whisk", 12), ("whisk", 124), ("prove", 12), ("prove", 60), ("prove", 508)]
    for kind, ell in jobs:
        fx = (whisk_fixture if kind == "whisk" else prove_fixture)(ell, cb)
        if ell <= 12:  # the two restatements must agree before a fixture is written
            from oracle import merlin
            merlin.set_keccak(merlin.keccak_f1600)
            bls.set_decompress_batch(None)
            ref = (whisk_fixture if kind == "whisk" else prove_fixture)(ell, py)
            assert ref == fx, f"pure-Python and C oracles disagree for {kind} ell={ell}"
            merlin.set_keccak(cb.keccak)
            bls.set_decompress_batch(cb.decompress)
        path = os.path.join(HERE, f"{kind}_ell{ell}.json")
        with open(path, "w") as fh:
            json.dump(fx, fh, indent=1)
        print(path, "valid" if fx["valid"] else "INVALID", len(fx["proof"]) // 2, "bytes")


if __name__ == "__main__":
    main()
