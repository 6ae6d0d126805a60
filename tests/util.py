"""Test helpers: conversions between the oracle's python ints / affine tuples
and gnark's in-memory layouts used at the C ABI."""
from __future__ import annotations

from oracle import bls12381 as bls

P, R = bls.P, bls.R
RP = pow(2, 384, P)
RP_INV = pow(RP, -1, P)
RR = pow(2, 256, R)
RR_INV = pow(RR, -1, R)


def fp_enc(x: int) -> bytes:
    return (x * RP % P).to_bytes(48, "little")


def fp_dec(b: bytes) -> int:
    return int.from_bytes(b, "little") * RP_INV % P


def fr_enc(x: int) -> bytes:
    return (x % R * RR % R).to_bytes(32, "little")


def fr_dec(b: bytes) -> int:
    return int.from_bytes(b, "little") * RR_INV % R


def aff_enc(pt) -> bytes:
    if pt is None:
        return bytes(96)
    return fp_enc(pt[0]) + fp_enc(pt[1])


def aff_dec(b: bytes):
    x = int.from_bytes(b[:48], "little")
    y = int.from_bytes(b[48:96], "little")
    if x == 0 and y == 0:
        return None
    return (x * RP_INV % P, y * RP_INV % P)


def affs_enc(pts) -> bytes:
    return b"".join(aff_enc(p) for p in pts)


def affs_dec(b: bytes):
    return [aff_dec(b[i:i + 96]) for i in range(0, len(b), 96)]


def frs_enc(xs) -> bytes:
    return b"".join(fr_enc(x) for x in xs)


def jac_enc(pt, z: int = 1) -> bytes:
    """Jacobian (x z^2, y z^3, z); infinity -> (1, 1, 0)."""
    if pt is None:
        return fp_enc(1) + fp_enc(1) + bytes(48)
    return fp_enc(pt[0] * z * z % P) + fp_enc(pt[1] * z * z * z % P) + fp_enc(z)


def jac_dec(b: bytes):
    X, Y, Z = fp_dec(b[:48]), fp_dec(b[48:96]), fp_dec(b[96:144])
    if Z == 0:
        return None
    zi = pow(Z, -1, P)
    return (X * zi * zi % P, Y * zi * zi * zi % P)
