"""Byte encodings at the C ABI (gnark-crypto memory layouts) for host programs that drive the
library from Python: BLS12-381 constants and Montgomery-form packing.  Plain integers only —
no curve arithmetic happens on this side."""
from __future__ import annotations

P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
G1_GEN = (
    0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
    0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1,
)

RP = pow(2, 384, P)       # fp.Element Montgomery radix
RP_INV = pow(RP, -1, P)
RR = pow(2, 256, R)       # fr.Element Montgomery radix
RR_INV = pow(RR, -1, R)


def fp_enc(x: int) -> bytes:
    """gnark fp.Element: 6 little-endian 64-bit limbs of x * 2^384 mod p."""
    return (x * RP % P).to_bytes(48, "little")


def fp_dec(b: bytes) -> int:
    return int.from_bytes(b, "little") * RP_INV % P


def fr_enc(x: int) -> bytes:
    """gnark fr.Element: 4 little-endian 64-bit limbs of x * 2^256 mod r."""
    return (x % R * RR % R).to_bytes(32, "little")


def fr_dec(b: bytes) -> int:
    return int.from_bytes(b, "little") * RR_INV % R


def aff_enc(pt) -> bytes:
    """gnark G1Affine {X, Y}; None (infinity) is (0, 0)."""
    if pt is None:
        return bytes(96)
    return fp_enc(pt[0]) + fp_enc(pt[1])


def aff_dec(b: bytes):
    x = int.from_bytes(b[:48], "little")
    y = int.from_bytes(b[48:96], "little")
    if x == 0 and y == 0:
        return None
    return (x * RP_INV % P, y * RP_INV % P)
