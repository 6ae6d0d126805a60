"""Pins the CPU oracle to public known-answer tests (the reference itself holds
no golden vectors — SURVEY.md §4/§8c)."""
import hashlib

from oracle import bls12381 as bls
from oracle import merlin
from oracle import protocol as P
from oracle.rand import Rand


def test_curve_constants():
    z = -bls.Z_ABS
    assert bls.R == z**4 - z**2 + 1
    assert bls.P == (z - 1) ** 2 * bls.R // 3 + z
    assert bls.P % 4 == 3
    assert bls.is_on_curve(bls.G1_GEN)
    assert bls.g1_mul_jac_raw(bls.G1_GEN, bls.R)[2] == 0


def test_generator_encodings():
    # eth2 BLS public keys for sk = 1 and sk = 2
    assert bls.g1_compress(bls.G1_GEN).hex() == (
        "97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb")
    assert bls.g1_compress(bls.g1_mul(bls.G1_GEN, 2)).hex() == (
        "a572cbea904d67468808c8eb50a9450c9721db309128012543902d0ac358a62ae28f75bb8f1c7c42c39a8c5529bf0f4e")
    assert bls.g1_compress(None) == bytes([0xC0]) + bytes(47)
    # ... and for sk = 3, 4, 5 (public eth2 interop / test-vector keys): scalar multiplication, the
    # sign bit of the encoding and decompression are pinned by them
    kat = {
        3: "89ece308f9d1f0131765212deca99697b112d61f9be9a5f1f3780a51335b3ff981747a0b2ca2179b96d2c0c9024e5224",
        4: "ac9b60d5afcbd5663a8a44b7c5a02f19e9a77ab0a35bd65809bb5c67ec582c897feb04decc694b13e08587f3ff9b5b60",
        5: "b0e7791fb972fe014159aa33a98622da3cdc98ff707965e536d8636b5fcc5ac7a91a8c46e59a00dca575af0f18fb13dc",
    }
    for k, enc in kat.items():
        pt = bls.g1_mul(bls.G1_GEN, k)
        assert bls.g1_compress(pt).hex() == enc
        assert bls.g1_decompress(bytes.fromhex(enc)) == pt


def test_keccak_against_hashlib():
    for msg in (b"", b"abc", bytes(range(135)), bytes(200)):
        st = bytearray(200)
        rate = 136
        data = bytearray(msg) + b"\x06"
        while len(data) % rate:
            data += b"\x00"
        data[-1] |= 0x80
        for off in range(0, len(data), rate):
            for i in range(rate):
                st[i] ^= data[off + i]
            merlin.keccak_f1600(st)
        assert bytes(st[:32]) == hashlib.sha3_256(msg).digest()


def test_merlin_kat():
    t = merlin.MerlinTranscript(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == (
        "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615")


def test_rand_stream_is_shake256():
    r = Rand(7)
    a = r.read(10) + r.read(5000) + r.read(33)
    assert a == hashlib.shake_256((7).to_bytes(8, "big")).digest(5043)
    r = Rand(0)
    v = r.get_fr()
    assert 0 <= v < bls.R
    perm = Rand(3).generate_permutation(50)
    assert sorted(perm) == list(range(50))


def test_ipa_literal():
    # common/util_test.go:26 — the only literal expected value in the reference
    assert P.ipa([1, 2, 3, 4], [2, 3, 4, 5]) == 40


def test_codec_roundtrip_and_rejects():
    import pytest
    pts = [bls.g1_mul(bls.G1_GEN, k) for k in (1, 2, 3, 12345678901234567890)] + [None]
    for pt in pts:
        enc = bls.g1_compress(pt)
        assert bls.g1_decompress(enc) == pt
    bad = bytearray(bls.g1_compress(pts[0]))
    bad[0] |= 0x40 | 0x20  # 0b111 mask
    with pytest.raises(bls.DecodeError):
        bls.g1_decompress(bytes(bad))
    with pytest.raises(bls.DecodeError):  # x >= p
        bls.g1_decompress(bytes([0x9F]) + b"\xff" * 47)
    # a curve point outside the r-torsion must be rejected
    x = 1
    while True:
        y = bls.fp_sqrt((x**3 + 4) % bls.P)
        if y is not None and not bls.g1_in_subgroup((x, y)):
            break
        x += 1
    with pytest.raises(bls.DecodeError):
        bls.g1_decompress(bls.g1_compress((x, y)))
    assert bls.g1_decompress(bls.g1_compress((x, y)), subgroup_check=False) == (x, y)


def test_msm_variants_agree():
    r = Rand(11)
    pts = r.get_g1_affines(20) + [None]
    sc = r.get_frs(19) + [0, 5]
    want = bls.g1_msm_naive(pts, sc)
    for c in (2, 3, 5, 8):
        assert bls.g1_msm(pts, sc, c=c) == want
