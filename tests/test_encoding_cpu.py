"""go-curdleproofs_b200/encoding.py (host-side packing used by bench.py and the tools) agrees with
the oracle's constants and with the test helpers."""
import importlib
import random

import util
from oracle import bls12381 as b


def test_constants_and_packing():
    enc = importlib.import_module("go-curdleproofs_b200.encoding")
    assert enc.P == b.P and enc.R == b.R and enc.G1_GEN == b.G1_GEN
    assert b.is_on_curve(enc.G1_GEN)
    random.seed(4)
    for _ in range(50):
        x, k = random.randrange(b.P), random.randrange(b.R)
        assert enc.fp_enc(x) == util.fp_enc(x) and enc.fp_dec(enc.fp_enc(x)) == x
        assert enc.fr_enc(k) == util.fr_enc(k) and enc.fr_dec(enc.fr_enc(k)) == k
    pt = b.g1_mul(b.G1_GEN, 7)
    assert enc.aff_enc(pt) == util.aff_enc(pt) and enc.aff_dec(enc.aff_enc(pt)) == pt
    assert enc.aff_enc(None) == bytes(96) and enc.aff_dec(bytes(96)) is None
