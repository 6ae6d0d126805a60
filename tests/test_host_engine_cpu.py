"""Host-side pieces of the product library that need no GPU: common.Rand,
Merlin transcript, Fr arithmetic — checked against the oracle / public KATs."""
import ctypes as C
import random

from oracle import bls12381 as b
from oracle.rand import Rand as ORand
from util import R, fr_dec, fr_enc


def test_rand_matches_oracle(pkg):
    for seed in (0, 42, 43, 2**63 + 5):
        r = pkg.Rand(seed)
        o = ORand(seed)
        got = r.get_frs(40)
        assert [fr_dec(got[i:i + 32]) for i in range(0, len(got), 32)] == o.get_frs(40)
        assert r.generate_permutation(124) == o.generate_permutation(124)
        assert fr_dec(r.get_fr()) == o.get_fr()


def test_merlin_kat_and_fr(pkg):
    lib = pkg.load_library()
    random.seed(3)
    for _ in range(20):
        a, x = random.randrange(1, R), random.randrange(1, R)
        out = C.create_string_buffer(32)
        fr_out = C.create_string_buffer(32)
        assert lib.cdl_host_selftest(out, fr_enc(a), fr_enc(x), fr_out) == 0
        assert out.raw.hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"
        want = b.fr_inv((a * x + a - x) % R) * pow(a, 5, R) % R
        assert fr_dec(fr_out.raw) == want
    # edge values of the inversion (b = 0 makes the inverted value a itself)
    for a in (1, 2, 3, R - 1, R - 2, (R + 1) // 2, (R - 1) // 2, 2**255 % R, 2**128, 2**64 - 1):
        out = C.create_string_buffer(32)
        fr_out = C.create_string_buffer(32)
        assert lib.cdl_host_selftest(out, fr_enc(a), fr_enc(0), fr_out) == 0
        assert fr_dec(fr_out.raw) == pow(a, -1, R) * pow(a, 5, R) % R


def test_eight_way_transcript_hashing_matches_scalar(pkg):
    """Proofs of a batch hash as cooperating fibers whose Keccak-f permutations are batched eight at a
    time (csrc/host/fiber.hpp, keccak_x8.cpp): every transcript / SHAKE stream must produce exactly the
    bytes it produces on its own, for group sizes that fill, underfill and overflow a group."""
    import os
    import subprocess
    import sys

    lib = pkg.load_library()
    digests = {}
    for n in (1, 2, 7, 8, 9, 33, 100):
        out = C.create_string_buffer(32)
        used = C.c_int32(-1)
        assert lib.cdl_host_selftest_fibers(n, 12, out, C.byref(used)) == 0, n
        assert used.value in (0, 1)
        digests[n] = out.raw.hex()
    # the same digests with the eight-way path disabled (plain scalar hashing)
    code = ("import importlib,ctypes as C,sys; sys.path.insert(0, %r); p=importlib.import_module('go-curdleproofs_b200'); "
            "l=p.load_library(); o=C.create_string_buffer(32); u=C.c_int32(); "
            "print([(l.cdl_host_selftest_fibers(n,12,o,C.byref(u)), u.value, o.raw.hex()) for n in (1,2,7,8,9,33,100)])"
            % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env = dict(os.environ, CDL_NO_FIBERS="1")
    res = eval(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout)
    assert [r[0] for r in res] == [0] * 7 and [r[1] for r in res] == [0] * 7
    assert [r[2] for r in res] == [digests[n] for n in (1, 2, 7, 8, 9, 33, 100)]
