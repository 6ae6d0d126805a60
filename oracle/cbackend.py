"""ctypes front-end of the C restatement (oracle/c/curdle_oracle.c) exposing
the same backend interface as protocol.PyBackend.  TEST INFRASTRUCTURE ONLY
(see oracle/__init__.py): checker at large sizes + the timed CPU "port"
baseline.  Cross-checked against the pure-Python restatement in
tests/test_oracle_c.py."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from . import bls12381 as bls
from . import merlin

HERE = os.path.dirname(os.path.abspath(__file__))
CDIR = os.path.join(HERE, "c")
LIB = os.path.join(CDIR, "libcurdle_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(CDIR, "curdle_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-C", CDIR], check=True)
    return LIB


def _pt(p) -> bytes:
    if p is None:
        return bytes(96)
    return p[0].to_bytes(48, "little") + p[1].to_bytes(48, "little")


def _unpt(b: bytes):
    x = int.from_bytes(b[:48], "little")
    y = int.from_bytes(b[48:96], "little")
    return None if x == 0 and y == 0 else (x, y)


class CBackend:
    name = "c-6x64-montgomery"

    def __init__(self, threads: int | None = None, accelerate_keccak: bool = True):
        self.lib = C.CDLL(build())
        self.threads = threads or (os.cpu_count() or 1)
        vp, sz = C.c_void_p, C.c_size_t
        self.lib.co_g1_mul_batch.argtypes = [vp, vp, sz, sz, vp]
        self.lib.co_g1_fold.argtypes = [vp, vp, vp, sz, vp]
        self.lib.co_g1_msm.argtypes = [vp, vp, sz, vp, C.c_int]
        self.lib.co_g1_sum.argtypes = [vp, sz, vp]
        self.lib.co_g1_compress.argtypes = [vp, sz, vp]
        self.lib.co_g1_decompress.argtypes = [vp, sz, vp, vp]
        self.lib.co_keccak_f1600.argtypes = [vp]
        if accelerate_keccak:
            merlin.set_keccak(self.keccak)
            bls.set_decompress_batch(self.decompress)

    def keccak(self, state: bytearray) -> None:
        buf = (C.c_uint8 * 200).from_buffer(state)
        self.lib.co_keccak_f1600(buf)

    # --- raw byte-level entry points (used by the large-size tests / bench)
    def msm_raw(self, points: bytes, scalars: bytes, n: int, threads: int | None = None) -> bytes:
        out = C.create_string_buffer(96)
        self.lib.co_g1_msm(points, scalars, n, out, threads or self.threads)
        return out.raw

    # --- backend interface
    def msm(self, points, scalars):
        if len(points) != len(scalars):
            raise ValueError("len(points) != len(scalars)")
        n = len(points)
        pts = b"".join(_pt(p) for p in points)
        sc = b"".join((s % bls.R).to_bytes(32, "little") for s in scalars)
        return _unpt(self.msm_raw(pts, sc, n))

    def mul_batch(self, pts, ks):
        n = len(pts)
        if n == 0:
            return []
        out = C.create_string_buffer(96 * n)
        self.lib.co_g1_mul_batch(b"".join(_pt(p) for p in pts),
                                 b"".join((k % bls.R).to_bytes(32, "little") for k in ks), n, 1, out)
        return [_unpt(out.raw[96 * i:96 * i + 96]) for i in range(n)]

    def mul(self, pt, k):
        return self.mul_batch([pt], [k])[0]

    def fold(self, L, Rr, x):
        n = len(L)
        out = C.create_string_buffer(96 * max(1, n))
        self.lib.co_g1_fold(b"".join(_pt(p) for p in L), b"".join(_pt(p) for p in Rr),
                            (x % bls.R).to_bytes(32, "little"), n, out)
        return [_unpt(out.raw[96 * i:96 * i + 96]) for i in range(n)]

    def add(self, a, b):
        return bls.g1_add(a, b)

    def sub(self, a, b):
        return bls.g1_sub(a, b)

    def sum(self, pts):
        out = C.create_string_buffer(96)
        self.lib.co_g1_sum(b"".join(_pt(p) for p in pts), len(pts), out)
        return _unpt(out.raw)

    def compress(self, pts) -> bytes:
        out = C.create_string_buffer(48 * max(1, len(pts)))
        self.lib.co_g1_compress(b"".join(_pt(p) for p in pts), len(pts), out)
        return out.raw[:48 * len(pts)]

    def decompress(self, enc: bytes):
        n = len(enc) // 48
        out = C.create_string_buffer(96 * max(1, n))
        st = C.create_string_buffer(max(1, n))
        self.lib.co_g1_decompress(enc, n, out, st)
        return [_unpt(out.raw[96 * i:96 * i + 96]) for i in range(n)], list(st.raw[:n])
