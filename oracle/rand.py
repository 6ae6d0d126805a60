"""common.Rand — the reference's deterministic SHAKE256 RNG
(/root/reference/common/rand.go).  TEST INFRASTRUCTURE ONLY.

golang.org/x/crypto/sha3.NewShake256 (go.mod:9) is standard SHAKE256, so
hashlib.shake_256 is an exact stand-in: SHAKE output is prefix-consistent and
a streaming Read() is a slice of one long digest.
"""
from __future__ import annotations

import hashlib

from . import bls12381 as bls


class Rand:
    def __init__(self, seed: int, backend=None):  # rand.go:19-33
        self._seed = int(seed).to_bytes(8, "big")
        self._buf = b""
        self._pos = 0
        self._backend = backend

    def read(self, n: int) -> bytes:
        need = self._pos + n
        if need > len(self._buf):
            size = max(4096, 2 * len(self._buf), need)
            self._buf = hashlib.shake_256(self._seed).digest(size)
        out = self._buf[self._pos:need]
        self._pos = need
        return out

    def get_fr(self) -> int:  # rand.go:35-47 (rejection sampling, 32-byte BE)
        while True:
            v = int.from_bytes(self.read(32), "big")
            if v < bls.R:
                return v

    def get_frs(self, n: int):  # rand.go:49-59
        return [self.get_fr() for _ in range(n)]

    def get_g1_affine(self):  # rand.go:72-83
        s = self.get_fr()
        if self._backend is not None:
            return self._backend.mul_batch([bls.G1_GEN], [s])[0]
        return bls.g1_mul(bls.G1_GEN, s)

    def get_g1_affines(self, n: int):  # rand.go:85-95
        ss = [self.get_fr() for _ in range(n)]
        if self._backend is not None:
            return self._backend.mul_batch([bls.G1_GEN] * n, ss)
        return bls.g1_mul_batch([bls.G1_GEN] * n, ss)

    get_g1_jac = get_g1_affine  # rand.go:61-70 (same point, lifted)

    def generate_permutation(self, n: int):  # rand.go:97-113
        perm = list(range(n))
        for i in range(n):
            tmp = int.from_bytes(self.read(16)[:2], "big")
            j = tmp % (i + 1)
            perm[i], perm[j] = perm[j], perm[i]
        return perm
