"""Pure-Python big-int restatement of the BLS12-381 G1 / Fr arithmetic and wire
formats the reference reaches through gnark-crypto v0.11.0 (go.mod:6; source
not in /root/reference — algorithms restated from the public BLS12-381 /
ZCash-serialisation specifications).  TEST INFRASTRUCTURE ONLY (see
oracle/__init__.py).  PARITY: unpinned vs a Go run; pinned to public KATs in
tests/test_oracle_kat.py.

Representation
  Fr element : python int in [0, r)
  Fp element : python int in [0, p)
  G1 point   : None (infinity) or an affine tuple (x, y)
Every value that is observable in the reference (transcript bytes, proof
bytes, verdicts) is canonical, so representation of intermediates is free
(SURVEY.md §8c).

Reference call sites covered (file:line in /root/reference):
  G1Jac.MultiExp            e.g. curdleproof.go:72, msmaccumulator.go:59
  G1Affine.ScalarMultiplication  common/util.go:57, grandproductargument.go:97
  G1Affine.Add / Sub        innerproductargument.go:160, grandproductargument.go:246
  BatchJacobianToAffineG1   transcript/transcript.go:26
  G1Affine.Bytes / SetBytes transcript/transcript.go:35, whisk/types.go:81-95
  Encoder / Decoder         curdleproof.go:320-387
"""
from __future__ import annotations

P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
B_COEFF = 4
G1_X = 0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB
G1_Y = 0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1
G1_GEN = (G1_X, G1_Y)
# BLS parameter z (negative): p = (z-1)^2 r / 3 + z,  r = z^4 - z^2 + 1
Z_ABS = 0xD201000000010000

FR_BYTES = 32
FP_BYTES = 48
G1_COMPRESSED = 48
G1_UNCOMPRESSED = 96

# ZCash-style flag bits as gnark-crypto uses them (ecc/bls12-381/marshal.go)
M_MASK = 0b111 << 5
M_UNCOMPRESSED = 0b000 << 5
M_UNCOMPRESSED_INF = 0b010 << 5
M_COMPRESSED_SMALLEST = 0b100 << 5
M_COMPRESSED_LARGEST = 0b101 << 5
M_COMPRESSED_INF = 0b110 << 5


class DecodeError(ValueError):
    pass


# ----------------------------------------------------------------------------
# Fr helpers (fr.Element semantics: Inverse(0) == 0, BatchInvert keeps zeros)
# ----------------------------------------------------------------------------
def fr_inv(a: int) -> int:
    a %= R
    return 0 if a == 0 else pow(a, -1, R)


def fr_batch_inv(v):
    return [fr_inv(a) for a in v]


def fr_to_bytes(a: int) -> bytes:
    """fr.Element.Bytes(): 32-byte big-endian canonical."""
    return (a % R).to_bytes(32, "big")


def fr_from_bytes_canonical(b: bytes) -> int:
    """fr.Element.SetBytesCanonical: error unless len==32 and value < r."""
    if len(b) != 32:
        raise DecodeError("fr: invalid length")
    v = int.from_bytes(b, "big")
    if v >= R:
        raise DecodeError("fr: non-canonical")
    return v


# ----------------------------------------------------------------------------
# G1 group law.  Jacobian (X, Y, Z) with a = 0 internally.
# ----------------------------------------------------------------------------
def is_on_curve(pt) -> bool:
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - B_COEFF) % P == 0


def _jac_dbl(X1, Y1, Z1):
    if Z1 == 0:
        return (1, 1, 0)
    A = X1 * X1 % P
    Bb = Y1 * Y1 % P
    C = Bb * Bb % P
    D = 2 * ((X1 + Bb) * (X1 + Bb) - A - C) % P
    E = 3 * A % P
    F = E * E % P
    X3 = (F - 2 * D) % P
    Y3 = (E * (D - X3) - 8 * C) % P
    Z3 = 2 * Y1 * Z1 % P
    return (X3, Y3, Z3)


def _jac_add(p1, p2):
    X1, Y1, Z1 = p1
    X2, Y2, Z2 = p2
    if Z1 == 0:
        return p2
    if Z2 == 0:
        return p1
    Z1Z1 = Z1 * Z1 % P
    Z2Z2 = Z2 * Z2 % P
    U1 = X1 * Z2Z2 % P
    U2 = X2 * Z1Z1 % P
    S1 = Y1 * Z2 * Z2Z2 % P
    S2 = Y2 * Z1 * Z1Z1 % P
    if U1 == U2:
        if S1 == S2:
            return _jac_dbl(X1, Y1, Z1)
        return (1, 1, 0)
    H = (U2 - U1) % P
    Rr = (S2 - S1) % P
    HH = H * H % P
    HHH = H * HH % P
    V = U1 * HH % P
    X3 = (Rr * Rr - HHH - 2 * V) % P
    Y3 = (Rr * (V - X3) - S1 * HHH) % P
    Z3 = Z1 * Z2 * H % P
    return (X3, Y3, Z3)


def _to_jac(pt):
    if pt is None:
        return (1, 1, 0)
    return (pt[0], pt[1], 1)


def _from_jac(j):
    X, Y, Z = j
    if Z == 0:
        return None
    zi = pow(Z, -1, P)
    zi2 = zi * zi % P
    return (X * zi2 % P, Y * zi2 * zi % P)


def _batch_from_jac(js):
    """Montgomery batch inversion (BatchJacobianToAffineG1)."""
    n = len(js)
    pref = [1] * (n + 1)
    for i, (_, _, Z) in enumerate(js):
        pref[i + 1] = pref[i] * (Z if Z else 1) % P
    inv = pow(pref[n], -1, P)
    out = [None] * n
    for i in range(n - 1, -1, -1):
        X, Y, Z = js[i]
        if Z == 0:
            continue
        zi = inv * pref[i] % P
        inv = inv * Z % P
        zi2 = zi * zi % P
        out[i] = (X * zi2 % P, Y * zi2 * zi % P)
    return out


def g1_neg(pt):
    if pt is None:
        return None
    return (pt[0], (-pt[1]) % P)


def g1_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    y3 = (lam * (x1 - x3) - y1) % P
    return (x3, y3)


def g1_sub(a, b):
    return g1_add(a, g1_neg(b))


def g1_eq(a, b) -> bool:
    return a == b


def g1_mul_jac(pt, k: int):
    """k·pt, plain left-to-right double-and-add, result Jacobian."""
    k %= R
    acc = (1, 1, 0)
    if pt is None or k == 0:
        return acc
    base = (pt[0], pt[1], 1)
    for bit in bin(k)[2:]:
        acc = _jac_dbl(*acc)
        if bit == "1":
            acc = _jac_add(acc, base)
    return acc


def g1_mul(pt, k: int):
    """G1Affine.ScalarMultiplication / G1Jac.ScalarMultiplication: k·pt."""
    return _from_jac(g1_mul_jac(pt, k))


def g1_mul_batch(pts, ks):
    """[k_i·P_i]; one shared batch inversion at the end."""
    return _batch_from_jac([g1_mul_jac(p, k) for p, k in zip(pts, ks)])


def g1_msm_naive(points, scalars):
    acc = (1, 1, 0)
    for pt, k in zip(points, scalars):
        acc = _jac_add(acc, g1_mul_jac(pt, k))
    return _from_jac(acc)


def g1_msm(points, scalars, c: int | None = None):
    """G1Jac.MultiExp(points, scalars): Σ scalars[i]·points[i].

    Bucket method (unsigned c-bit windows).  Only the resulting group element
    is observable, so the window schedule is free.  Infinity
    bases are skipped (identity); errors iff lengths differ.
    """
    if len(points) != len(scalars):
        raise ValueError("len(points) != len(scalars)")
    n = len(points)
    if n == 0:
        return None
    if n < 8:
        return g1_msm_naive(points, scalars)
    if c is None:
        c = 3 if n < 32 else (5 if n < 256 else (7 if n < 2048 else 9))
    scal = [s % R for s in scalars]
    nwin = (255 + c - 1) // c
    mask = (1 << c) - 1
    jpts = [_to_jac(p) for p in points]
    total = (1, 1, 0)
    for w in range(nwin - 1, -1, -1):
        for _ in range(c):
            total = _jac_dbl(*total)
        buckets = [None] * (mask + 1)
        sh = w * c
        for j, s in zip(jpts, scal):
            d = (s >> sh) & mask
            if d and j[2] != 0:
                bkt = buckets[d]
                buckets[d] = j if bkt is None else _jac_add(bkt, j)
        run = (1, 1, 0)
        acc = (1, 1, 0)
        for d in range(mask, 0, -1):
            if buckets[d] is not None:
                run = _jac_add(run, buckets[d])
            acc = _jac_add(acc, run)
        total = _jac_add(total, acc)
    return _from_jac(total)


def g1_sum(points):
    acc = (1, 1, 0)
    for p in points:
        acc = _jac_add(acc, _to_jac(p))
    return _from_jac(acc)


def g1_in_subgroup(pt) -> bool:
    if pt is None:
        return True
    return g1_mul_jac_raw(pt, R)[2] == 0


def g1_mul_jac_raw(pt, k: int):
    """k·pt without reducing k mod r (needed for the subgroup check)."""
    acc = (1, 1, 0)
    base = (pt[0], pt[1], 1)
    for bit in bin(k)[2:]:
        acc = _jac_dbl(*acc)
        if bit == "1":
            acc = _jac_add(acc, base)
    return acc


# ----------------------------------------------------------------------------
# Wire formats
# ----------------------------------------------------------------------------
def fp_lexicographically_largest(y: int) -> bool:
    return y > (P - 1) // 2


def g1_compress(pt) -> bytes:
    """G1Affine.Bytes(): 48-byte compressed ZCash form."""
    if pt is None:
        return bytes([M_COMPRESSED_INF]) + bytes(47)
    x, y = pt
    out = bytearray(x.to_bytes(48, "big"))
    out[0] |= M_COMPRESSED_LARGEST if fp_lexicographically_largest(y) else M_COMPRESSED_SMALLEST
    return bytes(out)


def fp_sqrt(a: int):
    """p ≡ 3 (mod 4): candidate a^((p+1)/4); None if a is a non-residue."""
    a %= P
    s = pow(a, (P + 1) // 4, P)
    return s if s * s % P == a else None


def g1_set_bytes(buf: bytes, subgroup_check: bool = True):
    """G1Affine.SetBytes(buf) → (point, bytes_consumed).

    Accepts compressed (48) or uncompressed (96) encodings, rejects the three
    invalid flag patterns, non-canonical x/y (>= p), non-zero padding on
    infinity, x not on curve, and (by default) points outside the r-torsion.
    """
    if len(buf) < G1_COMPRESSED:
        raise DecodeError("short buffer")
    m = buf[0] & M_MASK
    if m in (0b111 << 5, 0b011 << 5, 0b001 << 5):
        raise DecodeError("invalid encoding")
    if m in (M_UNCOMPRESSED, M_UNCOMPRESSED_INF):
        if len(buf) < G1_UNCOMPRESSED:
            raise DecodeError("short buffer")
        if m == M_UNCOMPRESSED_INF:
            if (buf[0] & ~M_MASK & 0xFF) or any(buf[1:G1_UNCOMPRESSED]):
                raise DecodeError("invalid infinity encoding")
            return None, G1_UNCOMPRESSED
        x = int.from_bytes(buf[:48], "big")
        y = int.from_bytes(buf[48:96], "big")
        if x >= P or y >= P:
            raise DecodeError("non-canonical coordinate")
        pt = (x, y)
        # gnark checks subgroup membership only (which implies on-curve for
        # the r-torsion test it uses); we require both.
        if not is_on_curve(pt) or (subgroup_check and not g1_in_subgroup(pt)):
            raise DecodeError("invalid point: subgroup check failed")
        return pt, G1_UNCOMPRESSED
    if m == M_COMPRESSED_INF:
        if (buf[0] & ~M_MASK & 0xFF) or any(buf[1:G1_COMPRESSED]):
            raise DecodeError("invalid infinity encoding")
        return None, G1_COMPRESSED
    xb = bytearray(buf[:48])
    xb[0] &= ~M_MASK & 0xFF
    x = int.from_bytes(xb, "big")
    if x >= P:
        raise DecodeError("non-canonical coordinate")
    y = fp_sqrt((x * x * x + B_COEFF) % P)
    if y is None:
        raise DecodeError("invalid compressed coordinate: square root doesn't exist")
    if fp_lexicographically_largest(y) != (m == M_COMPRESSED_LARGEST):
        y = P - y
    pt = (x, y)
    if subgroup_check and not g1_in_subgroup(pt):
        raise DecodeError("invalid point: subgroup check failed")
    return pt, G1_COMPRESSED


def g1_decompress(buf: bytes, subgroup_check: bool = True):
    pt, used = g1_set_bytes(buf, subgroup_check)
    return pt


# Optional accelerated batch decompressor (set by oracle.cbackend): takes n*48
# bytes of *compressed* encodings, returns (points, status codes).
_decompress_batch = None


def set_decompress_batch(fn) -> None:
    global _decompress_batch
    _decompress_batch = fn


_REASONS = {1: "invalid encoding", 2: "non-canonical coordinate",
            3: "invalid compressed coordinate: square root doesn't exist",
            4: "invalid point: subgroup check failed", 5: "invalid infinity encoding"}


def g1_decompress_many(encs):
    """SetBytes on a list of 48-byte compressed encodings (whisk/types.go:86-95);
    raises DecodeError for the first bad one."""
    if _decompress_batch is None:
        return [g1_decompress(e) for e in encs]
    for e in encs:
        if len(e) != G1_COMPRESSED:
            raise DecodeError("short buffer")
    pts, st = _decompress_batch(b"".join(encs))
    for code in st:
        if code:
            raise DecodeError(_REASONS.get(code, "decode error"))
    return pts


class Encoder:
    """bls12381.NewEncoder(w): compressed points, 32-byte BE scalars, uint32-BE
    slice-length prefix (assumed, unpinned — SURVEY.md §8c-i)."""

    def __init__(self):
        self.buf = bytearray()

    def point(self, pt):
        self.buf += g1_compress(pt)

    def points(self, pts):
        self.buf += len(pts).to_bytes(4, "big")
        for p in pts:
            self.buf += g1_compress(p)

    def scalar(self, s: int):
        self.buf += fr_to_bytes(s)

    def bytes(self) -> bytes:
        return bytes(self.buf)


class Decoder:
    """bls12381.NewDecoder(r) with default subgroup checks."""

    def __init__(self, data: bytes):
        self.data = bytes(data)
        self.pos = 0

    def _take(self, n: int) -> bytes:
        if self.pos + n > len(self.data):
            raise DecodeError("unexpected EOF")
        b = self.data[self.pos:self.pos + n]
        self.pos += n
        return b

    def point(self):
        head = self._take(G1_COMPRESSED)
        m = head[0] & M_MASK
        if m in (0b111 << 5, 0b011 << 5, 0b001 << 5):
            raise DecodeError("invalid encoding")
        if m in (M_UNCOMPRESSED, M_UNCOMPRESSED_INF):
            head = head + self._take(G1_COMPRESSED)
        elif _decompress_batch is not None:
            return g1_decompress_many([head])[0]
        pt, _ = g1_set_bytes(head)
        return pt

    def points(self):
        n = int.from_bytes(self._take(4), "big")
        if n * G1_COMPRESSED > len(self.data) - self.pos:
            raise DecodeError("unexpected EOF")
        if _decompress_batch is not None:
            # gnark decodes slice elements first and validates them afterwards;
            # either way any bad element fails the whole Decode.
            heads = []
            for _ in range(n):
                h = self._take(G1_COMPRESSED)
                m = h[0] & M_MASK
                if m in (0b111 << 5, 0b011 << 5, 0b001 << 5):
                    raise DecodeError("invalid encoding")
                if m in (M_UNCOMPRESSED, M_UNCOMPRESSED_INF):
                    pt, _ = g1_set_bytes(h + self._take(G1_COMPRESSED))
                    heads.append(pt)
                else:
                    heads.append(h)
            comp = [h for h in heads if isinstance(h, (bytes, bytearray))]
            dec = iter(g1_decompress_many(comp))
            return [next(dec) if isinstance(h, (bytes, bytearray)) else h for h in heads]
        return [self.point() for _ in range(n)]

    def scalar(self) -> int:
        return fr_from_bytes_canonical(self._take(FR_BYTES))
