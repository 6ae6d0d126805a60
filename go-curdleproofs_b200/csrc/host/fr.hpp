// Host-side scalar field Fr of BLS12-381 (4 x 64-bit Montgomery limbs) — the
// part of gnark-crypto's fr.Element the reference's *host* code uses between GPU
// calls: challenge arithmetic, vector folding of cs/ds/x, inner products, powers
// of beta, verifier scalar unfolding (e.g. innerproductargument.go:80-88,
// 155-158, 223-234; grandproductargument.go:57-129; common/util.go:26-35).
// Same memory layout as fr.Element, so values upload to the device unchanged.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace cdlh {

typedef unsigned __int128 u128;

struct Fr {
  uint64_t l[4];
};

static const uint64_t FR_MOD[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull,
                                   0x73eda753299d7d48ull};
static const uint64_t FR_INV = 0xfffffffeffffffffull;  // -r^-1 mod 2^64
static const Fr FR_ONE = {{0x00000001fffffffeull, 0x5884b7fa00034802ull, 0x998c4fefecbc4ff5ull, 0x1824b159acc5056full}};
static const Fr FR_R2 = {{0xc999e990f3f29c6dull, 0x2b6cedcb87925c23ull, 0x05d314967254398full, 0x0748d9d99f59ff11ull}};
static const Fr FR_ZERO = {{0, 0, 0, 0}};

inline bool fr_is_zero(const Fr& a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }
inline bool fr_eq(const Fr& a, const Fr& b) { return memcmp(a.l, b.l, 32) == 0; }

inline bool fr_geq_mod(const uint64_t* a) {
  for (int i = 3; i >= 0; i--) {
    if (a[i] > FR_MOD[i]) return true;
    if (a[i] < FR_MOD[i]) return false;
  }
  return true;
}
inline void fr_sub_mod(uint64_t* a) {
  uint64_t b = 0;
  for (int i = 0; i < 4; i++) {
    u128 t = (u128)a[i] - FR_MOD[i] - b;
    a[i] = (uint64_t)t;
    b = (uint64_t)(t >> 64) & 1;
  }
}
inline Fr fr_add(const Fr& a, const Fr& b) {
  Fr r;
  uint64_t c = 0;
  for (int i = 0; i < 4; i++) {
    u128 t = (u128)a.l[i] + b.l[i] + c;
    r.l[i] = (uint64_t)t;
    c = (uint64_t)(t >> 64);
  }
  if (c || fr_geq_mod(r.l)) fr_sub_mod(r.l);
  return r;
}
inline Fr fr_sub(const Fr& a, const Fr& b) {
  Fr r;
  uint64_t bw = 0;
  for (int i = 0; i < 4; i++) {
    u128 t = (u128)a.l[i] - b.l[i] - bw;
    r.l[i] = (uint64_t)t;
    bw = (uint64_t)(t >> 64) & 1;
  }
  if (bw) {
    uint64_t c = 0;
    for (int i = 0; i < 4; i++) {
      u128 t = (u128)r.l[i] + FR_MOD[i] + c;
      r.l[i] = (uint64_t)t;
      c = (uint64_t)(t >> 64);
    }
  }
  return r;
}
inline Fr fr_neg(const Fr& a) { return fr_is_zero(a) ? a : fr_sub(FR_ZERO, a); }

// Montgomery product, CIOS with the multiplication and reduction rows interleaved.  The modulus
// leaves its top bit clear (0x73ed.. < 2^63), so the running value never needs a fifth word
// (the "no-carry" form gnark-crypto's generated fr.Mul uses as well).
inline Fr fr_mul(const Fr& a, const Fr& b) {
  uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0;
  for (int i = 0; i < 4; i++) {
    u128 c = (u128)a.l[0] * b.l[i] + t0;
    uint64_t lo = (uint64_t)c, A = (uint64_t)(c >> 64);
    const uint64_t m = lo * FR_INV;
    u128 d = (u128)m * FR_MOD[0] + lo;
    uint64_t C = (uint64_t)(d >> 64);
    c = (u128)a.l[1] * b.l[i] + t1 + A;
    A = (uint64_t)(c >> 64);
    d = (u128)m * FR_MOD[1] + (uint64_t)c + C;
    C = (uint64_t)(d >> 64);
    t0 = (uint64_t)d;
    c = (u128)a.l[2] * b.l[i] + t2 + A;
    A = (uint64_t)(c >> 64);
    d = (u128)m * FR_MOD[2] + (uint64_t)c + C;
    C = (uint64_t)(d >> 64);
    t1 = (uint64_t)d;
    c = (u128)a.l[3] * b.l[i] + t3 + A;
    A = (uint64_t)(c >> 64);
    d = (u128)m * FR_MOD[3] + (uint64_t)c + C;
    C = (uint64_t)(d >> 64);
    t2 = (uint64_t)d;
    t3 = C + A;
  }
  Fr r = {{t0, t1, t2, t3}};
  if (fr_geq_mod(r.l)) fr_sub_mod(r.l);
  return r;
}
inline Fr fr_sqr(const Fr& a) { return fr_mul(a, a); }

inline Fr fr_from_u64(uint64_t v) {  // fr.NewElement
  Fr c = {{v, 0, 0, 0}};
  return fr_mul(c, FR_R2);
}
inline Fr fr_from_canonical(const uint64_t* w) {
  Fr c = {{w[0], w[1], w[2], w[3]}};
  return fr_mul(c, FR_R2);
}
inline void fr_to_canonical(uint64_t* w, const Fr& a) {
  Fr one = {{1, 0, 0, 0}};
  Fr c = fr_mul(a, one);
  memcpy(w, c.l, 32);
}
// fr.Element.Bytes(): 32-byte big-endian canonical
inline void fr_to_bytes_be(uint8_t* out, const Fr& a) {
  uint64_t w[4];
  fr_to_canonical(w, a);
  for (int i = 0; i < 4; i++)
    for (int k = 0; k < 8; k++) out[8 * i + k] = (uint8_t)(w[3 - i] >> (56 - 8 * k));
}
// fr.Element.SetBytesCanonical: false when the value is >= r
inline bool fr_from_bytes_be_canonical(Fr& r, const uint8_t* in) {
  uint64_t w[4];
  for (int i = 0; i < 4; i++) {
    uint64_t v = 0;
    for (int k = 0; k < 8; k++) v = (v << 8) | in[8 * i + k];
    w[3 - i] = v;
  }
  if (fr_geq_mod(w)) return false;
  r = fr_from_canonical(w);
  return true;
}
inline Fr fr_pow_u64(const Fr& a, uint64_t e) {  // fr.Element.Exp with a small exponent
  Fr acc = FR_ONE, base = a;
  while (e) {
    if (e & 1) acc = fr_mul(acc, base);
    base = fr_sqr(base);
    e >>= 1;
  }
  return acc;
}
// fr.Element.Inverse; Inverse(0) = 0.  Binary extended Euclid on the raw limbs (about 500
// shift / subtract steps on 4 words instead of ~400 Montgomery products): for the Montgomery
// representation A = a*R it yields A^-1 = a^-1 * R^-1, and one product with R^3 gives a^-1 * R.
inline bool fr_raw_geq(const uint64_t* a, const uint64_t* b) {
  for (int i = 3; i >= 0; i--) {
    if (a[i] > b[i]) return true;
    if (a[i] < b[i]) return false;
  }
  return true;
}
inline void fr_raw_sub(uint64_t* a, const uint64_t* b) {  // a -= b (a >= b)
  uint64_t bw = 0;
  for (int i = 0; i < 4; i++) {
    u128 t = (u128)a[i] - b[i] - bw;
    a[i] = (uint64_t)t;
    bw = (uint64_t)(t >> 64) & 1;
  }
}
inline void fr_raw_half_mod(uint64_t* x) {  // x = x / 2 mod r  (x < r)
  uint64_t c = 0;
  if (x[0] & 1) {
    for (int i = 0; i < 4; i++) {
      u128 t = (u128)x[i] + FR_MOD[i] + c;
      x[i] = (uint64_t)t;
      c = (uint64_t)(t >> 64);
    }
  }
  for (int i = 0; i < 3; i++) x[i] = (x[i] >> 1) | (x[i + 1] << 63);
  x[3] = (x[3] >> 1) | (c << 63);
}
inline void fr_raw_shr1(uint64_t* x) {
  for (int i = 0; i < 3; i++) x[i] = (x[i] >> 1) | (x[i + 1] << 63);
  x[3] >>= 1;
}
inline void fr_raw_sub_mod(uint64_t* a, const uint64_t* b) {  // a = a - b mod r  (a, b < r)
  if (fr_raw_geq(a, b)) {
    fr_raw_sub(a, b);
  } else {
    uint64_t t[4] = {FR_MOD[0], FR_MOD[1], FR_MOD[2], FR_MOD[3]};
    fr_raw_sub(t, b);  // r - b
    uint64_t c = 0;
    for (int i = 0; i < 4; i++) {
      u128 s = (u128)a[i] + t[i] + c;
      a[i] = (uint64_t)s;
      c = (uint64_t)(s >> 64);
    }
  }
}
inline Fr fr_inv(const Fr& a) {
  if (fr_is_zero(a)) return FR_ZERO;
  static const Fr R3 = fr_mul(FR_R2, FR_R2);  // R^2 * R^2 * R^-1
  uint64_t u[4] = {a.l[0], a.l[1], a.l[2], a.l[3]};
  uint64_t v[4] = {FR_MOD[0], FR_MOD[1], FR_MOD[2], FR_MOD[3]};
  uint64_t x1[4] = {1, 0, 0, 0}, x2[4] = {0, 0, 0, 0};
  auto is_one = [](const uint64_t* w) { return w[0] == 1 && (w[1] | w[2] | w[3]) == 0; };
  while (!is_one(u) && !is_one(v)) {
    while (!(u[0] & 1)) { fr_raw_shr1(u); fr_raw_half_mod(x1); }
    while (!(v[0] & 1)) { fr_raw_shr1(v); fr_raw_half_mod(x2); }
    if (fr_raw_geq(u, v)) { fr_raw_sub(u, v); fr_raw_sub_mod(x1, x2); }
    else { fr_raw_sub(v, u); fr_raw_sub_mod(x2, x1); }
  }
  const uint64_t* x = is_one(u) ? x1 : x2;
  Fr raw = {{x[0], x[1], x[2], x[3]}};
  return fr_mul(raw, R3);
}
// fr.BatchInvert (Montgomery's trick: one inversion, 3 products per element); zeros stay zero
inline std::vector<Fr> fr_batch_inv(const std::vector<Fr>& v) {
  const size_t n = v.size();
  std::vector<Fr> out(n);
  Fr acc = FR_ONE;
  for (size_t i = 0; i < n; i++) {
    out[i] = acc;
    if (!fr_is_zero(v[i])) acc = fr_mul(acc, v[i]);
  }
  acc = fr_inv(acc);
  for (size_t i = n; i-- > 0;) {
    if (fr_is_zero(v[i])) { out[i] = FR_ZERO; continue; }
    Fr t = fr_mul(acc, out[i]);
    acc = fr_mul(acc, v[i]);
    out[i] = t;
  }
  return out;
}
// common.IPA (common/util.go:26-35)
inline Fr fr_inner(const Fr* a, const Fr* b, size_t n) {
  Fr acc = FR_ZERO;
  for (size_t i = 0; i < n; i++) acc = fr_add(acc, fr_mul(a[i], b[i]));
  return acc;
}

}  // namespace cdlh
