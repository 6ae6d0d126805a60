// GLV decomposition for BLS12-381 G1.
//
// phi(x, y) = (beta*x, y) acts on the r-torsion as multiplication by
// lambda = z^2 - 1 (lambda^2 + lambda + 1 = r).  A scalar k < r is first folded to
// N = min(k, r-k) < 2^254 (negating the point), then split as
//     N = k1 + k2*lambda,  k2 = round(N / lambda),  |k1| <= lambda/2,
// so that |k1|, k2 < 2^127: 32 signed 4-bit windows instead of 64, half the
// doublings of every scalar multiplication and MSM window chain.
// k * P = s0 * ( s1*|k1| * P + k2 * phi(P) ).
// The quotient is an exact Barrett division: with m = ceil(2^384 / lambda),
// floor(X*m / 2^384) == floor(X / lambda) for every X < 2^255.
#pragma once
#include "fields.cuh"

namespace cdl {

struct Glv {
  uint32_t k1[4];  // |k1|
  uint32_t k2[4];
  bool neg1;       // sign of the P part      (s0 * s1 < 0)
  bool neg2;       // sign of the phi(P) part (s0 < 0)
};

CDL_HD void glv_decompose(Glv& g, const uint32_t* k /* canonical, < r */) {
  const uint32_t M[9] = {0x896c72deu, 0xda5e4f8du, 0x268bf7a3u, 0x389f49a7u, 0xf6cfee30u,
                         0x63f6e522u, 0xe01faaddu, 0x7c6becf1u, 0x00000001u};
  const uint32_t LAM[4] = {0xffffffffu, 0x00000000u, 0x0001a402u, 0xac45a401u};
  const uint32_t LAMHALF[4] = {0x7fffffffu, 0x00000000u, 0x8000d201u, 0x5622d200u};
  const uint32_t RHALF[8] = {0x80000000u, 0x7fffffffu, 0x7fff2dffu, 0xa9ded201u,
                             0x04d0ec02u, 0x199cec04u, 0x94cebea4u, 0x39f6d3a9u};
  // s0: k > (r-1)/2 ?
  bool s0 = false;
  {
    uint64_t b = 0;
    for (int i = 0; i < 8; i++) {
      uint64_t t = (uint64_t)RHALF[i] - k[i] - b;
      b = (t >> 63) & 1;
    }
    s0 = b != 0;
  }
  uint32_t N[8];
  if (s0) {
    uint64_t b = 0;
    for (int i = 0; i < 8; i++) {
      uint64_t t = (uint64_t)FrParams::mod(i) - k[i] - b;
      N[i] = (uint32_t)t;
      b = (t >> 63) & 1;
    }
  } else {
    for (int i = 0; i < 8; i++) N[i] = k[i];
  }
  // X = N + floor(lambda/2)
  uint32_t X[8];
  {
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
      uint64_t t = (uint64_t)N[i] + (i < 4 ? LAMHALF[i] : 0u) + c;
      X[i] = (uint32_t)t;
      c = t >> 32;
    }
  }
  // q = (X * M) >> 384 : schoolbook 8 x 9 limbs, keep limbs 12..16
  uint32_t prod[17];
  for (int i = 0; i < 17; i++) prod[i] = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t c = 0;
    for (int j = 0; j < 9; j++) {
      uint64_t t = (uint64_t)X[i] * M[j] + prod[i + j] + c;
      prod[i + j] = (uint32_t)t;
      c = t >> 32;
    }
    prod[i + 9] = (uint32_t)c;
  }
  uint32_t q[4] = {prod[12], prod[13], prod[14], prod[15]};
  // k1 = N - q*lambda  (256-bit two's complement)
  uint32_t ql[8];
  for (int i = 0; i < 8; i++) ql[i] = 0;
  for (int i = 0; i < 4; i++) {
    uint64_t c = 0;
    for (int j = 0; j < 4; j++) {
      uint64_t t = (uint64_t)q[i] * LAM[j] + ql[i + j] + c;
      ql[i + j] = (uint32_t)t;
      c = t >> 32;
    }
    ql[i + 4] = (uint32_t)c;
  }
  uint32_t d[8];
  {
    uint64_t b = 0;
    for (int i = 0; i < 8; i++) {
      uint64_t t = (uint64_t)N[i] - ql[i] - b;
      d[i] = (uint32_t)t;
      b = (t >> 63) & 1;
    }
  }
  bool s1 = (d[7] >> 31) != 0;
  if (s1) {  // negate
    uint64_t c = 1;
    for (int i = 0; i < 8; i++) {
      uint64_t t = (uint64_t)(~d[i]) + c;
      d[i] = (uint32_t)t;
      c = t >> 32;
    }
  }
  for (int i = 0; i < 4; i++) { g.k1[i] = d[i]; g.k2[i] = q[i]; }
  g.neg1 = s0 != s1;
  g.neg2 = s0;
}

// signed 4-bit digits of a 127-bit magnitude: 32 digits in [-8, 8]
CDL_HD void recode_w4_128(int8_t* digits, const uint32_t* k) {
  uint32_t carry = 0;
  for (int i = 0; i < 32; i++) {
    uint32_t d = ((k[i >> 3] >> ((i & 7) * 4)) & 15u) + carry;
    carry = d > 8u;
    digits[i] = (int8_t)((int)d - (int)(carry << 4));
  }
}

// biased form for carry-free digit lookup: kp = k + 0x0888..8 (31 eights);
// digit(w) = nibble(w) - 8 for w < 31, raw nibble for w = 31 (<= 8 since k < 2^127)
CDL_HD void glv_bias(uint32_t* kp, const uint32_t* k) {
  uint64_t c = 0;
  for (int i = 0; i < 4; i++) {
    uint64_t t = (uint64_t)k[i] + (i == 3 ? 0x08888888u : 0x88888888u) + c;
    kp[i] = (uint32_t)t;
    c = t >> 32;
  }
}
CDL_HD int glv_digit(const uint32_t* kp, int w) {
  int nib = (int)((kp[w >> 3] >> ((w & 7) * 4)) & 15u);
  return w == 31 ? nib : nib - 8;
}

}  // namespace cdl
