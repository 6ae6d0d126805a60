// BLS12-381 base field Fp (381 bit, 12 x 32-bit limbs) and scalar field Fr
// (255 bit, 8 x 32-bit limbs) parameters for the Montgomery core in mont.cuh.
// Replaces gnark-crypto's fp.Element / fr.Element arithmetic underneath every
// row of SURVEY.md §8a (the reference reaches it through go.mod:6).
#pragma once
#include "mont.cuh"

namespace cdl {

#define CDL_FP_MOD  {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u, 0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau}
#define CDL_FP_ONE  {0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu, 0x53c758bau, 0x5f489857u, 0x70525745u, 0x77ce5853u, 0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u}
#define CDL_FP_R2   {0x1c341746u, 0xf4df1f34u, 0x09d104f1u, 0x0a76e6a6u, 0x4c95b6d5u, 0x8de5476cu, 0x939d83c0u, 0x67eb88a9u, 0xb519952du, 0x9a793e85u, 0x92cae3aau, 0x11988fe5u}
#define CDL_FR_MOD  {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u}
#define CDL_FR_ONE  {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau, 0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u}
#define CDL_FR_R2   {0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu, 0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u}

static const uint32_t FP_MOD_H[12] = CDL_FP_MOD;
static const uint32_t FP_ONE_H[12] = CDL_FP_ONE;
static const uint32_t FP_R2_H[12] = CDL_FP_R2;
static const uint32_t FR_MOD_H[8] = CDL_FR_MOD;
static const uint32_t FR_ONE_H[8] = CDL_FR_ONE;
static const uint32_t FR_R2_H[8] = CDL_FR_R2;
#if defined(__CUDACC__)
static __device__ __constant__ uint32_t FP_MOD_D[12] = CDL_FP_MOD;
static __device__ __constant__ uint32_t FP_ONE_D[12] = CDL_FP_ONE;
static __device__ __constant__ uint32_t FP_R2_D[12] = CDL_FP_R2;
static __device__ __constant__ uint32_t FR_MOD_D[8] = CDL_FR_MOD;
static __device__ __constant__ uint32_t FR_ONE_D[8] = CDL_FR_ONE;
static __device__ __constant__ uint32_t FR_R2_D[8] = CDL_FR_R2;
#endif

#if defined(__CUDA_ARCH__)
#define CDL_SEL(dev, host) dev
#else
#define CDL_SEL(dev, host) host
#endif

struct FpParams {
  static constexpr int N = 12;
  static constexpr uint32_t M0 = 0xfffcfffdu;  // -p^-1 mod 2^32
  static CDL_HD uint32_t mod(int i) { return CDL_SEL(FP_MOD_D, FP_MOD_H)[i]; }
  static CDL_HD uint32_t one(int i) { return CDL_SEL(FP_ONE_D, FP_ONE_H)[i]; }
  static CDL_HD uint32_t r2(int i) { return CDL_SEL(FP_R2_D, FP_R2_H)[i]; }
};

struct FrParams {
  static constexpr int N = 8;
  static constexpr uint32_t M0 = 0xffffffffu;  // -r^-1 mod 2^32
  static CDL_HD uint32_t mod(int i) { return CDL_SEL(FR_MOD_D, FR_MOD_H)[i]; }
  static CDL_HD uint32_t one(int i) { return CDL_SEL(FR_ONE_D, FR_ONE_H)[i]; }
  static CDL_HD uint32_t r2(int i) { return CDL_SEL(FR_R2_D, FR_R2_H)[i]; }
};

using FpM = Mont<FpParams>;
using FrM = Mont<FrParams>;
using Fp = FpM::El;  // 48 bytes == gnark fp.Element
using Fr = FrM::El;  // 32 bytes == gnark fr.Element

// public exponents, little-endian 32-bit words
#define CDL_FP_PM2   {0xffffaaa9u, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u, 0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau}
// (p+1)/4
#define CDL_FP_SQRT  {0xffffeaabu, 0xee7fbfffu, 0xac54ffffu, 0x07aaffffu, 0x3dac3d89u, 0xd9cc34a8u, 0x3ce144afu, 0xd91dd2e1u, 0x90d2eb35u, 0x92c6e9edu, 0x8e5ff9a6u, 0x0680447au}
// (p-1)/2
#define CDL_FP_HALF  {0xffffd555u, 0xdcff7fffu, 0x58a9ffffu, 0x0f55ffffu, 0x7b587b12u, 0xb3986950u, 0x79c2895fu, 0xb23ba5c2u, 0x21a5d66bu, 0x258dd3dbu, 0x1cbff34du, 0x0d0088f5u}
// Montgomery form of curve constant b = 4 and of beta (cube root of unity with
// phi(x,y) = (beta*x, y) = lambda*(x,y), lambda = z^2 - 1)
#define CDL_FP_B     {0x000cfff3u, 0xaa270000u, 0xfc34000au, 0x53cc0032u, 0x6b0a807fu, 0x478fe97au, 0xe6ba24d7u, 0xb1d37ebeu, 0xbf78ab2fu, 0x8ec9733bu, 0x3d83de7eu, 0x09d64551u}
#define CDL_FP_BETA  {0x8671f071u, 0xcd03c9e4u, 0x1fcda5d2u, 0x5dab2246u, 0xd3851b95u, 0x587042afu, 0x01bacb9eu, 0x8eb60ebeu, 0x83d050d2u, 0x03f97d6eu, 0x54638741u, 0x18f02065u}

static const uint32_t FP_PM2_H[12] = CDL_FP_PM2;
static const uint32_t FP_SQRT_H[12] = CDL_FP_SQRT;
static const uint32_t FP_HALF_H[12] = CDL_FP_HALF;
static const uint32_t FP_B_H[12] = CDL_FP_B;
static const uint32_t FP_BETA_H[12] = CDL_FP_BETA;
#if defined(__CUDACC__)
static __device__ __constant__ uint32_t FP_PM2_D[12] = CDL_FP_PM2;
static __device__ __constant__ uint32_t FP_SQRT_D[12] = CDL_FP_SQRT;
static __device__ __constant__ uint32_t FP_HALF_D[12] = CDL_FP_HALF;
static __device__ __constant__ uint32_t FP_B_D[12] = CDL_FP_B;
static __device__ __constant__ uint32_t FP_BETA_D[12] = CDL_FP_BETA;
#endif

CDL_FN void fp_inv_fermat(Fp& r, const Fp& a) {  // a^(p-2); inv(0) = 0
  FpM::pow_words<12>(r, a, CDL_SEL(FP_PM2_D, FP_PM2_H));
}

// r = a^-1, inv(0) = 0 (Montgomery form in and out).  Binary extended Euclid in a branch-free,
// fixed-length form: the invariants u = x1*A, v = x2*A (mod p) hold for the raw limbs A = a*R; a
// step makes u even (swapping so that u >= v and subtracting when it is odd) and halves it, so
// bitlen(u) + bitlen(v) drops by at least one per step and 768 steps always end with u = 0,
// v = 1, x2 = A^-1.  About 180 add/shift/select instructions per step instead of the ~480
// Montgomery products of Fermat's a^(p-2): a third of the latency of every normalisation
// (one per scalar multiplication, per MSM result, per Horner chain) and half its issue slots.
CDL_FN void fp_inv(Fp& r, const Fp& a) {
  constexpr int N = 12;
  uint32_t u[N], v[N], x1[N], x2[N];
#pragma unroll
  for (int i = 0; i < N; i++) {
    u[i] = a.v[i];
    v[i] = FpParams::mod(i);
    x1[i] = i == 0 ? 1u : 0u;
    x2[i] = 0u;
  }
#pragma unroll 1
  for (int it = 0; it < 768; it++) {
    const uint32_t odd = 0u - (u[0] & 1u);
    uint32_t lt;
    {
      CC c;
      (void)sub_cc(c, u[0], v[0]);
#pragma unroll
      for (int i = 1; i < N; i++) (void)subc_cc(c, u[i], v[i]);
      lt = subc(c, 0, 0);  // all ones when u < v
    }
    const uint32_t sw = odd & lt;
#pragma unroll
    for (int i = 0; i < N; i++) {
      uint32_t t = (u[i] ^ v[i]) & sw;
      u[i] ^= t;
      v[i] ^= t;
      uint32_t s = (x1[i] ^ x2[i]) & sw;
      x1[i] ^= s;
      x2[i] ^= s;
    }
    {  // u -= v when u is odd (now u >= v)
      CC c;
      u[0] = sub_cc(c, u[0], v[0] & odd);
#pragma unroll
      for (int i = 1; i < N - 1; i++) u[i] = subc_cc(c, u[i], v[i] & odd);
      u[N - 1] = subc(c, u[N - 1], v[N - 1] & odd);
    }
    {  // x1 = x1 - x2 (mod p) when u was odd
      CC c;
      x1[0] = sub_cc(c, x1[0], x2[0] & odd);
#pragma unroll
      for (int i = 1; i < N; i++) x1[i] = subc_cc(c, x1[i], x2[i] & odd);
      const uint32_t borrow = subc(c, 0, 0);
      CC c2;
      x1[0] = add_cc(c2, x1[0], FpParams::mod(0) & borrow);
#pragma unroll
      for (int i = 1; i < N - 1; i++) x1[i] = addc_cc(c2, x1[i], FpParams::mod(i) & borrow);
      x1[N - 1] = addc(c2, x1[N - 1], FpParams::mod(N - 1) & borrow);
    }
#pragma unroll
    for (int i = 0; i < N - 1; i++) u[i] = (u[i] >> 1) | (u[i + 1] << 31);
    u[N - 1] >>= 1;
    {  // x1 = x1 / 2 (mod p): add p first when odd (x1 + p < 2^382, no carry out of the top limb)
      const uint32_t xo = 0u - (x1[0] & 1u);
      CC c;
      x1[0] = add_cc(c, x1[0], FpParams::mod(0) & xo);
#pragma unroll
      for (int i = 1; i < N - 1; i++) x1[i] = addc_cc(c, x1[i], FpParams::mod(i) & xo);
      x1[N - 1] = addc(c, x1[N - 1], FpParams::mod(N - 1) & xo);
#pragma unroll
      for (int i = 0; i < N - 1; i++) x1[i] = (x1[i] >> 1) | (x1[i + 1] << 31);
      x1[N - 1] >>= 1;
    }
  }
  // x2 = (a*R)^-1 = a^-1 * R^-1 as a plain residue; times R^3 in the Montgomery product gives a^-1 * R
  Fp raw, r2, r3;
#pragma unroll
  for (int i = 0; i < N; i++) {
    raw.v[i] = x2[i];
    r2.v[i] = FpParams::r2(i);
  }
  FpM::mul(r3, r2, r2);
  FpM::mul(r, raw, r3);
}

// candidate square root a^((p+1)/4) (p = 3 mod 4); returns whether it squares to a
CDL_FN bool fp_sqrt(Fp& r, const Fp& a) {
  Fp s, t;
  FpM::pow_words<12>(s, a, CDL_SEL(FP_SQRT_D, FP_SQRT_H));
  FpM::sqr(t, s);
  r = s;
  return FpM::eq(t, a);
}

CDL_HD void fp_set_b(Fp& r) {
#pragma unroll
  for (int i = 0; i < 12; i++) r.v[i] = CDL_SEL(FP_B_D, FP_B_H)[i];
}

CDL_HD void fp_set_beta(Fp& r) {
#pragma unroll
  for (int i = 0; i < 12; i++) r.v[i] = CDL_SEL(FP_BETA_D, FP_BETA_H)[i];
}

// gnark fp.Element.LexicographicallyLargest: canonical(a) > (p-1)/2
CDL_FN bool fp_lex_largest(const Fp& a_mont) {
  Fp c;
  FpM::from_mont(c, a_mont);
  CC cc;  // (p-1)/2 - c borrows  <=>  c > (p-1)/2
  (void)sub_cc(cc, CDL_SEL(FP_HALF_D, FP_HALF_H)[0], c.v[0]);
#pragma unroll
  for (int i = 1; i < 12; i++) (void)subc_cc(cc, CDL_SEL(FP_HALF_D, FP_HALF_H)[i], c.v[i]);
  return subc(cc, 0, 0) != 0;
}

// canonical limbs < p ?
CDL_HD bool fp_is_canonical(const Fp& c) {
  CC cc;
  (void)sub_cc(cc, c.v[0], FpParams::mod(0));
#pragma unroll
  for (int i = 1; i < 12; i++) (void)subc_cc(cc, c.v[i], FpParams::mod(i));
  return subc(cc, 0, 0) != 0;
}

}  // namespace cdl
