// CUDA kernels for the G1 hot path (sm_100a).  Integer-pipe work: no tensor
// cores, no GEMM reshaping.  One thread owns one 381-bit field element chain;
// throughput comes from IMAD.WIDE issue rate x occupancy.
#pragma once
#include <cuda_runtime.h>
#include "g1.cuh"

namespace cdl {

// ---------------------------------------------------------------- codecs
// flag bits of the ZCash/gnark compressed encoding
constexpr uint32_t kFlagCompressed = 0x80, kFlagInfinity = 0x40, kFlagLargest = 0x20;

// canonical little-endian limbs -> 48 big-endian bytes
__device__ __forceinline__ void fp_store_be(uint8_t* out, const Fp& c) {
#pragma unroll
  for (int i = 0; i < 12; i++) {
    uint32_t w = c.v[11 - i];
    out[4 * i + 0] = (uint8_t)(w >> 24);
    out[4 * i + 1] = (uint8_t)(w >> 16);
    out[4 * i + 2] = (uint8_t)(w >> 8);
    out[4 * i + 3] = (uint8_t)w;
  }
}

__device__ __forceinline__ void fp_load_be(Fp& c, const uint8_t* in) {
#pragma unroll
  for (int i = 0; i < 12; i++) {
    c.v[11 - i] = ((uint32_t)in[4 * i] << 24) | ((uint32_t)in[4 * i + 1] << 16) |
                  ((uint32_t)in[4 * i + 2] << 8) | (uint32_t)in[4 * i + 3];
  }
}

// G1Affine.Bytes(): 48-byte compressed form
__device__ __forceinline__ void g1_compress_dev(uint8_t* out, const G1Affine& p) {
  if (aff_is_inf(p)) {
    out[0] = kFlagCompressed | kFlagInfinity;
    for (int i = 1; i < 48; i++) out[i] = 0;
    return;
  }
  Fp xc;
  FpM::from_mont(xc, p.x);
  fp_store_be(out, xc);
  out[0] |= fp_lex_largest(p.y) ? (kFlagCompressed | kFlagLargest) : kFlagCompressed;
}

// [r]P == infinity, r the group order (255-bit, top nibble 7: fits recode_w4)
__device__ __forceinline__ bool g1_in_subgroup_dev(const G1Affine& p) {
  uint32_t rr[8];
#pragma unroll
  for (int i = 0; i < 8; i++) rr[i] = FR_MOD_D[i];
  G1Jac t;
  jac_scalar_mul(t, p, rr);
  return jac_is_inf(t);
}

// G1Affine.SetBytes for one compressed encoding; returns 0 or a reason code
__device__ __forceinline__ uint32_t g1_decompress_dev(G1Affine& p, const uint8_t* in) {
  uint32_t flags = in[0] & 0xe0u;
  if (!(flags & kFlagCompressed)) return 1;  // uncompressed forms are not accepted on 48-byte inputs
  if (flags == 0xe0u) return 1;              // 0b111 is an invalid mask
  if (flags & kFlagInfinity) {
    uint32_t o = in[0] & 0x1fu;
    for (int i = 1; i < 48; i++) o |= in[i];
    if (o) return 5;
    aff_set_inf(p);
    return 0;
  }
  Fp xc;
  fp_load_be(xc, in);
  xc.v[11] &= 0x1fffffffu;
  if (!fp_is_canonical(xc)) return 2;
  Fp x, y2, y, b;
  FpM::to_mont(x, xc);
  FpM::sqr(y2, x);
  FpM::mul(y2, y2, x);
  fp_set_b(b);
  FpM::add(y2, y2, b);
  if (!fp_sqrt(y, y2)) return 3;
  bool want_largest = (flags & kFlagLargest) != 0;
  if (fp_lex_largest(y) != want_largest) FpM::neg(y, y);
  p.x = x;
  p.y = y;
  if (!g1_in_subgroup_dev(p)) return 4;
  return 0;
}

__global__ void k_compress(const G1Affine* __restrict__ in, uint8_t* __restrict__ out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1Affine p = in[i];
  g1_compress_dev(out + 48 * (size_t)i, p);
}

__global__ void k_decompress(const uint8_t* __restrict__ in, G1Affine* __restrict__ out,
                             uint8_t* __restrict__ status, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1Affine p;
  aff_set_inf(p);
  uint32_t st = g1_decompress_dev(p, in + 48 * (size_t)i);
  if (st) aff_set_inf(p);
  out[i] = p;
  status[i] = (uint8_t)st;
}

// ---------------------------------------------------------------- diagnostics
__global__ void k_fp_mul(const Fp* __restrict__ a, const Fp* __restrict__ b, Fp* __restrict__ out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp x = a[i], y = b[i], r;
  FpM::mul(r, x, y);
  out[i] = r;
}

// kind 0: 8 independent 32-bit IMAD chains per thread
__global__ void k_peak_imad(uint32_t* out, int iters, uint32_t seed) {
  uint32_t x0 = threadIdx.x + seed, x1 = x0 * 3, x2 = x0 * 5, x3 = x0 * 7, x4 = x0 * 11, x5 = x0 * 13,
           x6 = x0 * 17, x7 = x0 * 19;
  uint32_t m = seed | 1, c = seed * 2654435761u;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x0) : "r"(m), "r"(c));
      asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x1) : "r"(m), "r"(c));
      asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x2) : "r"(m), "r"(c));
      asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x3) : "r"(m), "r"(c));
      asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x4) : "r"(m), "r"(c));
      asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x5) : "r"(m), "r"(c));
      asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x6) : "r"(m), "r"(c));
      asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x7) : "r"(m), "r"(c));
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
}

// kind 1: 8 independent 32x32+64 -> 64 IMAD.WIDE chains per thread
__global__ void k_peak_imad_wide(uint64_t* out, int iters, uint32_t seed) {
  uint64_t x0 = threadIdx.x + seed, x1 = x0 * 3, x2 = x0 * 5, x3 = x0 * 7, x4 = x0 * 11, x5 = x0 * 13,
           x6 = x0 * 17, x7 = x0 * 19;
  uint32_t m = seed | 1, c = threadIdx.x * 2654435761u + seed;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x0) : "r"(m), "r"(c));
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x1) : "r"(m), "r"(c));
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x2) : "r"(m), "r"(c));
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x3) : "r"(m), "r"(c));
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x4) : "r"(m), "r"(c));
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x5) : "r"(m), "r"(c));
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x6) : "r"(m), "r"(c));
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x7) : "r"(m), "r"(c));
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
}

// kind 2: dependent Montgomery products, one chain per thread
__global__ void k_peak_modmul(Fp* out, int iters, uint32_t seed) {
  Fp x, y;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    x.v[i] = FP_ONE_D[i] ^ (threadIdx.x * (i + 1) & 0xffff);
    y.v[i] = FP_R2_D[i] ^ (seed & 0xff);
  }
  x.v[11] &= 0x0fffffffu;
  y.v[11] &= 0x0fffffffu;
  for (int i = 0; i < iters; i++) {
    FpM::mul(x, x, y);
    FpM::mul(y, y, x);
  }
  FpM::add(x, x, y);
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

// ---------------------------------------------------------------- elementwise
// out[i] = s[i*stride] * P[i] (+ L[i] when L != nullptr), affine result.
// stride 0: one shared scalar (Whisk rescale, IPA / SameMSM folds) — every lane
// runs the identical digit schedule, no divergence.  Scalars arrive in gnark's
// Montgomery fr.Element form and are brought to canonical form here.
__global__ void __launch_bounds__(128)
k_scalar_mul(const G1Affine* __restrict__ P, const Fr* __restrict__ s, int stride,
             const G1Affine* __restrict__ L, G1Affine* __restrict__ out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr km = s[(size_t)i * stride], k;
  FrM::from_mont(k, km);
  G1Affine p = P[i];
  G1Jac r;
  jac_scalar_mul(r, p, k.v);
  if (L != nullptr) {
    G1Affine l = L[i];
    jac_add_mixed(r, r, l);
  }
  G1Affine a;
  jac_to_affine(a, r);
  out[i] = a;
}

__global__ void __launch_bounds__(128)
k_jac_to_affine(const G1Jac* __restrict__ in, G1Affine* __restrict__ out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1Jac p = in[i];
  G1Affine a;
  jac_to_affine(a, p);
  out[i] = a;
}

}  // namespace cdl
