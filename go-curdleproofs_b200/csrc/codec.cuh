// Device-side ZCash/gnark point codec helpers (G1Affine.Bytes / SetBytes).
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
  if (!g1_in_subgroup_endo(p)) return 4;
  return 0;
}

}  // namespace cdl
