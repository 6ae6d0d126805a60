// Diagnostics: batched Fp product (K1) and integer-pipe peak microbenchmarks (K0).
#include <cuda_runtime.h>
#include "g1.cuh"
#include "launch.h"

namespace cdl {

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

// kind 3 / 4: two / three independent dependent-product chains per thread (ILP probe)
__global__ void k_peak_modmul2(Fp* out, int iters, uint32_t seed) {
  Fp x, y, u, v;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    x.v[i] = FP_ONE_D[i] ^ (threadIdx.x * (i + 1) & 0xffff);
    y.v[i] = FP_R2_D[i] ^ (seed & 0xff);
    u.v[i] = FP_ONE_D[i] ^ (threadIdx.x * (i + 3) & 0xffff);
    v.v[i] = FP_R2_D[i] ^ (seed & 0xf0);
  }
  x.v[11] &= 0x0fffffffu; y.v[11] &= 0x0fffffffu; u.v[11] &= 0x0fffffffu; v.v[11] &= 0x0fffffffu;
  for (int i = 0; i < iters; i++) {
    FpM::mul(x, x, y);
    FpM::mul(u, u, v);
    FpM::mul(y, y, x);
    FpM::mul(v, v, u);
  }
  FpM::add(x, x, y);
  FpM::add(u, u, v);
  FpM::add(x, x, u);
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

__global__ void k_peak_modmul3(Fp* out, int iters, uint32_t seed) {
  Fp x, y, u, v, p, q;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    x.v[i] = FP_ONE_D[i] ^ (threadIdx.x * (i + 1) & 0xffff);
    y.v[i] = FP_R2_D[i] ^ (seed & 0xff);
    u.v[i] = FP_ONE_D[i] ^ (threadIdx.x * (i + 3) & 0xffff);
    v.v[i] = FP_R2_D[i] ^ (seed & 0xf0);
    p.v[i] = FP_ONE_D[i] ^ (threadIdx.x * (i + 5) & 0xffff);
    q.v[i] = FP_R2_D[i] ^ (seed & 0x0f);
  }
  x.v[11] &= 0x0fffffffu; y.v[11] &= 0x0fffffffu; u.v[11] &= 0x0fffffffu; v.v[11] &= 0x0fffffffu;
  p.v[11] &= 0x0fffffffu; q.v[11] &= 0x0fffffffu;
  for (int i = 0; i < iters; i++) {
    FpM::mul(x, x, y);
    FpM::mul(u, u, v);
    FpM::mul(p, p, q);
    FpM::mul(y, y, x);
    FpM::mul(v, v, u);
    FpM::mul(q, q, p);
  }
  FpM::add(x, x, y);
  FpM::add(u, u, v);
  FpM::add(p, p, q);
  FpM::add(x, x, u);
  FpM::add(x, x, p);
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

void launch_fp_mul(const Fp* a, const Fp* b, Fp* out, int n, cudaStream_t s) {
  k_fp_mul<<<(n + 127) / 128, 128, 0, s>>>(a, b, out, n);
}
void launch_peak(int kind, void* out, int blocks, int tpb, int iters, uint32_t seed, cudaStream_t s) {
  if (kind == 0) k_peak_imad<<<blocks, tpb, 0, s>>>((uint32_t*)out, iters, seed);
  else if (kind == 1) k_peak_imad_wide<<<blocks, tpb, 0, s>>>((uint64_t*)out, iters, seed);
  else if (kind == 2) k_peak_modmul<<<blocks, tpb, 0, s>>>((Fp*)out, iters, seed);
  else if (kind == 3) k_peak_modmul2<<<blocks, tpb, 0, s>>>((Fp*)out, iters, seed);
  else k_peak_modmul3<<<blocks, tpb, 0, s>>>((Fp*)out, iters, seed);
}

}  // namespace cdl
