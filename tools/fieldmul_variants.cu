// Field-multiplication formulations for sm_100a, BUILT and measured, next to the
// product's 12 x 32-bit carry-chain CIOS core (csrc/mont.cuh):
//
//   A  "carry-free" reduced radix: 14 limbs x 28 bits, every partial product a plain IMAD.WIDE.U32 with
//      a 64-bit accumulate (no carry flag anywhere), carries resolved by shifts once per column;
//      Montgomery reduction with R = 2^392, lazily reduced operands (< 2p, no final subtraction needed)
//   B  FP64: 17 limbs x 23 bits held as doubles, partial products accumulated EXACTLY by DFMA (column
//      sums stay below 2^53), quotient digits through the 2^52 magic-number trick; R = 2^391
//
// Every variant is checked against the product core on random operands (the results differ only by the
// Montgomery factor: r_std = r_A * 2^8 = r_B * 2^7 mod p) and then timed as dependent chains, exactly like
// the product's own k_peak_modmul probe.  Standalone:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_bin/fieldmul_variants tools/fieldmul_variants.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../go-curdleproofs_b200/csrc/g1.cuh"

using namespace cdl;

// ------------------------------------------------------------------ A: 14 x 28-bit limbs
constexpr int NA = 14;
__device__ __constant__ uint32_t P28[NA] = {0xfffaaabu, 0xfefffffu, 0x3ffffb9u, 0xfffeb15u, 0x6241eabu, 0xa0f6b0fu, 0xf6730d2u,
                                            0xf38512bu, 0x4774b84u, 0x4bacd76u, 0xba7b643u, 0xe69a4b1u, 0x1ea397fu, 0x1a011u};
constexpr uint32_t M28 = 0xffcfffdu;  // -p^-1 mod 2^28
constexpr uint32_t MASK28 = (1u << 28) - 1;

struct ElA { uint32_t v[NA]; };

__device__ __forceinline__ uint64_t madw(uint32_t a, uint32_t b, uint64_t c) {
  uint64_t r;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c));
  return r;
}

// r = a * b * 2^-392 mod p, limbs < 2^28 + small, value < 2p for inputs < 2p
__device__ __forceinline__ void mul_a(ElA& r, const ElA& a, const ElA& b) {
  uint64_t t[2 * NA];
#pragma unroll
  for (int k = 0; k < 2 * NA; k++) t[k] = 0;
#pragma unroll
  for (int i = 0; i < NA; i++)
#pragma unroll
    for (int j = 0; j < NA; j++) t[i + j] = madw(a.v[j], b.v[i], t[i + j]);
#pragma unroll
  for (int i = 0; i < NA; i++) {
    const uint32_t m = ((uint32_t)t[i] * M28) & MASK28;
#pragma unroll
    for (int j = 0; j < NA; j++) t[i + j] = madw(m, P28[j], t[i + j]);
    t[i + 1] += t[i] >> 28;  // t[i] is now divisible by 2^28
  }
#pragma unroll
  for (int k = 0; k < NA - 1; k++) {
    r.v[k] = (uint32_t)t[NA + k] & MASK28;
    t[NA + k + 1] += t[NA + k] >> 28;
  }
  r.v[NA - 1] = (uint32_t)t[2 * NA - 1];
}

__device__ void to_a(ElA& r, const Fp& x) {  // repack 12 x 32 -> 14 x 28
#pragma unroll
  for (int k = 0; k < NA; k++) {
    int bit = 28 * k, w = bit >> 5, s = bit & 31;
    uint64_t two = (uint64_t)x.v[w] | (w + 1 < 12 ? (uint64_t)x.v[w + 1] << 32 : 0);
    r.v[k] = (uint32_t)(two >> s) & MASK28;
  }
}
__device__ void from_a(Fp& r, const ElA& x) {  // normalise + repack (value < 2p < 2^384)
  uint32_t n[NA];
  uint32_t c = 0;
  for (int k = 0; k < NA; k++) { uint32_t v = x.v[k] + c; n[k] = v & MASK28; c = v >> 28; }
  n[NA - 1] += c << 28;
  for (int w = 0; w < 12; w++) {
    int bit = 32 * w, k = bit / 28, s = bit % 28;
    uint64_t acc = (uint64_t)n[k] >> s;
    int have = 28 - s;
    for (int q = k + 1; have < 32 && q < NA; q++) { acc |= (uint64_t)n[q] << have; have += 28; }
    r.v[w] = (uint32_t)acc;
  }
}

// ------------------------------------------------------------------ B: 17 x 23-bit limbs in doubles
constexpr int NB = 17;
__device__ __constant__ double P23[NB] = {8366763.0, 8388607.0, 8316923.0, 696319.0, 4194283.0, 4490197.0, 4041789.0, 1599824.0, 1228647.0,
                                          648970.0, 1170734.0, 6121147.0, 8086580.0, 4809588.0, 6289830.0, 587036.0, 6657.0};
constexpr uint32_t M23 = 0x7cfffdu;  // -p^-1 mod 2^23
constexpr uint32_t MASK23 = (1u << 23) - 1;
constexpr double TWO52 = 4503599627370496.0;

struct ElB { double v[NB]; };

__device__ __forceinline__ double dfma(double a, double b, double c) { return __fma_rn(a, b, c); }

// r = a * b * 2^-391 mod p; limbs are exact integers < 2^23 + small in doubles, value < 2p
__device__ __forceinline__ void mul_b(ElB& r, const ElB& a, const ElB& b) {
  double t[NB + 1];
#pragma unroll
  for (int k = 0; k <= NB; k++) t[k] = 0.0;
#pragma unroll
  for (int i = 0; i < NB; i++) {
#pragma unroll
    for (int j = 0; j < NB; j++) t[j] = dfma(a.v[j], b.v[i], t[j]);
    // quotient digit: low 23 bits of the exact integer t[0] (< 2^52) through the 2^52 offset
    const uint32_t lo = (uint32_t)__double_as_longlong(t[0] + TWO52) & MASK23;
    const uint32_t q = (lo * M23) & MASK23;
    const double qd = __longlong_as_double(0x4330000000000000ll | (long long)q) - TWO52;
#pragma unroll
    for (int j = 0; j < NB; j++) t[j] = dfma(qd, P23[j], t[j]);
    const double carry = t[0] * (1.0 / 8388608.0);  // exact: t[0] is divisible by 2^23
#pragma unroll
    for (int j = 0; j < NB; j++) t[j] = t[j + 1];
    t[0] += carry;
    t[NB] = 0.0;
  }
  // normalise the column sums (each < 2^52) to 23-bit limbs with an integer carry chain
  unsigned long long c = 0;
#pragma unroll
  for (int k = 0; k < NB; k++) {
    unsigned long long v = ((unsigned long long)__double_as_longlong(t[k] + TWO52) & 0xfffffffffffffull) + c;
    const uint32_t limb = k == NB - 1 ? (uint32_t)v : (uint32_t)v & MASK23;
    c = k == NB - 1 ? 0 : v >> 23;
    r.v[k] = __longlong_as_double(0x4330000000000000ll | (long long)limb) - TWO52;
  }
}

__device__ void to_b(ElB& r, const Fp& x) {
  for (int k = 0; k < NB; k++) {
    int bit = 23 * k, w = bit >> 5, s = bit & 31;
    uint64_t two = (w < 12 ? (uint64_t)x.v[w] : 0) | (w + 1 < 12 ? (uint64_t)x.v[w + 1] << 32 : 0);
    r.v[k] = (double)((uint32_t)(two >> s) & MASK23);
  }
}
__device__ void from_b(Fp& r, const ElB& x) {
  uint32_t n[NB];
  for (int k = 0; k < NB; k++) n[k] = (uint32_t)x.v[k];
  for (int w = 0; w < 12; w++) {
    int bit = 32 * w, k = bit / 23, s = bit % 23;
    uint64_t acc = (uint64_t)n[k] >> s;
    int have = 23 - s;
    for (int q = k + 1; have < 32 && q < NB; q++) { acc |= (uint64_t)n[q] << have; have += 23; }
    r.v[w] = (uint32_t)acc;
  }
}

// ------------------------------------------------------------------ checks
__device__ void reduce_once(Fp& x) {  // x < 2p -> x mod p
  uint32_t t[12];
  for (int i = 0; i < 12; i++) t[i] = x.v[i];
  FpM::final_sub(t);
  for (int i = 0; i < 12; i++) x.v[i] = t[i];
}

__global__ void k_check(const Fp* a, const Fp* b, int n, int* bad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp x = a[i], y = b[i], want;
  FpM::mul_inline(want, x, y);  // x*y*2^-384
  {
    ElA xa, ya, ra;
    to_a(xa, x);
    to_a(ya, y);
    mul_a(ra, xa, ya);
    mul_a(ra, ra, ya);  // chained once more: lazily reduced operand in
    Fp got, w2;
    from_a(got, ra);
    reduce_once(got);
    for (int k = 0; k < 16; k++) FpM::dbl(got, got);  // two products: 2^-784 vs 2^-768
    FpM::mul_inline(w2, want, y);
    if (!FpM::eq(got, w2)) atomicOr(bad, 1);
  }
  {
    ElB xb, yb, rb;
    to_b(xb, x);
    to_b(yb, y);
    mul_b(rb, xb, yb);
    mul_b(rb, rb, yb);
    Fp got, w2;
    from_b(got, rb);
    reduce_once(got);
    for (int k = 0; k < 14; k++) FpM::dbl(got, got);
    FpM::mul_inline(w2, want, y);
    if (!FpM::eq(got, w2)) atomicOr(bad, 2);
  }
}

// ------------------------------------------------------------------ dependent chains (throughput)
__global__ void k_chain_std(Fp* out, int iters) {
  Fp x, y;
  for (int i = 0; i < 12; i++) { x.v[i] = FP_ONE_D[i] ^ (threadIdx.x * (i + 1) & 0xffff); y.v[i] = FP_R2_D[i]; }
  x.v[11] &= 0x0fffffffu;
  for (int i = 0; i < iters; i++) { FpM::mul_inline(x, x, y); FpM::mul_inline(y, y, x); }
  FpM::add(x, x, y);
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
__global__ void k_chain_a(Fp* out, int iters) {
  Fp x0, y0;
  for (int i = 0; i < 12; i++) { x0.v[i] = FP_ONE_D[i] ^ (threadIdx.x * (i + 1) & 0xffff); y0.v[i] = FP_R2_D[i]; }
  x0.v[11] &= 0x0fffffffu;
  ElA x, y;
  to_a(x, x0);
  to_a(y, y0);
  for (int i = 0; i < iters; i++) { mul_a(x, x, y); mul_a(y, y, x); }
  from_a(x0, x);
  from_a(y0, y);
  for (int i = 0; i < 12; i++) x0.v[i] ^= y0.v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0;
}
__global__ void k_chain_b(Fp* out, int iters) {
  Fp x0, y0;
  for (int i = 0; i < 12; i++) { x0.v[i] = FP_ONE_D[i] ^ (threadIdx.x * (i + 1) & 0xffff); y0.v[i] = FP_R2_D[i]; }
  x0.v[11] &= 0x0fffffffu;
  ElB x, y;
  to_b(x, x0);
  to_b(y, y0);
  for (int i = 0; i < iters; i++) { mul_b(x, x, y); mul_b(y, y, x); }
  from_b(x0, x);
  from_b(y0, y);
  for (int i = 0; i < 12; i++) x0.v[i] ^= y0.v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0;
}
// both pipes at once: warps alternate between the IMAD core and the DFMA core
__global__ void k_chain_mixed(Fp* out, int iters) {
  Fp x0, y0;
  for (int i = 0; i < 12; i++) { x0.v[i] = FP_ONE_D[i] ^ (threadIdx.x * (i + 1) & 0xffff); y0.v[i] = FP_R2_D[i]; }
  x0.v[11] &= 0x0fffffffu;
  if ((threadIdx.x >> 5) & 1) {
    ElB x, y;
    to_b(x, x0);
    to_b(y, y0);
    for (int i = 0; i < iters; i++) { mul_b(x, x, y); mul_b(y, y, x); }
    from_b(x0, x);
    from_b(y0, y);
  } else {
    for (int i = 0; i < iters; i++) { FpM::mul_inline(x0, x0, y0); FpM::mul_inline(y0, y0, x0); }
  }
  for (int i = 0; i < 12; i++) x0.v[i] ^= y0.v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0;
}

template <class F>
static float timeit(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; r++) {
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r && ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  printf("%s, %d SMs, %.0f MHz\n", prop.name, sms, prop.clockRate / 1000.0);
  // correctness on random operands below p
  const int n = 1 << 14;
  std::vector<uint32_t> ha((size_t)n * 12), hb((size_t)n * 12);
  uint64_t st = 0x243f6a8885a308d3ull;
  auto rnd = [&] { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (uint32_t)(st >> 11); };
  for (int i = 0; i < n; i++)
    for (int j = 0; j < 12; j++) {
      ha[(size_t)i * 12 + j] = j == 11 ? rnd() & 0x0fffffffu : rnd();
      hb[(size_t)i * 12 + j] = j == 11 ? rnd() & 0x0fffffffu : rnd();
    }
  for (int j = 0; j < 12; j++) { ha[j] = 0; hb[12 + j] = 0; ha[24 + j] = j == 11 ? 0x0fffffffu : 0xffffffffu; }  // edge rows
  Fp *da, *db;
  int* dbad;
  cudaMalloc(&da, (size_t)n * 48);
  cudaMalloc(&db, (size_t)n * 48);
  cudaMalloc(&dbad, 4);
  cudaMemset(dbad, 0, 4);
  cudaMemcpy(da, ha.data(), (size_t)n * 48, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), (size_t)n * 48, cudaMemcpyHostToDevice);
  k_check<<<n / 128, 128>>>(da, db, n, dbad);
  int bad = -1;
  cudaMemcpy(&bad, dbad, 4, cudaMemcpyDeviceToHost);
  cudaError_t err = cudaDeviceSynchronize();
  printf("check vs the product core on %d random operand pairs: 28-bit carry-free %s, FP64 23-bit %s%s\n", n,
         (bad & 1) ? "MISMATCH" : "ok", (bad & 2) ? "MISMATCH" : "ok", err == cudaSuccess ? "" : " (CUDA error)");
  Fp* out;
  cudaMalloc(&out, sizeof(Fp) * sms * 16 * 128);
  const int iters = 1000;
  printf("dependent chains, 2 products per iteration, %d iterations, 128-thread CTAs:\n", iters);
  printf("%-44s %10s %10s %10s\n", "variant", "8 warps/SM", "16 w/SM", "32 w/SM");
  auto row = [&](const char* name, auto kern) {
    printf("%-44s", name);
    for (int wps : {8, 16, 32}) {
      int blocks = sms * wps / 4;
      float ms = timeit([&] { kern<<<blocks, 128>>>(out, iters); });
      printf(" %8.2f G", 2.0 * iters * blocks * 128 / ms / 1e6);
    }
    printf("   modmul/s\n");
  };
  row("12 x 32-bit carry-chain CIOS (product core)", k_chain_std);
  row("14 x 28-bit carry-free IMAD.WIDE", k_chain_a);
  row("17 x 23-bit FP64 DFMA (exact accumulation)", k_chain_b);
  row("half the warps IMAD core, half DFMA core", k_chain_mixed);
  return (bad == 0 && err == cudaSuccess) ? 0 : 1;
}
