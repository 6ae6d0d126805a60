// Kernel A/B microbenchmarks (standalone; includes the product's kernels directly):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_bin/kbench tools/kbench.cu
//   nvcc ... -DCDL_NO_DEDICATED_SQR -o tools/_bin/kbench_nosqr tools/kbench.cu
// Prints the time of dependent product / squaring chains and of the elementwise scalar-multiplication
// kernel with 1 and 4 points per thread (per-thread batch inversion), distinct and shared scalars.
#include <cstdio>
#include <vector>
#include "../go-curdleproofs_b200/csrc/k_elem.cu"

using namespace cdl;

__global__ void k_chain_mul(Fp* out, int iters) {
  Fp x, y;
#pragma unroll
  for (int i = 0; i < 12; i++) { x.v[i] = FP_ONE_D[i] ^ (threadIdx.x * (i + 1) & 0xffff); y.v[i] = FP_R2_D[i]; }
  x.v[11] &= 0x0fffffffu;
  for (int i = 0; i < iters; i++) { FpM::mul(x, x, y); FpM::mul(y, y, x); }
  FpM::add(x, x, y);
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
__global__ void k_chain_sqr(Fp* out, int iters) {
  Fp x, y;
#pragma unroll
  for (int i = 0; i < 12; i++) { x.v[i] = FP_ONE_D[i] ^ (threadIdx.x * (i + 1) & 0xffff); y.v[i] = FP_R2_D[i]; }
  x.v[11] &= 0x0fffffffu;
  for (int i = 0; i < iters; i++) { FpM::sqr(x, x); FpM::sqr(y, y); }
  FpM::add(x, x, y);
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
__global__ void k_chain_dbl(G1Jac* out, const G1Affine* in, int iters) {
  G1Jac p;
  jac_from_affine(p, in[blockIdx.x * blockDim.x + threadIdx.x]);
#pragma unroll 1
  for (int i = 0; i < iters; i++) jac_dbl(p, p);
  out[blockIdx.x * blockDim.x + threadIdx.x] = p;
}

template <class F>
static float timeit(F f, int reps = 3) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < reps + 1; r++) {
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

int main(int argc, char** argv) {
  int n = argc > 1 ? atoi(argv[1]) : (1 << 19);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  int sms = prop.multiProcessorCount;
#ifdef CDL_NO_DEDICATED_SQR
  printf("%s: sqr = mul(a, a)\n", prop.name);
#else
  printf("%s: dedicated squaring\n", prop.name);
#endif
  Fp* dfp;
  cudaMalloc(&dfp, sizeof(Fp) * sms * 8 * 128);
  const int iters = 2000;
  for (int wps : {8, 16}) {
    int blocks = sms * wps / 4;
    float tm = timeit([&] { k_chain_mul<<<blocks, 128>>>(dfp, iters); });
    float ts = timeit([&] { k_chain_sqr<<<blocks, 128>>>(dfp, iters); });
    double ops = 2.0 * iters * blocks * 128;
    printf("chains, %2d warps/SM: mul %.3f ms (%.2f G/s)   sqr %.3f ms (%.2f G/s)   sqr/mul time %.3f\n", wps, tm,
           ops / tm / 1e6, ts, ops / ts / 1e6, ts / tm);
  }
  // points: scalar multiples of the generator
  std::vector<uint32_t> hs((size_t)n * 8);
  uint64_t st = 88172645463325252ull;
  for (auto& w : hs) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; w = (uint32_t)(st >> 16); }
  for (int i = 0; i < n; i++) hs[(size_t)i * 8 + 7] &= 0x3fffffffu;  // < r
  Fr* ds;
  G1Affine *dP, *dO;
  cudaMalloc(&ds, (size_t)n * 32);
  cudaMalloc(&dP, (size_t)n * 96);
  cudaMalloc(&dO, (size_t)n * 96);
  cudaMemcpy(ds, hs.data(), (size_t)n * 32, cudaMemcpyHostToDevice);
  static const uint32_t GX[12] = {0xfd530c16u, 0x5cb38790u, 0x9976fff5u, 0x7817fc67u, 0x143ba1c1u, 0x154f95c7u, 0xf3d0e747u, 0xf0ae6acdu, 0x21dbf440u, 0xedce6eccu, 0x9e0bfb75u, 0x12017741u};
  static const uint32_t GY[12] = {0x0ce72271u, 0xbaac93d5u, 0x7918fd8eu, 0x8c22631au, 0x570725ceu, 0xdd595f13u, 0x50405194u, 0x51ac5829u, 0xad0059c0u, 0x0e1c8c3fu, 0x5008a26au, 0x0bbc3efcu};
  std::vector<G1Affine> hp(n);
  for (int i = 0; i < n; i++) for (int j = 0; j < 12; j++) { hp[i].x.v[j] = GX[j]; hp[i].y.v[j] = GY[j]; }
  cudaMemcpy(dP, hp.data(), (size_t)n * 96, cudaMemcpyHostToDevice);
  launch_scalar_mul(dP, ds, 1, nullptr, dO, n, 0);
  cudaMemcpy(dP, dO, (size_t)n * 96, cudaMemcpyDeviceToDevice);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("setup failed: %s\n", cudaGetErrorString(err)); return 1; }
  for (int stride : {1, 0}) {
    for (int m : {n, n / 2, n / 4, n / 16}) {
      float t1 = timeit([&] { launch_scalar_mul(dP, ds, stride, dP, dO, m, 0); });
      printf("k_scalar_mul n=%d %s scalars: %.3f ms (%.2f M/s)\n", m, stride ? "distinct" : "shared", t1, m / t1 / 1e3);
    }
  }
  G1Jac* dJ;
  cudaMalloc(&dJ, (size_t)n * 144);
  float td = timeit([&] { k_chain_dbl<<<n / 128, 128>>>(dJ, dP, 128); });
  printf("jac_dbl chain: %d points x 128 doublings %.3f ms (%.2f G dbl/s)\n", n, td, 128.0 * n / td / 1e6);
  for (int E : {1, 2, 8}) {
    float tj = timeit([&] {
      if (E == 1) k_jac_to_affine<1><<<(n + 63) / 64, 64>>>(dJ, dO, n);
      else if (E == 2) k_jac_to_affine<2><<<((n + 1) / 2 + 63) / 64, 64>>>(dJ, dO, n);
      else k_jac_to_affine<8><<<((n + 7) / 8 + 63) / 64, 64>>>(dJ, dO, n);
    });
    printf("k_jac_to_affine<%d> n=%d: %.3f ms\n", E, n, tj);
  }
  return 0;
}
