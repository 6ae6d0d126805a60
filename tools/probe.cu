// Integer-pipe instruction-rate probes for sm_100a (standalone; `nvcc -o tools/_bin/probe tools/probe.cu`).
// Prints warp-instructions per cycle per SM for the SASS forms a 32-bit-limb Montgomery product is made of.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)

// plain IMAD.WIDE.U32 (64-bit accumulate, no carry)
__global__ void k_wide(uint32_t* out, int iters, uint32_t seed) {
  uint64_t x[8];
  uint32_t m = seed | 1, c = threadIdx.x * 2654435761u + seed;
#pragma unroll
  for (int u = 0; u < 8; u++) x[u] = threadIdx.x * (u + 3);
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int u = 0; u < 8; u++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[u]) : "r"(m), "r"(c));
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int u = 0; u < 8; u++) s ^= x[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(s ^ (s >> 32));
}
// IMAD.WIDE.U32 with carry-out + IADD3.X into a counter
__global__ void k_wide_cout(uint32_t* out, int iters, uint32_t seed) {
  uint32_t lo[8], hi[8], cnt[8];
  uint32_t m = seed | 1, c = threadIdx.x * 2654435761u + seed;
#pragma unroll
  for (int u = 0; u < 8; u++) { lo[u] = threadIdx.x * (u + 3); hi[u] = seed * (u + 7); cnt[u] = 0; }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int u = 0; u < 8; u++)
        asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
                     : "+r"(lo[u]), "+r"(hi[u]), "+r"(cnt[u]) : "r"(m), "r"(c));
    }
  }
  uint32_t x = 0;
#pragma unroll
  for (int u = 0; u < 8; u++) x ^= lo[u] ^ hi[u] ^ cnt[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
// carry chain: IMAD.WIDE.U32.X across 8 accumulators
__global__ void k_wide_x(uint32_t* out, int iters, uint32_t seed) {
  uint32_t lo[8], hi[8];
  uint32_t m = seed | 1, c = threadIdx.x * 2654435761u + seed;
#pragma unroll
  for (int u = 0; u < 8; u++) { lo[u] = threadIdx.x * (u + 3); hi[u] = seed * (u + 7); }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
      asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[0]), "+r"(hi[0]) : "r"(m), "r"(c));
#pragma unroll
      for (int u = 1; u < 7; u++)
        asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[u]), "+r"(hi[u]) : "r"(m), "r"(c));
      asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[7]), "+r"(hi[7]) : "r"(m), "r"(c));
    }
  }
  uint32_t x = 0;
#pragma unroll
  for (int u = 0; u < 8; u++) x ^= lo[u] ^ hi[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
// ALU: IADD3 (3-input add), 8 independent chains
__global__ void k_iadd3(uint32_t* out, int iters, uint32_t seed) {
  uint32_t x[8];
  uint32_t m = seed | 1, c = threadIdx.x * 2654435761u + seed;
#pragma unroll
  for (int u = 0; u < 8; u++) x[u] = threadIdx.x * (u + 3);
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int u = 0; u < 8; u++) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(x[u]) : "r"(m), "r"(c));
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int u = 0; u < 8; u++) s ^= x[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// ALU: SHF funnel shift + LOP3
__global__ void k_shf(uint32_t* out, int iters, uint32_t seed) {
  uint32_t x[8];
  uint32_t m = seed | 1;
#pragma unroll
  for (int u = 0; u < 8; u++) x[u] = threadIdx.x * (u + 3) + seed;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int u = 0; u < 8; u++) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(x[u]) : "r"(m));
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int u = 0; u < 8; u++) s ^= x[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: one plain IMAD.WIDE + one IADD3 + one SHF per group (co-issue test)
__global__ void k_mixed(uint32_t* out, int iters, uint32_t seed) {
  uint64_t x[8];
  uint32_t y[8], z[8];
  uint32_t m = seed | 1, c = threadIdx.x * 2654435761u + seed;
#pragma unroll
  for (int u = 0; u < 8; u++) { x[u] = threadIdx.x * (u + 3); y[u] = u + seed; z[u] = u * 7 + seed; }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int u = 0; u < 8; u++) {
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[u]) : "r"(m), "r"(c));
        asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(y[u]) : "r"(m), "r"(c));
      }
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int u = 0; u < 8; u++) s ^= x[u] ^ y[u] ^ z[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(s ^ (s >> 32));
}

template <class K>
static void run(const char* name, K kernel, double inst_per_iter, int sms, double mhz) {
  uint32_t* d;
  cudaMalloc(&d, sizeof(uint32_t) * sms * 8 * 256);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 2000;
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    kernel<<<sms * 8, 256>>>(d, iters, 12345u + rep);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  double warp_inst = (double)sms * 8 * 8 * iters * inst_per_iter;  // 8 warps per block
  double cycles = best * 1e-3 * mhz * 1e6;
  printf("%-12s %8.3f ms  %6.3f warp-inst/clk/SM  (%5.1f lanes/clk/SM)\n", name, best, warp_inst / cycles / sms,
         32.0 * warp_inst / cycles / sms);
  cudaFree(d);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  double mhz = p.clockRate / 1000.0;
  printf("%s, %d SMs, %.0f MHz\n", p.name, p.multiProcessorCount, mhz);
  int sms = p.multiProcessorCount;
  run("wide", k_wide, 64, sms, mhz);
  run("wide+cout", k_wide_cout, 64, sms, mhz);       // counts the IMAD.WIDE only (the IADD3.X rides along)
  run("wide.X", k_wide_x, 64, sms, mhz);
  run("iadd3", k_iadd3, 64, sms, mhz);
  run("shf", k_shf, 64, sms, mhz);
  run("wide|iadd3", k_mixed, 64, sms, mhz);          // IMAD.WIDE count; an equal number of IADD3 co-issue
  return 0;
}
