// Round-2 instruction-rate probes for sm_100a (standalone: `nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -o tools/_bin/probe2 tools/probe2.cu`).  Round 1's tools/probe.cu fed loop-invariant operands to
// `mad.wide.u32`, which ptxas strength-reduced into IADD3 + IADD3.X, so its "plain IMAD.WIDE" and
// "wide|iadd3" rows never measured an IMAD.WIDE.  Every chain here multiplies a value that changes each
// iteration; the SASS opcode mix of every kernel is checked with `cuobjdump -sass` (profiles/r2_probe2_sass.txt).
//
// Output: warp-instructions per clock per SM and lanes per clock per SM for each instruction (mix).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define NCH 8

// IMAD.WIDE.U32 without carry: x[u] += m[u] * c[r]; m changes every outer iteration, so no product is
// loop-invariant or shared (ptxas otherwise strength-reduces or splits the accumulate into IADD3s)
__global__ void k_wide(uint32_t* out, int iters, uint32_t seed) {
  uint64_t x[NCH];
  uint32_t m[NCH], c[8];
#pragma unroll
  for (int u = 0; u < NCH; u++) { x[u] = (uint64_t)threadIdx.x * (u + 3) + seed; m[u] = seed * (u + 1) + threadIdx.x; }
#pragma unroll
  for (int r = 0; r < 8; r++) c[r] = threadIdx.x * 2654435761u + seed * (r + 1);
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int u = 0; u < NCH; u++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[u]) : "r"(m[u]), "r"(c[r]));
    }
#pragma unroll
    for (int u = 0; u < NCH; u++) m[u] += (uint32_t)x[(u + 1) % NCH];
  }
  uint64_t s = 0;
#pragma unroll
  for (int u = 0; u < NCH; u++) s ^= x[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(s ^ (s >> 32));
}
// IMAD.WIDE.U32 with RZ addend: x = lo(x) * hi(x) (both halves stay live, nothing to simplify)
__global__ void k_wide_rz(uint32_t* out, int iters, uint32_t seed) {
  uint64_t x[NCH];
#pragma unroll
  for (int u = 0; u < NCH; u++) x[u] = ((uint64_t)(threadIdx.x * (u + 3) + seed) << 32) | (seed * (u + 5) + 7u);
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int u = 0; u < NCH; u++)
        asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mul.wide.u32 %0, lo, hi;}" : "+l"(x[u]));
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int u = 0; u < NCH; u++) s ^= x[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(s ^ (s >> 32));
}
// IMAD.WIDE.U32 with a 64-bit addend written as mad.lo.cc + madc.hi (the form ptxas keeps fused)
__global__ void k_wide_acc(uint32_t* out, int iters, uint32_t seed) {
  uint32_t lo[NCH], hi[NCH], m[NCH], c[8];
#pragma unroll
  for (int u = 0; u < NCH; u++) { lo[u] = threadIdx.x * (u + 3); hi[u] = seed * (u + 7); m[u] = seed * (u + 1) + threadIdx.x; }
#pragma unroll
  for (int r = 0; r < 8; r++) c[r] = threadIdx.x * 2654435761u + seed * (r + 1);
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int u = 0; u < NCH; u++)
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[u]), "+r"(hi[u]) : "r"(m[u]), "r"(c[r]));
    }
#pragma unroll
    for (int u = 0; u < NCH; u++) m[u] += i;
  }
  uint32_t x = 0;
#pragma unroll
  for (int u = 0; u < NCH; u++) x ^= lo[u] ^ hi[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
// IMAD (32-bit low): x = x * c + m
__global__ void k_imad(uint32_t* out, int iters, uint32_t seed) {
  uint32_t x[NCH];
  uint32_t c = threadIdx.x * 2654435761u + seed, m = seed | 1;
#pragma unroll
  for (int u = 0; u < NCH; u++) x[u] = threadIdx.x * (u + 3) + seed;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int u = 0; u < NCH; u++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[u]) : "r"(c), "r"(m));
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int u = 0; u < NCH; u++) s ^= x[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// IMAD.HI: x = hi(x * c) + m
__global__ void k_imad_hi(uint32_t* out, int iters, uint32_t seed) {
  uint32_t x[NCH];
  uint32_t c = threadIdx.x * 2654435761u + seed, m = seed | 1;
#pragma unroll
  for (int u = 0; u < NCH; u++) x[u] = threadIdx.x * (u + 3) + seed;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int u = 0; u < NCH; u++) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[u]) : "r"(c), "r"(m));
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int u = 0; u < NCH; u++) s ^= x[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// carry chain IMAD.WIDE.U32.X across 8 accumulators (as in the 32-bit-limb CIOS rows)
__global__ void k_wide_x(uint32_t* out, int iters, uint32_t seed) {
  uint32_t lo[NCH], hi[NCH];
  uint32_t c = threadIdx.x * 2654435761u + seed;
#pragma unroll
  for (int u = 0; u < NCH; u++) { lo[u] = threadIdx.x * (u + 3); hi[u] = seed * (u + 7); }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
      asm volatile("mad.lo.cc.u32 %0, %0, %2, %0;\n\tmadc.hi.cc.u32 %1, %0, %2, %1;" : "+r"(lo[0]), "+r"(hi[0]) : "r"(c));
#pragma unroll
      for (int u = 1; u < NCH - 1; u++)
        asm volatile("madc.lo.cc.u32 %0, %3, %2, %0;\n\tmadc.hi.cc.u32 %1, %3, %2, %1;" : "+r"(lo[u]), "+r"(hi[u]) : "r"(c), "r"(lo[u - 1]));
      asm volatile("madc.lo.cc.u32 %0, %3, %2, %0;\n\tmadc.hi.u32 %1, %3, %2, %1;" : "+r"(lo[NCH - 1]), "+r"(hi[NCH - 1]) : "r"(c), "r"(lo[NCH - 2]));
    }
  }
  uint32_t x = 0;
#pragma unroll
  for (int u = 0; u < NCH; u++) x ^= lo[u] ^ hi[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
// DFMA (round toward zero, as the 52-bit-limb product splitting needs)
__global__ void k_dfma(uint32_t* out, int iters, uint32_t seed) {
  double x[NCH];
  double c = 1.0 + (threadIdx.x + seed) * 1e-9, m = 1e-3 * (seed | 1);
#pragma unroll
  for (int u = 0; u < NCH; u++) x[u] = (double)(threadIdx.x * (u + 3) + seed);
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int u = 0; u < NCH; u++) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(x[u]) : "d"(c), "d"(m));
    }
  }
  double s = 0;
#pragma unroll
  for (int u = 0; u < NCH; u++) s += x[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)__double2ll_rz(s);
}
__global__ void k_dadd(uint32_t* out, int iters, uint32_t seed) {
  double x[NCH];
  double m = 1e-3 * (seed | 1) + threadIdx.x;
#pragma unroll
  for (int u = 0; u < NCH; u++) x[u] = (double)(threadIdx.x * (u + 3) + seed);
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int u = 0; u < NCH; u++) asm volatile("add.rz.f64 %0, %0, %1;" : "+d"(x[u]) : "d"(m));
    }
  }
  double s = 0;
#pragma unroll
  for (int u = 0; u < NCH; u++) s += x[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)__double2ll_rz(s);
}
// ALU: IADD3 x = x + x_prev + c (one instruction per counted op)
__global__ void k_iadd3(uint32_t* out, int iters, uint32_t seed) {
  uint32_t x[NCH];
  uint32_t c = threadIdx.x * 2654435761u + seed;
#pragma unroll
  for (int u = 0; u < NCH; u++) x[u] = threadIdx.x * (u + 3) + seed;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int u = 0; u < NCH; u++)
        asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(x[u]) : "r"(x[(u + 1) % NCH]), "r"(c));
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int u = 0; u < NCH; u++) s ^= x[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 64-bit integer add (IADD3 + IADD3.X, or a native 64-bit add if the target has one)
__global__ void k_add64(uint32_t* out, int iters, uint32_t seed) {
  uint64_t x[NCH];
#pragma unroll
  for (int u = 0; u < NCH; u++) x[u] = (uint64_t)threadIdx.x * (u + 3) + ((uint64_t)seed << 31);
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int u = 0; u < NCH; u++) asm volatile("add.u64 %0, %0, %1;" : "+l"(x[u]) : "l"(x[(u + 1) % NCH]));
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int u = 0; u < NCH; u++) s ^= x[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(s ^ (s >> 32));
}
// SHF + LOP3 (limb re-alignment work of a reduced-radix core)
__global__ void k_shf(uint32_t* out, int iters, uint32_t seed) {
  uint32_t x[NCH];
#pragma unroll
  for (int u = 0; u < NCH; u++) x[u] = threadIdx.x * (u + 3) + seed;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int u = 0; u < NCH; u++) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(x[u]) : "r"(x[(u + 1) % NCH]));
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int u = 0; u < NCH; u++) s ^= x[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- co-issue mixes: A = 4 chains of one kind, B = 4 chains of another, interleaved 1:1
#define MIX_KERNEL(NAME, DECL, INIT, OPA, OPB, FOLD)                                      \
  __global__ void NAME(uint32_t* out, int iters, uint32_t seed) {                         \
    DECL;                                                                                 \
    uint32_t c = threadIdx.x * 2654435761u + seed;                                        \
    uint32_t mm[4], cc[8];                                                                \
    for (int u = 0; u < 4; u++) mm[u] = seed * (u + 1) + threadIdx.x;                     \
    for (int r = 0; r < 8; r++) cc[r] = c + seed * (r + 1);                               \
    INIT;                                                                                 \
    for (int i = 0; i < iters; i++) {                                                     \
      _Pragma("unroll") for (int r = 0; r < 8; r++) {                                     \
        _Pragma("unroll") for (int u = 0; u < 4; u++) { OPA; OPB; }                       \
      }                                                                                   \
      _Pragma("unroll") for (int u = 0; u < 4; u++) mm[u] += i;                           \
    }                                                                                     \
    uint64_t s = mm[0] ^ cc[0];                                                           \
    FOLD;                                                                                 \
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(s ^ (s >> 32));               \
  }

#define OP_WIDE(X) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(X) : "r"(mm[u]), "r"(cc[r]))
#define OP_IADD3(X, Y) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(X) : "r"(Y), "r"(c))
#define OP_DFMA(X) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(X) : "d"(dc), "d"(dm))
#define OP_IMAD(X) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(X) : "r"(c), "r"(seed))
#define OP_WIDEX(L, H) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(L), "+r"(H) : "r"(mm[u]), "r"(cc[r]))

MIX_KERNEL(k_mix_wide_iadd3, uint64_t a[4]; uint32_t b[4],
           for (int u = 0; u < 4; u++) { a[u] = threadIdx.x * (u + 3) + seed; b[u] = u * 7 + seed; },
           OP_WIDE(a[u]), OP_IADD3(b[u], b[(u + 1) & 3]),
           for (int u = 0; u < 4; u++) s ^= a[u] ^ b[u])
MIX_KERNEL(k_mix_dfma_iadd3, double a[4]; uint32_t b[4]; double dc = 1.0 + threadIdx.x * 1e-9; double dm = 1e-3 * seed,
           for (int u = 0; u < 4; u++) { a[u] = threadIdx.x * (u + 3) + seed; b[u] = u * 7 + seed; },
           OP_DFMA(a[u]), OP_IADD3(b[u], b[(u + 1) & 3]),
           for (int u = 0; u < 4; u++) s ^= (uint64_t)__double_as_longlong(a[u]) ^ b[u])
MIX_KERNEL(k_mix_dfma_wide, double a[4]; uint64_t b[4]; double dc = 1.0 + threadIdx.x * 1e-9; double dm = 1e-3 * seed,
           for (int u = 0; u < 4; u++) { a[u] = threadIdx.x * (u + 3) + seed; b[u] = u * 7 + seed; },
           OP_DFMA(a[u]), OP_WIDE(b[u]),
           for (int u = 0; u < 4; u++) s ^= (uint64_t)__double_as_longlong(a[u]) ^ b[u])
MIX_KERNEL(k_mix_dfma_widex, double a[4]; uint32_t bl[4]; uint32_t bh[4]; double dc = 1.0 + threadIdx.x * 1e-9; double dm = 1e-3 * seed,
           for (int u = 0; u < 4; u++) { a[u] = threadIdx.x * (u + 3) + seed; bl[u] = u * 7 + seed; bh[u] = u + seed; },
           OP_DFMA(a[u]), OP_WIDEX(bl[u], bh[u]),
           for (int u = 0; u < 4; u++) s ^= (uint64_t)__double_as_longlong(a[u]) ^ bl[u] ^ bh[u])
MIX_KERNEL(k_mix_imad_iadd3, uint32_t a[4]; uint32_t b[4],
           for (int u = 0; u < 4; u++) { a[u] = threadIdx.x * (u + 3) + seed; b[u] = u * 7 + seed; },
           OP_IMAD(a[u]), OP_IADD3(b[u], b[(u + 1) & 3]),
           for (int u = 0; u < 4; u++) s ^= a[u] ^ b[u])
MIX_KERNEL(k_mix_widex_iadd3, uint32_t al[4]; uint32_t ah[4]; uint32_t b[4],
           for (int u = 0; u < 4; u++) { al[u] = threadIdx.x * (u + 3) + seed; ah[u] = u; b[u] = u * 7 + seed; },
           OP_WIDEX(al[u], ah[u]), OP_IADD3(b[u], b[(u + 1) & 3]),
           for (int u = 0; u < 4; u++) s ^= al[u] ^ ah[u] ^ b[u])
// three pipes at once: DFMA + IMAD.WIDE + 2 x IADD3
__global__ void k_mix3(uint32_t* out, int iters, uint32_t seed) {
  double a[4];
  uint64_t b[4];
  uint32_t e[4];
  uint32_t c = threadIdx.x * 2654435761u + seed;
  uint32_t mm[4], cc[8];
  for (int u = 0; u < 4; u++) mm[u] = seed * (u + 1) + threadIdx.x;
  for (int r = 0; r < 8; r++) cc[r] = c + seed * (r + 1);
  double dc = 1.0 + threadIdx.x * 1e-9, dm = 1e-3 * seed;
  for (int u = 0; u < 4; u++) { a[u] = threadIdx.x * (u + 3) + seed; b[u] = u * 7 + seed; e[u] = u * 5 + seed; }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int u = 0; u < 4; u++) { OP_DFMA(a[u]); OP_WIDE(b[u]); OP_IADD3(e[u], e[(u + 1) & 3]); OP_IADD3(e[(u + 2) & 3], e[u]); }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) mm[u] += i;
  }
  uint64_t s = 0;
  for (int u = 0; u < 4; u++) s ^= (uint64_t)__double_as_longlong(a[u]) ^ b[u] ^ e[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(s ^ (s >> 32));
}

template <class K>
static void run(const char* name, K kernel, double inst_per_iter, int sms, double mhz, const char* note = "") {
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
  printf("%-18s %8.3f ms  %6.3f warp-inst/clk/SM  (%5.1f lanes/clk/SM)  %s\n", name, best, warp_inst / cycles / sms,
         32.0 * warp_inst / cycles / sms, note);
  cudaFree(d);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  double mhz = p.clockRate / 1000.0;
  printf("%s, %d SMs, %.0f MHz (rates assume the SM clock stays at this value)\n", p.name, p.multiProcessorCount, mhz);
  int sms = p.multiProcessorCount;
  run("imad.wide", k_wide, 64, sms, mhz, "IMAD.WIDE.U32, no carry");
  run("imad.wide.rz", k_wide_rz, 64, sms, mhz, "IMAD.WIDE.U32 Rd, Ra, Rb, RZ");
  run("imad.wide.acc", k_wide_acc, 64, sms, mhz, "mad.lo.cc+madc.hi pairs (see SASS mix)");
  run("imad.wide.x", k_wide_x, 64, sms, mhz, "IMAD.WIDE.U32.X carry chain");
  run("imad.lo", k_imad, 64, sms, mhz, "IMAD (32-bit)");
  run("imad.hi", k_imad_hi, 64, sms, mhz, "IMAD.HI");
  run("dfma", k_dfma, 64, sms, mhz, "DFMA.RZ");
  run("dadd", k_dadd, 64, sms, mhz, "DADD.RZ");
  run("iadd3", k_iadd3, 64, sms, mhz, "IADD3");
  run("add.u64", k_add64, 64, sms, mhz, "64-bit add (counted once)");
  run("shf", k_shf, 64, sms, mhz, "SHF");
  run("wide|iadd3", k_mix_wide_iadd3, 32, sms, mhz, "pairs: IMAD.WIDE + IADD3");
  run("widex|iadd3", k_mix_widex_iadd3, 32, sms, mhz, "pairs: IMAD.WIDE carry + IADD3");
  run("imad|iadd3", k_mix_imad_iadd3, 32, sms, mhz, "pairs: IMAD + IADD3");
  run("dfma|iadd3", k_mix_dfma_iadd3, 32, sms, mhz, "pairs: DFMA + IADD3");
  run("dfma|wide", k_mix_dfma_wide, 32, sms, mhz, "pairs: DFMA + IMAD.WIDE");
  run("dfma|widex", k_mix_dfma_widex, 32, sms, mhz, "pairs: DFMA + IMAD.WIDE carry");
  run("dfma|wide|2iadd3", k_mix3, 32, sms, mhz, "quads: DFMA + IMAD.WIDE + 2 IADD3");
  return 0;
}
