// Batched small-MSM kernel: one CTA per independent MSM, one thread per bucket.
//
// Replaces (*G1Jac).MultiExp at every size Prove/Verify produce (1 .. 5*ell+8
// terms; SURVEY.md §8a rows a1,a3-a6,a8,a10-a15), many MSMs per launch (the 4/6
// MSMs of an IPA / SameMSM round, times the number of proofs in a batch).
//
// GLV halves, signed 4-bit windows, W = 32 windows x 8 buckets = 256 threads.  Phases:
//   1. recode   : scalars Montgomery -> canonical -> k = +-|k1| +- k2*lambda, both
//                 halves biased by 0x0888..8 so that window digits are
//                 ((k' >> 4w) & 15) - 8 with no carry chain at lookup time;
//                 the two biased halves (32 B/term) are staged in shared memory.
//   2. accumulate: thread (w, d) walks the terms, adds every point whose digit
//                 in window w is +-d into its private XYZZ bucket (mixed add).
//                 Lanes search for their next term independently (cheap) and
//                 meet for the field arithmetic, so lanes are busy on different
//                 points at once.
//   3. reduce   : sum_d d*B_d per window as suffix scan + butterfly over the 8
//                 lanes of a window with warp shuffles (6 adds instead of 16).
//   4. combine  : binary tree over windows, level L doubles the upper operand
//                 4*2^L times (124 doublings on the critical path, 5 adds), a quad of
//                 lanes per pair (quad.cuh: warp-cooperative doubling / addition).
//   5. normalise: one inversion, affine result (+ optional compressed bytes).
#define CDL_FP_MUL_CALL 1  // one shared product body: the hot loops fit the instruction caches (mont.cuh)
#include "codec.cuh"
#include "quad.cuh"
#include "launch.h"

namespace cdl {

constexpr int kMsmC = 4;
constexpr int kMsmBuckets = 8;     // 2^(c-1)
constexpr int kMsmWindows = 32;    // signed 4-bit windows of the 127-bit GLV halves
constexpr int kMsmThreads = kMsmBuckets * kMsmWindows;

__device__ __forceinline__ void xyzz_shfl_down(G1Xyzz& r, const G1Xyzz& p, int delta, int width) {
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&p);
  uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < 48; i++) d[i] = __shfl_down_sync(0xffffffffu, s[i], delta, width);
}

// Phases 1-3 for one CTA / one task: on return lane 0 of every 8-lane group
// holds the window sum S_w (w = tid >> 3) in `acc`.  Uses T*36 bytes of smem.
__device__ __forceinline__ void msm_bucket_phases(const G1Affine* __restrict__ points, const uint32_t* __restrict__ idx,
                                                  const Fr* __restrict__ scalars, const MsmTask& task,
                                                  uint8_t* smem_raw, G1Xyzz& acc) {
  const int T = (int)task.term_cnt;
  const int tid = threadIdx.x;
  // layout: kp[T][8] words (biased |k1|, biased k2), then pidx[T] words (bit 31 / 30: negate P / phi(P))
  uint32_t* kp = reinterpret_cast<uint32_t*>(smem_raw);
  uint32_t* pidx = kp + (size_t)T * 8;

  // ---- phase 1: recode.  k = +-|k1| +- k2*lambda (GLV, both halves < 2^127), biased so that a
  // signed window digit is one nibble - 8 (glv.cuh): half the windows and half the doublings of
  // the final combine compared with windows over the 255-bit scalar.
  for (int t = tid; t < T; t += kMsmThreads) {
    Fr km = scalars[task.term_off + t], k;
    FrM::from_mont(k, km);
    Glv g;
    glv_decompose(g, k.v);
    uint32_t b1[4], b2[4];
    glv_bias(b1, g.k1);
    glv_bias(b2, g.k2);
#pragma unroll
    for (int i = 0; i < 4; i++) { kp[t * 8 + i] = b1[i]; kp[t * 8 + 4 + i] = b2[i]; }
    uint32_t raw = idx[task.term_off + t];
    bool flip = (raw >> 31) != 0;
    pidx[t] = (raw & 0x3fffffffu) | ((g.neg1 != flip) ? 0x80000000u : 0u) | ((g.neg2 != flip) ? 0x40000000u : 0u);
  }
  __syncthreads();

  // ---- phase 2: bucket accumulation over the 2T (term, half) pairs
  const int w = tid >> 3;
  const int mybucket = (tid & 7) + 1;
  xyzz_set_inf(acc);
  Fp beta;
  fp_set_beta(beta);
  int t = 0;  // index into the 2T pairs: pair = 2*term + half
  while (true) {
    int found = -1;
    int sgn = 0;
    while (t < 2 * T) {
      int d = glv_digit(kp + (t >> 1) * 8 + (t & 1) * 4, w);
      int a = d < 0 ? -d : d;
      if (a == mybucket) { found = t; sgn = d; t++; break; }
      t++;
    }
    if (!__any_sync(0xffffffffu, found >= 0)) break;
    if (found >= 0) {
      uint32_t pi = pidx[found >> 1];
      G1Affine q = points[pi & 0x3fffffffu];
      const int h = found & 1;
      if (h) FpM::mul(q.x, q.x, beta);
      bool neg = (((pi >> (31 - h)) & 1u) != 0) != (sgn < 0);
      if (neg) FpM::neg(q.y, q.y);
      xyzz_add_mixed(acc, acc, q);
    }
  }

  // ---- phase 3: window sum S_w = sum_d d * B_d over the 8 lanes of the window
  // steps 0..2: suffix sums R_d = sum_{j>=d} B_j (offsets 1,2,4);
  // steps 3..5: S_w = sum_d R_d (offsets 4,2,1).  One loop => one inlined add.
  G1Xyzz other;
#pragma unroll 1
  for (int step = 0; step < 6; step++) {
    int off = step < 3 ? (1 << step) : (4 >> (step - 3));
    xyzz_shfl_down(other, acc, off, kMsmBuckets);
    bool take = step < 3 ? ((tid & 7) + off < kMsmBuckets) : ((tid & 7) < off);
    if (take) xyzz_add(acc, acc, other);
  }
}

// Latency path (few MSMs in flight): one CTA does everything for its task.
// points: pool of affine points; idx[t] selects the base of term t (bit 31:
// negate).  scalars[t]: gnark Montgomery fr.Element.  out_aff[task.out_idx] /
// out_c48[j]: result of task j (either pointer may be null).
__global__ void __launch_bounds__(kMsmThreads, 1)
k_msm_small(const G1Affine* __restrict__ points, const uint32_t* __restrict__ idx,
            const Fr* __restrict__ scalars, const MsmTask* __restrict__ tasks,
            G1Affine* __restrict__ out_aff, uint8_t* __restrict__ out_c48) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const MsmTask task = tasks[blockIdx.x];
  const int tid = threadIdx.x;
  const int w = tid >> 3;
  G1Xyzz acc;
  msm_bucket_phases(points, idx, scalars, task, smem_raw, acc);

  // ---- phase 4: combine windows.  Binary tree over the 32 window sums, level L doubles the upper
  // operand 4 * 2^L times: 124 doublings + 5 additions on the critical path.  Every pair is handled by
  // a QUAD of lanes (quad.cuh): the independent products of a doubling / addition run in parallel
  // lanes, three / four product latencies instead of nine / fourteen.
  __syncthreads();  // the recode staging is dead from here on; reuse shared memory
  G1Xyzz* win = reinterpret_cast<G1Xyzz*>(smem_raw);
  if ((tid & 7) == 0) win[w] = acc;
  __syncthreads();
  const Quad q;
  const int g = tid >> 2;  // 64 quads per CTA
#pragma unroll 1
  for (int level = 0, active = kMsmWindows / 2; active >= 1; level++, active >>= 1) {
    G1Xyzz lo, hi;
    if (g < active) {
      lo = win[2 * g];
      hi = win[2 * g + 1];
      int nd = kMsmC << level;
#pragma unroll 1
      for (int i = 0; i < nd; i++) qxyzz_dbl(q, hi, hi);
      qxyzz_add(q, lo, lo, hi);
    }
    __syncthreads();
    if (g < active && q.lane == 0) win[g] = lo;
    __syncthreads();
  }

  // ---- phase 5: normalise + store
  if (g == 0) {
    G1Affine a;
    qxyzz_to_affine(q, a, win[0]);
    if (q.lane == 0) {
      if (out_aff) out_aff[task.out_idx] = a;
      if (out_c48) g1_compress_dev(out_c48 + 48 * (size_t)blockIdx.x, a);
    }
  }
}

// shared memory bytes for the largest task of a launch
static size_t msm_small_smem_bytes(size_t max_terms) {
  size_t a = max_terms * 36;                       // kp + pidx
  size_t b = (size_t)kMsmWindows * sizeof(G1Xyzz);  // combine scratch
  return a > b ? a : b;
}

cudaError_t msm_small_init() {
  return cudaFuncSetAttribute(k_msm_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMsmMaxSmem);
}

void launch_msm_small(const G1Affine* points, const uint32_t* idx, const Fr* scalars, const MsmTask* tasks,
                      int ntasks, size_t max_terms, G1Affine* out_aff, uint8_t* out_c48, cudaStream_t st) {
  k_msm_small<<<ntasks, kMsmThreads, msm_small_smem_bytes(max_terms), st>>>(points, idx, scalars, tasks, out_aff,
                                                                             out_c48);
}

}  // namespace cdl
