// Throughput path for many small MSMs per launch (batched proving / verification), and for a
// few long ones (the verifier's 5*ell + 8 terms for a single proof).
//
//   k_msm_recode    thread per term: fr.Element (Montgomery) -> canonical -> GLV split
//                   k = s0*(s1*|k1| + k2*lambda), both < 2^127, stored biased so that a window
//                   digit is one nibble - 8; beta*x of the base (the x of phi(P)) once per term
//   k_msm_warp_gmem warp per chunk (<= 128 terms, msm_tp_pick_chunk) of a task, lane = one of the
//                   32 signed 4-bit windows: every lane walks the chunk's terms serially and adds
//                   +-P / +-phi(P) into its 8 XYZZ buckets, then reduces them with the running-sum
//                   trick from the highest non-empty bucket.  All 32 lanes do the same amount of
//                   work on every term (the term's point is one broadcast load), so lane
//                   efficiency is ~15/16 whatever the MSM size — unlike thread-per-bucket, whose
//                   lanes idle on load imbalance when a window has few points per bucket.
//                   The buckets (49 KB per warp) live in a global scratch whose regions are
//                   reused per SM slot and pinned in L2 (see k_msm_warp_gmem).
//   k_msm_warp / k_msm_warp_smem<>: the same walk with the buckets in local memory / in shared
//                   memory; measured alternatives, selectable with CDL_MSM_WARP=l / j / x
//   k_msm_chunk_sum thread per (task, window): sums the chunk partials
//   k_msm_combine_tp thread per task: walks the windows top-down (4 doublings + 1 addition),
//                   normalises, stores the affine point and its 48-byte encoding.  The 124-doubling
//                   chain is serial per MSM but runs at full lane efficiency across thousands of
//                   MSMs, and overlaps the other lanes' kernels in a batched call.
#include <algorithm>
#include <cstdlib>
#include <mutex>

#define CDL_FP_MUL_CALL 1
#include "codec.cuh"
#include "launch.h"

namespace cdl {

constexpr int kTpWindows = 32;

// Streaming (evict-first) loads for data a warp reads once — the recoded terms and the base
// points — so that they do not push the warps' bucket arrays (local memory, re-read on every
// addition) out of L2.
template <class T>
__device__ __forceinline__ T ld_stream8(const T* p) {
  static_assert(sizeof(T) % 8 == 0, "8-byte granules");
  T v;
  const uint2* s = reinterpret_cast<const uint2*>(p);
  uint2* d = reinterpret_cast<uint2*>(&v);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 8); i++) d[i] = __ldcs(s + i);
  return v;
}
template <class T>
__device__ __forceinline__ T ld_stream16(const T* p) {
  static_assert(sizeof(T) % 16 == 0, "16-byte granules");
  T v;
  const uint4* s = reinterpret_cast<const uint4*>(p);
  uint4* d = reinterpret_cast<uint4*>(&v);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = __ldcs(s + i);
  return v;
}

struct alignas(16) MsmRec {
  uint32_t k1p[4];
  uint32_t k2p[4];
  uint32_t pidx;   // pool index of the base
  uint32_t flags;  // bit 0: negate the P part, bit 1: negate the phi(P) part
  uint32_t pad[2];
  Fp bx;           // beta * x of the base: the x coordinate of phi(P), computed once per term here
                   // instead of once per term by all 32 lanes of the bucket warp
};

__global__ void k_msm_recode(const G1Affine* __restrict__ points, const uint32_t* __restrict__ idx,
                             const Fr* __restrict__ scalars, MsmRec* __restrict__ rec, int nterm) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nterm) return;
  Fr km = scalars[t], k;
  FrM::from_mont(k, km);
  Glv g;
  glv_decompose(g, k.v);
  MsmRec r;
  glv_bias(r.k1p, g.k1);
  glv_bias(r.k2p, g.k2);
  r.pidx = idx[t] & 0x7fffffffu;
  bool flip = (idx[t] >> 31) != 0;
  r.flags = ((g.neg1 != flip) ? 1u : 0u) | ((g.neg2 != flip) ? 2u : 0u);
  r.pad[0] = r.pad[1] = 0;
  Fp beta, x = points[r.pidx].x;
  fp_set_beta(beta);
  FpM::mul(r.bx, x, beta);
  rec[t] = r;
}

#ifndef CDL_WARP_MIN_CTAS
#define CDL_WARP_MIN_CTAS 2
#endif
__global__ void __launch_bounds__(128, CDL_WARP_MIN_CTAS)
k_msm_warp(const G1Affine* __restrict__ points, const MsmRec* __restrict__ rec, const MsmSub* __restrict__ subs,
           int nsub, G1Jac* __restrict__ win) {
  const int sub = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (sub >= nsub) return;  // whole warp
  const int w = threadIdx.x & 31;
  const MsmSub s = subs[sub];
  G1Xyzz bk[8];
  uint32_t nonempty = 0;
#pragma unroll 1
  for (uint32_t t = 0; t < s.term_cnt; t++) {
    const MsmRec r = ld_stream16(rec + s.term_off + t);
    G1Affine p = ld_stream16(points + r.pidx);
    if (aff_is_inf(p)) continue;  // uniform across the warp
    const Fp& bx = r.bx;
#pragma unroll 1
    for (int h = 0; h < 2; h++) {
      int d = glv_digit(h == 0 ? r.k1p : r.k2p, w);
      if (d == 0) continue;
      bool neg = (d < 0) != (((r.flags >> h) & 1u) != 0);
      int a = (d < 0 ? -d : d) - 1;
      G1Affine q;
      q.x = h == 0 ? p.x : bx;
      q.y = p.y;
      if (neg) FpM::neg(q.y, q.y);
      if (!((nonempty >> a) & 1u)) {
        xyzz_from_affine(bk[a], q);
        nonempty |= 1u << a;
      } else {
        G1Xyzz b = bk[a];
        xyzz_add_mixed(b, b, q);
        bk[a] = b;
      }
    }
  }
  // sum_d d * B_d by the running-sum trick
  G1Xyzz run, acc;
  xyzz_set_inf(run);
  xyzz_set_inf(acc);
  // buckets above the highest non-empty one contribute nothing: start there (short chunks in the
  // late folding rounds touch one or two buckets per window)
#pragma unroll 1
  for (int a = nonempty ? 31 - __clz(nonempty) : -1; a >= 0; a--) {
    if ((nonempty >> a) & 1u) {
      G1Xyzz b = bk[a];
      xyzz_add(run, run, b);
    }
    xyzz_add(acc, acc, run);
  }
  G1Jac j;
  xyzz_to_jac(j, acc);
  win[(size_t)w * nsub + sub] = j;
}

// Same walk with the 8 buckets of every lane in SHARED memory (Jacobian, 8 x 144 B x 32 lanes =
// 36 KB per warp, one warp per CTA, six CTAs per SM): word i of bucket a of lane l lives at
// sm[(a*36 + i)*32 + l], so every access is bank-conflict free whatever bucket each lane picks.
// The private-array form above keeps the buckets in local memory, whose 49 KB per warp thrash
// L1/L2 and show up as DRAM traffic hundreds of times the algorithmic bytes.
// Same walk with the buckets in an explicitly managed GLOBAL scratch: one 49 KB region per
// resident warp slot (SM x CTA slot x warp), claimed from a per-SM bitmap when the CTA starts and
// released when it ends, so that every CTA that ever runs on an SM reuses the same addresses and
// the whole working set (resident warps x 49 KB = 58 MB) stays in L2 with ordinary write-back
// caching.  uint4 granules, layout [bucket][granule][lane]: a warp access is 512 contiguous bytes.
constexpr int kSlotsPerSm = 4;

// %smid values are below %nsmid, which can exceed the SM count when units are fused off
__global__ void k_query_nsmid(uint32_t* out) {
  uint32_t n;
  asm volatile("mov.u32 %0, %%nsmid;" : "=r"(n));
  out[0] = n;
}
constexpr size_t kWarpBucketBytes = 8 * sizeof(G1Xyzz) * 32;

__global__ void __launch_bounds__(128, 2)
k_msm_warp_gmem(const G1Affine* __restrict__ points, const MsmRec* __restrict__ rec, const MsmSub* __restrict__ subs,
                int nsub, G1Jac* __restrict__ win, uint4* __restrict__ scratch, uint32_t* __restrict__ bitmap,
                uint32_t nsmid) {
  __shared__ uint32_t slot_sh;
  uint32_t smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  if (smid >= nsmid) __trap();
  if (threadIdx.x == 0) {
    uint32_t k = 0;
    for (uint32_t tries = 0;; tries++) {
      uint32_t old = atomicOr(&bitmap[smid], 1u << k);
      if (!((old >> k) & 1u)) break;
      k = (k + 1) % kSlotsPerSm;
      if (tries > (1u << 22)) __trap();  // a leaked slot must surface as an error, never as a hang
    }
    slot_sh = k;
  }
  __syncthreads();
  const uint32_t slot = slot_sh;
  const int sub = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int w = threadIdx.x & 31;
  // slot-major: the two slots normally in use on every SM form one contiguous prefix of the buffer
  uint4* bk = scratch + (((size_t)slot * nsmid + smid) * 4 + (threadIdx.x >> 5)) * (kWarpBucketBytes / 16) + w;
  if (sub < nsub) {
    const MsmSub s = subs[sub];
    uint32_t nonempty = 0;
#pragma unroll 1
    for (uint32_t t = 0; t < s.term_cnt; t++) {
      const MsmRec r = ld_stream16(rec + s.term_off + t);
      G1Affine p = ld_stream16(points + r.pidx);
      if (aff_is_inf(p)) continue;  // uniform across the warp
      const Fp& bx = r.bx;
#pragma unroll 1
      for (int h = 0; h < 2; h++) {
        int d = glv_digit(h == 0 ? r.k1p : r.k2p, w);
        if (d == 0) continue;
        bool neg = (d < 0) != (((r.flags >> h) & 1u) != 0);
        int a = (d < 0 ? -d : d) - 1;
        G1Affine q;
        q.x = h == 0 ? p.x : bx;
        q.y = p.y;
        if (neg) FpM::neg(q.y, q.y);
        uint4* slotp = bk + (size_t)a * 12 * 32;
        G1Xyzz b;
        uint4* bw = reinterpret_cast<uint4*>(&b);
        if (!((nonempty >> a) & 1u)) {
          xyzz_from_affine(b, q);
          nonempty |= 1u << a;
        } else {
#pragma unroll
          for (int i = 0; i < 12; i++) bw[i] = slotp[i * 32];
          xyzz_add_mixed(b, b, q);
        }
#pragma unroll
        for (int i = 0; i < 12; i++) slotp[i * 32] = bw[i];
      }
    }
    G1Xyzz run, acc;
    xyzz_set_inf(run);
    xyzz_set_inf(acc);
#pragma unroll 1
    for (int a = nonempty ? 31 - __clz(nonempty) : -1; a >= 0; a--) {
      if ((nonempty >> a) & 1u) {
        G1Xyzz b;
        uint4* bw = reinterpret_cast<uint4*>(&b);
        const uint4* slotp = bk + (size_t)a * 12 * 32;
#pragma unroll
        for (int i = 0; i < 12; i++) bw[i] = slotp[i * 32];
        xyzz_add(run, run, b);
      }
      xyzz_add(acc, acc, run);
    }
    G1Jac j;
    xyzz_to_jac(j, acc);
    win[(size_t)w * nsub + sub] = j;
  }
  __syncthreads();
  if (threadIdx.x == 0) atomicAnd(&bitmap[smid], ~(1u << slot));
}

// device-wide scratch of the gmem variant, shared by every context / lane on the device
struct GmemScratch {
  uint4* buf = nullptr;
  uint32_t* bitmap = nullptr;
  uint32_t nsmid = 0;       // number of %smid values on this device
  size_t hot_bytes = 0;     // prefix normally in use (2 CTA slots per SM)
  size_t persist_bytes = 0; // size of the L2 persisting carve-out this path asks for (0: unsupported)
  bool persist = false;     // the carve-out is currently set
};
static std::mutex g_scratch_mu;
static GmemScratch g_scratch[64];

// The persisting carve-out takes L2 away from everything else on the device, so it is only held
// while batched-MSM launches are being issued: the large-MSM path (whose point gathers want the
// whole L2) releases it, the next batched launch takes it back.
void msm_l2_carveout(bool on) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(g_scratch_mu);
  GmemScratch& g = g_scratch[dev & 63];
  if (!g.persist_bytes || g.persist == on) return;
  if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, on ? g.persist_bytes : 0) == cudaSuccess) g.persist = on;
  cudaGetLastError();
}
static GmemScratch* gmem_scratch() {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(g_scratch_mu);
  GmemScratch& g = g_scratch[dev & 63];
  if (!g.buf) {
    if (cudaMalloc(&g.bitmap, 4096) != cudaSuccess) return nullptr;
    cudaMemset(g.bitmap, 0, 4096);
    k_query_nsmid<<<1, 1>>>(g.bitmap + 1000);
    uint32_t nsmid = 0;
    if (cudaMemcpy(&nsmid, g.bitmap + 1000, 4, cudaMemcpyDeviceToHost) != cudaSuccess || nsmid == 0 || nsmid > 1000) {
      cudaFree(g.bitmap);
      cudaGetLastError();
      return nullptr;
    }
    cudaMemset(g.bitmap, 0, 4096);
    g.nsmid = nsmid;
    size_t bytes = (size_t)nsmid * kSlotsPerSm * 4 * kWarpBucketBytes;
    if (cudaMalloc(&g.buf, bytes) != cudaSuccess) { cudaFree(g.bitmap); g.buf = nullptr; cudaGetLastError(); return nullptr; }
    g.hot_bytes = (size_t)2 * nsmid * 4 * kWarpBucketBytes;
    // keep the bucket scratch resident in L2: persisting carve-out + access-policy window (set per stream)
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess && prop.persistingL2CacheMaxSize > 0 && !getenv("CDL_NO_L2_PERSIST")) {
      g.persist_bytes = std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, g.hot_bytes);
    }
    cudaGetLastError();
  }
  return &g;
}

template <class Bucket>
struct SmemBucketOps;
template <>
struct SmemBucketOps<G1Jac> {
  static __device__ __forceinline__ void from_affine(G1Jac& b, const G1Affine& q) { jac_from_affine(b, q); }
  static __device__ __forceinline__ void add_mixed(G1Jac& b, const G1Affine& q) { jac_add_mixed(b, b, q); }
  static __device__ __forceinline__ void add(G1Jac& r, const G1Jac& b) { jac_add(r, r, b); }
  static __device__ __forceinline__ void set_inf(G1Jac& b) { jac_set_inf(b); }
  static __device__ __forceinline__ void to_jac(G1Jac& j, const G1Jac& b) { j = b; }
};
template <>
struct SmemBucketOps<G1Xyzz> {
  static __device__ __forceinline__ void from_affine(G1Xyzz& b, const G1Affine& q) { xyzz_from_affine(b, q); }
  static __device__ __forceinline__ void add_mixed(G1Xyzz& b, const G1Affine& q) { xyzz_add_mixed(b, b, q); }
  static __device__ __forceinline__ void add(G1Xyzz& r, const G1Xyzz& b) { xyzz_add(r, r, b); }
  static __device__ __forceinline__ void set_inf(G1Xyzz& b) { xyzz_set_inf(b); }
  static __device__ __forceinline__ void to_jac(G1Jac& j, const G1Xyzz& b) { xyzz_to_jac(j, b); }
};

template <class Bucket>
__global__ void __launch_bounds__(32)
k_msm_warp_smem(const G1Affine* __restrict__ points, const MsmRec* __restrict__ rec,
                const MsmSub* __restrict__ subs, int nsub, G1Jac* __restrict__ win) {
  using Ops = SmemBucketOps<Bucket>;
  constexpr int NW = sizeof(Bucket) / 4;
  extern __shared__ uint32_t sm[];
  const int sub = blockIdx.x;
  const int w = threadIdx.x;
  const MsmSub s = subs[sub];
  uint32_t nonempty = 0;
#pragma unroll 1
  for (uint32_t t = 0; t < s.term_cnt; t++) {
    const MsmRec r = rec[s.term_off + t];
    G1Affine p = points[r.pidx];
    if (aff_is_inf(p)) continue;  // uniform across the warp
    const Fp& bx = r.bx;
#pragma unroll 1
    for (int h = 0; h < 2; h++) {
      int d = glv_digit(h == 0 ? r.k1p : r.k2p, w);
      if (d == 0) continue;
      bool neg = (d < 0) != (((r.flags >> h) & 1u) != 0);
      int a = (d < 0 ? -d : d) - 1;
      G1Affine q;
      q.x = h == 0 ? p.x : bx;
      q.y = p.y;
      if (neg) FpM::neg(q.y, q.y);
      uint32_t* slot = sm + (size_t)a * NW * 32 + w;
      Bucket b;
      uint32_t* bw = reinterpret_cast<uint32_t*>(&b);
      if (!((nonempty >> a) & 1u)) {
        Ops::from_affine(b, q);
        nonempty |= 1u << a;
      } else {
#pragma unroll
        for (int i = 0; i < NW; i++) bw[i] = slot[i * 32];
        Ops::add_mixed(b, q);
      }
#pragma unroll
      for (int i = 0; i < NW; i++) slot[i * 32] = bw[i];
    }
  }
  Bucket run, acc;
  Ops::set_inf(run);
  Ops::set_inf(acc);
#pragma unroll 1
  for (int a = nonempty ? 31 - __clz(nonempty) : -1; a >= 0; a--) {
    if ((nonempty >> a) & 1u) {
      Bucket b;
      uint32_t* bw = reinterpret_cast<uint32_t*>(&b);
      const uint32_t* slot = sm + (size_t)a * NW * 32 + w;
#pragma unroll
      for (int i = 0; i < NW; i++) bw[i] = slot[i * 32];
      Ops::add(run, b);
    }
    Ops::add(acc, run);
  }
  G1Jac j;
  Ops::to_jac(j, acc);
  win[(size_t)w * nsub + sub] = j;
}

// thread per (task, window): sum of the task's chunk partials, so that the serial Horner
// chain of k_msm_combine_tp sees one point per window however finely a task was cut
__global__ void __launch_bounds__(128)
k_msm_chunk_sum(const G1Jac* __restrict__ win, const MsmTask2* __restrict__ tasks, int ntasks, int nsub,
                G1Jac* __restrict__ wsum) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntasks * kTpWindows) return;
  const int w = t / ntasks, j = t - w * ntasks;
  const MsmTask2 task = tasks[j];
  G1Jac acc;
  jac_set_inf(acc);
#pragma unroll 1
  for (uint32_t c = 0; c < task.sub_cnt; c++) {
    G1Jac s = win[(size_t)w * nsub + task.sub_off + c];
    jac_add(acc, acc, s);
  }
  wsum[(size_t)w * ntasks + j] = acc;
}

__global__ void __launch_bounds__(64)
k_msm_combine_tp(const G1Jac* __restrict__ wsum, const MsmTask2* __restrict__ tasks, int ntasks,
                 G1Affine* __restrict__ out_aff, uint8_t* __restrict__ out_c48) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ntasks) return;
  const MsmTask2 task = tasks[j];
  G1Jac acc;
  jac_set_inf(acc);
#pragma unroll 1
  for (int w = kTpWindows - 1; w >= 0; w--) {
    if (w != kTpWindows - 1) {
#pragma unroll 1
      for (int i = 0; i < 4; i++) jac_dbl(acc, acc);
    }
    G1Jac s = wsum[(size_t)w * ntasks + j];
    jac_add(acc, acc, s);
  }
  G1Affine a;
  jac_to_affine(a, acc);
  if (out_aff) out_aff[task.out_idx] = a;
  if (out_c48) g1_compress_dev(out_c48 + 48 * (size_t)j, a);
}

size_t msm_tp_scratch_bytes(size_t nterm, size_t nsub, size_t ntasks) {
  size_t rec = (nterm * sizeof(MsmRec) + 255) & ~(size_t)255;
  size_t win = (nsub * kTpWindows * sizeof(G1Jac) + 255) & ~(size_t)255;
  return rec + win + ntasks * kTpWindows * sizeof(G1Jac);
}

// Chunk length for a launch.  Long chunks amortise the per-chunk bucket reduction (up to 16 full
// additions against 2 mixed additions per term); shorter ones give more warps and a shorter tail.
// Measured on B200 with batches of 2 048 Whisk round trips (MSM class per step): 64 terms 377 ms,
// 128: 378, 160: 374, 256: 384, 512: 404, 1 024: 427; batched verification alone (one 712-term MSM
// per proof) gains 5-12 % from 64-128 over 256.  CDL_MSM_CHUNK overrides the default of 128.
uint32_t msm_tp_pick_chunk(size_t nterm, int sm_count) {
  static const uint32_t forced = [] {
    const char* e = getenv("CDL_MSM_CHUNK");
    int v = e ? atoi(e) : 0;
    return (uint32_t)(v >= 8 && v <= 4096 ? v : 0);
  }();
  if (forced) return forced;
  // latency regime (a handful of large MSMs, e.g. the verifier's 5*ell + 8 terms for one proof):
  // too few 128-term chunks to occupy the machine, so cut finer and shorten each warp's serial walk
  if (nterm / kMsmChunk < (size_t)2 * sm_count) return 32;
  return kMsmChunk;
}

void launch_msm_tp(const G1Affine* points, const uint32_t* idx, const Fr* scalars, int nterm, const MsmSub* subs,
                   int nsub, const MsmTask2* tasks, int ntasks, G1Affine* out_aff, uint8_t* out_c48, void* scratch,
                   cudaStream_t st) {
  MsmRec* rec = (MsmRec*)scratch;
  size_t rec_bytes = ((size_t)nterm * sizeof(MsmRec) + 255) & ~(size_t)255;
  size_t win_bytes = ((size_t)nsub * kTpWindows * sizeof(G1Jac) + 255) & ~(size_t)255;
  G1Jac* win = (G1Jac*)((uint8_t*)scratch + rec_bytes);
  G1Jac* wsum = (G1Jac*)((uint8_t*)scratch + rec_bytes + win_bytes);
  if (nterm > 0) k_msm_recode<<<(nterm + 127) / 128, 128, 0, st>>>(points, idx, scalars, rec, nterm);
  static const int variant = [] {
    const char* e = getenv("CDL_MSM_WARP");
    return !e ? 3 : e[0] == 'j' ? 1 : e[0] == 'x' ? 2 : e[0] == 'l' ? 0 : 3;  // default: global scratch
  }();
  if (nsub > 0) {
    if (variant == 1) {
      k_msm_warp_smem<G1Jac><<<nsub, 32, 8 * sizeof(G1Jac) * 32, st>>>(points, rec, subs, nsub, win);
    } else if (variant == 2) {
      static const cudaError_t attr = cudaFuncSetAttribute(k_msm_warp_smem<G1Xyzz>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                           (int)(8 * sizeof(G1Xyzz) * 32));
      (void)attr;
      k_msm_warp_smem<G1Xyzz><<<nsub, 32, 8 * sizeof(G1Xyzz) * 32, st>>>(points, rec, subs, nsub, win);
    } else if (variant == 3) {
      GmemScratch* g = gmem_scratch();
      if (!g) {  // scratch allocation failed: private (local-memory) buckets
        k_msm_warp<<<(nsub + 3) / 4, 128, 0, st>>>(points, rec, subs, nsub, win);
      } else {
        if (g->persist_bytes) {
          msm_l2_carveout(true);
          cudaDeviceProp prop;
          int dev = 0;
          cudaGetDevice(&dev);
          static int max_window = [&] { int v = 0; cudaDeviceGetAttribute(&v, cudaDevAttrMaxAccessPolicyWindowSize, dev); return v; }();
          cudaStreamAttrValue av = {};
          av.accessPolicyWindow.base_ptr = g->buf;
          av.accessPolicyWindow.num_bytes = std::min<size_t>(g->hot_bytes, (size_t)max_window);
          av.accessPolicyWindow.hitRatio = 1.0f;
          av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
          av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
          cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av);
          (void)prop;
        }
        k_msm_warp_gmem<<<(nsub + 3) / 4, 128, 0, st>>>(points, rec, subs, nsub, win, g->buf, g->bitmap, g->nsmid);
      }
    } else {
      k_msm_warp<<<(nsub + 3) / 4, 128, 0, st>>>(points, rec, subs, nsub, win);
    }
  }
  k_msm_chunk_sum<<<(ntasks * kTpWindows + 127) / 128, 128, 0, st>>>(win, tasks, ntasks, nsub, wsum);
  k_msm_combine_tp<<<(ntasks + 63) / 64, 64, 0, st>>>(wsum, tasks, ntasks, out_aff, out_c48);
}

}  // namespace cdl
