// Host-callable kernel launchers; each is defined next to its kernel so that the
// kernels live in separate translation units (compiled in parallel).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include "g1.cuh"
#include "fixed_base.cuh"

namespace cdl {

struct MsmTask {
  uint32_t term_off;  // first term in idx[] / scalars[]
  uint32_t term_cnt;
  uint32_t out_idx;   // result slot: out_aff[out_idx]
  uint32_t pad;
};
// elementwise op on a pool of affine points:
//   pool[dst] = scalars[sc] * pool[src]  (+ pool[add] unless add == kNoPoint)
struct ElemOp {
  uint32_t src, add, dst, sc;
};
constexpr uint32_t kNoPoint = 0xffffffffu;
// pool[dst .. dst+count) = pool[src .. src+count); src == kNoPoint fills with the point at infinity.
// Ranges of one launch must not overlap each other's destinations.
struct CopyRange {
  uint32_t src, dst, count, pad;
};
constexpr size_t kMsmMaxSmem = 220 * 1024;         // dynamic shared memory opt-in for the small-MSM kernel
constexpr size_t kMsmMaxTerms = kMsmMaxSmem / 36;  // 36 B of staging per term

void launch_compress(const G1Affine* in, uint8_t* out, int n, cudaStream_t s);
void launch_decompress(const uint8_t* in, G1Affine* out, uint8_t* status, int n, cudaStream_t s);
void launch_fp_mul(const Fp* a, const Fp* b, Fp* out, int n, cudaStream_t s);
void launch_peak(int kind, void* out, int blocks, int tpb, int iters, uint32_t seed, cudaStream_t s);
void launch_scalar_mul(const G1Affine* P, const Fr* s, int stride, const G1Affine* L, G1Affine* out, int n,
                       cudaStream_t st);
void launch_jac_to_affine(const G1Jac* in, G1Affine* out, int n, cudaStream_t st);
// ft: fixed-base tables of the pool's first ft.nbase points (fixed_base.cuh); ops on those sources skip the
// double-and-add (the thread-per-op kernel only; small launches take the quad kernel)
void launch_elem_ops(G1Affine* pool, const ElemOp* ops, const Fr* scalars, int n, cudaStream_t st,
                     FixedTable ft = FixedTable());
size_t fixed_table_bytes(uint32_t nbase);
void launch_fixed_build(const G1Affine* bases, uint32_t nbase, G1Affine* tab, cudaStream_t st);
void launch_copy_ranges(G1Affine* pool, const CopyRange* ranges, int n, cudaStream_t st);
// verifier scalar pipeline on the device (k_verify_scalars.cu): per-proof inputs, all Montgomery fr.Element
constexpr int kVsMaxM = 16;  // rounds (n <= 2^16)
struct VsParams {
  Fr a1b, a2c0, a3d0, a4xf, a5xf, a6xf, a7, a8, beta_inv, sH_host;
  Fr gamma[kVsMaxM], gamma_inv[kVsMaxM], ch[kVsMaxM];
  uint32_t sc_base;  // first merged-base scalar of this proof in the stage's scalar array
  uint32_t as_base;  // first vector challenge of this proof in the `as` array
  uint32_t pad[2];
};
void launch_verify_scalars(const VsParams* params, const Fr* as, Fr* sc, uint32_t nproofs, uint32_t ell, uint32_t n,
                           uint32_t m, cudaStream_t st);
// indexed codecs on a pool: enc[i] <- pool[src[i]] ; pool[dst[i]] <- enc[i]
void launch_compress_idx(const G1Affine* pool, const uint32_t* src, uint8_t* out48, int n, cudaStream_t s);
void launch_decompress_idx(const uint8_t* in48, G1Affine* pool, const uint32_t* dst, uint8_t* status, int n,
                           cudaStream_t s);
// one CTA per task; out_aff / out_c48 may be null
cudaError_t msm_small_init();
// --- throughput path (k_msm_tp.cu): tasks are cut into chunks ("subs") of at most
// kMsmChunk terms; one warp per sub, one combine thread per task.
struct MsmSub {
  uint32_t term_off, term_cnt;
};
struct MsmTask2 {
  uint32_t sub_off, sub_cnt, out_idx, pad;
};
constexpr uint32_t kMsmIdxFixed = 1u << 30;  // idx[] flag (throughput path only): sum this term from the fixed-base tables
constexpr uint32_t kMsmChunk = 128;  // terms per bucket warp (msm_tp_pick_chunk)
size_t msm_tp_scratch_bytes(size_t nterm, size_t nsub, size_t ntasks);
uint32_t msm_tp_pick_chunk(size_t nterm, int sm_count);
void msm_l2_carveout(bool on);  // persisting-L2 window of the bucket scratch: held by the batched path only
// ft + fixed_subs: terms flagged kMsmIdxFixed (on the pool's first ft.nbase points) are summed from the fixed-base
// tables by one warp per listed chunk (k_msm_fixed; fixed_subs = device list of the chunks holding such terms)
// and skipped by the bucket warps
void launch_msm_tp(const G1Affine* points, const uint32_t* idx, const Fr* scalars, int nterm, const MsmSub* subs,
                   int nsub, const MsmTask2* tasks, int ntasks, G1Affine* out_aff, uint8_t* out_c48, void* scratch,
                   cudaStream_t st, FixedTable ft = FixedTable(), const uint32_t* fixed_subs = nullptr,
                   int nfixed_subs = 0);
// host helper: cut tasks into subs
template <class VecSub, class VecTask2>
inline void msm_build_subs(const MsmTask* tasks, size_t ntasks, uint32_t chunk, VecSub& subs, VecTask2& tasks2) {
  subs.clear();
  tasks2.clear();
  for (size_t j = 0; j < ntasks; j++) {
    MsmTask2 t2{(uint32_t)subs.size(), 0, tasks[j].out_idx, 0};
    for (uint32_t o = 0; o < tasks[j].term_cnt; o += chunk) {
      uint32_t c = tasks[j].term_cnt - o < chunk ? tasks[j].term_cnt - o : chunk;
      subs.push_back(MsmSub{tasks[j].term_off + o, c});
      t2.sub_cnt++;
    }
    tasks2.push_back(t2);
  }
}

// counts[0..n) -> offsets[0..n] exclusive prefix (total in offsets[n]); counts are zeroed; bsum: 4096 words
cudaError_t launch_exclusive_scan(uint32_t* counts, uint32_t n, uint32_t* bsum, uint32_t* offsets, cudaStream_t st,
                                  uint32_t pad = 0);

// Latency path (k_msm.cu): one CTA per task.  Callers switch to the throughput path at
// kMsmSplitThreshold tasks per launch.
constexpr int kMsmSplitThreshold = 96;
// ... or when one task is long enough that a single CTA's thread-per-bucket walk is the slower choice
constexpr size_t kMsmSplitTerms = 384;
void launch_msm_small(const G1Affine* points, const uint32_t* idx, const Fr* scalars, const MsmTask* tasks,
                      int ntasks, size_t max_terms, G1Affine* out_aff, uint8_t* out_c48, cudaStream_t st);


// --- single large MSM (k_msm_big.cu): signed-digit Pippenger, counting sort of bucket
// indices, thread-per-bucket accumulation.  A caller may own only the windows
// wfirst, wfirst + wstep, ... (multi-GPU window partition).
struct BigMsmDims {
  int n, c, W, M, wfirst, wstep, nlocal;
  uint32_t nb;     // nlocal * M buckets
  uint32_t large;  // buckets with more entries than this (after the batch-affine rounds) are summed by whole CTAs (slices)
  int R;           // batch-affine rounds: bucket regions padded to 2^R slots and halved R times (0 = XYZZ only)
};
constexpr size_t kBigMsmThreshold = 1024;  // cdl_g1_msm switches to the Pippenger path above this size
int big_msm_pick_c(size_t n);
// ba_rounds < 0: pick the number of batch-affine rounds by the mean bucket load
BigMsmDims big_msm_dims(size_t n, int c, int wfirst, int wstep, int ba_rounds = -1);
constexpr int kBigMaxBaRounds = 8;
size_t big_msm_scratch_bytes(const BigMsmDims& d);
// sorted (point, sign) entries of a launch: both GLV halves of every term in every owned window, plus
// the padding of every bucket's region to a multiple of 2^R slots (upper bound)
inline uint64_t big_msm_entries(const BigMsmDims& d) {
  return 2ull * (uint64_t)d.n * (uint64_t)d.nlocal + (uint64_t)d.nb * (((uint64_t)1 << d.R) - 1);
}
cudaError_t launch_big_msm(const G1Affine* points, const Fr* scalars, const BigMsmDims& d, int normalize,
                           void* scratch, int sm_count, G1Jac* d_out, cudaStream_t st);
void launch_big_combine(const G1Jac* in, int n, G1Jac* out, cudaStream_t st);

}  // namespace cdl
