// Host-callable kernel launchers; each is defined next to its kernel so that the
// kernels live in separate translation units (compiled in parallel).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include "g1.cuh"

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
constexpr size_t kMsmMaxSmem = 220 * 1024;         // dynamic shared memory opt-in for the small-MSM kernel
constexpr size_t kMsmMaxTerms = kMsmMaxSmem / 36;  // 36 B of staging per term

void launch_compress(const G1Affine* in, uint8_t* out, int n, cudaStream_t s);
void launch_decompress(const uint8_t* in, G1Affine* out, uint8_t* status, int n, cudaStream_t s);
void launch_fp_mul(const Fp* a, const Fp* b, Fp* out, int n, cudaStream_t s);
void launch_peak(int kind, void* out, int blocks, int tpb, int iters, uint32_t seed, cudaStream_t s);
void launch_scalar_mul(const G1Affine* P, const Fr* s, int stride, const G1Affine* L, G1Affine* out, int n,
                       cudaStream_t st);
void launch_jac_to_affine(const G1Jac* in, G1Affine* out, int n, cudaStream_t st);
void launch_elem_ops(G1Affine* pool, const ElemOp* ops, const Fr* scalars, int n, cudaStream_t st);
// indexed codecs on a pool: enc[i] <- pool[src[i]] ; pool[dst[i]] <- enc[i]
void launch_compress_idx(const G1Affine* pool, const uint32_t* src, uint8_t* out48, int n, cudaStream_t s);
void launch_decompress_idx(const uint8_t* in48, G1Affine* pool, const uint32_t* dst, uint8_t* status, int n,
                           cudaStream_t s);
// one CTA per task; out_aff / out_c48 may be null
cudaError_t msm_small_init();
// With >= kMsmSplitThreshold tasks and a scratch buffer of msm_window_scratch_bytes(ntasks)
// the throughput path (bucket kernel + combine kernel) is used; returns the number of kernels launched.
constexpr int kMsmSplitThreshold = 96;
size_t msm_window_scratch_bytes(int ntasks);
int launch_msm_small(const G1Affine* points, const uint32_t* idx, const Fr* scalars, const MsmTask* tasks,
                     int ntasks, size_t max_terms, G1Affine* out_aff, uint8_t* out_c48, void* win_scratch,
                     cudaStream_t st);

}  // namespace cdl
