// Library context: one CUDA device, one stream, grow-only device scratch buffers.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/curdle_b200.h"

struct cdl_ctx {
  int device = 0;
  int sm_count = 0;
  int clock_khz = 0;
  std::string name;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_sync = nullptr;  // cudaEventBlockingSync: waiting host threads sleep instead of spinning
  // Wait for everything queued on the stream.  Lanes and ranks share the host cores with the
  // Fiat-Shamir work, so a waiting thread must not burn one (cudaStreamSynchronize spins).
  // A call that serves one proof at a time is a chain of ~50 dependent stages: there the wake-up latency of
  // a sleeping wait (tens of microseconds per stage) is on the critical path and the thread spins instead.
  bool spin_wait = false;
  cudaError_t sync_stream() {
    if (spin_wait) return cudaStreamSynchronize(stream);
    if (!ev_sync && cudaEventCreateWithFlags(&ev_sync, cudaEventBlockingSync | cudaEventDisableTiming) != cudaSuccess)
      return cudaStreamSynchronize(stream);
    cudaError_t e = cudaEventRecord(ev_sync, stream);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(ev_sync);
  }
  std::mutex mu;
  std::string err;
  void* engine = nullptr;  // cdlh::Engine, created on first protocol-level call
  // multi-GPU communicator (capi_msm.cu): NCCL is dlopen'ed on first use
  void* nccl_comm = nullptr;
  int comm_rank = 0, comm_world = 1;
  int msm_c_override = 0;  // CDL_MSM_C / cdl_set_msm_window: 0 = pick by size
  int fixed_min_batch = -1;  // cdl_set_fixed_base_min_batch: instances per (lane) call from which the CRS fixed-base tables are used; -1 = CDL_FIXED_BASE_MINB or 64, 0 = never
  int msm_ba_override = -1;  // cdl_set_msm_batch_affine: batch-affine rounds of the large MSM, -1 = pick by bucket load
  // Batched protocol calls are cut into up to n_lanes sub-batches that advance
  // concurrently, each on its own lane context (same device, own stream, engine,
  // staging): one lane's host-side Fiat-Shamir work overlaps the other lanes'
  // kernels, and their latency-bound kernels overlap each other.
  cdl_ctx* parent = nullptr;
  std::vector<cdl_ctx*> lanes;
  int n_lanes = 4;               // CDL_LANES / cdl_set_lanes
  cudaEvent_t base_ev = nullptr; // origin of the kernel-interval timeline (busy-time accounting)
  cdl_ctx* root() { return parent ? parent : this; }
  static constexpr int kSlots = 8;
  void* slot[kSlots] = {};
  size_t cap[kSlots] = {};

  // grow-only scratch buffer `i` of at least `bytes`
  void* buf(int i, size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (cap[i] >= bytes) return slot[i];
    if (slot[i]) { cudaStreamSynchronize(stream); cudaFree(slot[i]); slot[i] = nullptr; cap[i] = 0; }
    size_t want = bytes + bytes / 2;
    if (cudaMalloc(&slot[i], want) != cudaSuccess) { slot[i] = nullptr; return nullptr; }
    cap[i] = want;
    return slot[i];
  }
  void free_all() {
    for (int i = 0; i < kSlots; i++) if (slot[i]) { cudaFree(slot[i]); slot[i] = nullptr; cap[i] = 0; }
  }
  int32_t fail(int32_t code, const char* fmt, ...) {
    char tmp[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tmp, sizeof tmp, fmt, ap);
    va_end(ap);
    err = tmp;
    return code;
  }
};

#define CDL_CUDA(ctx, call)                                                                   \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return (ctx)->fail(CDL_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                         __FILE__, __LINE__);                                                 \
  } while (0)
