// C ABI of libcurdle_b200.so (include/curdle_b200.h): context, device buffers,
// host-pointer entry points that mirror the gnark-crypto calls of the reference.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/curdle_b200.h"
#include "context.cuh"
#include "launch.h"

using namespace cdl;

namespace cdl {
static __global__ void k_iota(uint32_t* out, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = i;
}
}  // namespace cdl

static_assert(sizeof(Fp) == sizeof(cdl_fp), "fp layout");
static_assert(sizeof(Fr) == sizeof(cdl_fr), "fr layout");
static_assert(sizeof(G1Affine) == sizeof(cdl_g1_affine), "affine layout");
static_assert(sizeof(G1Jac) == sizeof(cdl_g1_jac), "jac layout");

extern "C" {

uint32_t cdl_abi_version(void) { return (1u << 16) | 4u; }

int32_t cdl_create(int device, cdl_ctx** out) {
  if (!out) return CDL_ERR_INVALID_ARG;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) return CDL_ERR_NO_DEVICE;
  if (device < 0 || device >= count) return CDL_ERR_INVALID_ARG;
  cdl_ctx* c = new cdl_ctx();
  c->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete c;
    return CDL_ERR_CUDA;
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  c->sm_count = prop.multiProcessorCount;
  c->clock_khz = prop.clockRate;
  c->name = prop.name;
  cudaEventCreate(&c->ev0);
  cudaEventCreate(&c->ev1);
  cudaEventCreate(&c->base_ev);
  cudaEventRecord(c->base_ev, c->stream);
  cudaStreamSynchronize(c->stream);
  // the small-MSM kernel stages up to ~6000 terms in shared memory
  msm_small_init();
  if (const char* e = getenv("CDL_MSM_C")) c->msm_c_override = atoi(e);
  if (const char* e = getenv("CDL_LANES")) { int v = atoi(e); if (v >= 1 && v <= 16) c->n_lanes = v; }
  *out = c;
  return CDL_OK;
}

// lane context `i` of a root context (created on first use; caller holds the root's mutex)
cdl_ctx* cdl_lane_(cdl_ctx* root, size_t i) {
  while (root->lanes.size() <= i) {
    cdl_ctx* l = new cdl_ctx();
    l->device = root->device;
    l->sm_count = root->sm_count;
    l->clock_khz = root->clock_khz;
    l->name = root->name;
    l->parent = root;
    l->n_lanes = 1;
    l->msm_c_override = root->msm_c_override;
    if (cudaStreamCreateWithFlags(&l->stream, cudaStreamNonBlocking) != cudaSuccess) { delete l; return nullptr; }
    cudaEventCreate(&l->ev0);
    cudaEventCreate(&l->ev1);
    root->lanes.push_back(l);
  }
  return root->lanes[i];
}

int32_t cdl_set_lanes(cdl_ctx* c, int32_t lanes) {
  if (!c || lanes < 1 || lanes > 16) return CDL_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  c->n_lanes = lanes;
  return CDL_OK;
}

void cdl_engine_free_(void* engine);  // capi_protocol.cu
int32_t cdl_big_msm_host_(cdl_ctx* c, const cdl_g1_affine* points, const cdl_fr* scalars, size_t n, cdl_g1_jac* out);  // capi_msm.cu

void cdl_destroy(cdl_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cdl_comm_destroy(c);
  for (cdl_ctx* l : c->lanes) cdl_destroy(l);
  c->lanes.clear();
  if (c->engine) cdl_engine_free_(c->engine);
  c->free_all();
  cudaEventDestroy(c->ev0);
  cudaEventDestroy(c->ev1);
  if (c->base_ev) cudaEventDestroy(c->base_ev);
  if (c->ev_sync) cudaEventDestroy(c->ev_sync);
  cudaStreamDestroy(c->stream);
  delete c;
}

const char* cdl_last_error(cdl_ctx* c) { return c ? c->err.c_str() : "null context"; }

int32_t cdl_device_info(cdl_ctx* c, int32_t* sm_count, int32_t* clock_khz, char* name, size_t cap) {
  if (!c) return CDL_ERR_INVALID_ARG;
  if (sm_count) *sm_count = c->sm_count;
  if (clock_khz) *clock_khz = c->clock_khz;
  if (name && cap) { strncpy(name, c->name.c_str(), cap - 1); name[cap - 1] = 0; }
  return CDL_OK;
}

// --------------------------------------------------------------------- MSM
int32_t cdl_g1_msm_batch(cdl_ctx* c, const cdl_g1_affine* points, const cdl_fr* scalars,
                         const uint32_t* offsets, size_t k, cdl_g1_affine* out) {
  if (!c || !offsets || !out || (k && offsets[k] && (!points || !scalars))) return CDL_ERR_INVALID_ARG;
  if (k == 0) return CDL_OK;
  std::lock_guard<std::mutex> lk(c->mu);
  CDL_CUDA(c, cudaSetDevice(c->device));
  size_t total = offsets[k];
  size_t max_terms = 0;
  std::vector<MsmTask> tasks(k);
  for (size_t j = 0; j < k; j++) {
    if (offsets[j + 1] < offsets[j]) return c->fail(CDL_ERR_INVALID_ARG, "msm_batch: offsets not monotone");
    tasks[j].term_off = offsets[j];
    tasks[j].term_cnt = offsets[j + 1] - offsets[j];
    tasks[j].out_idx = (uint32_t)j;
    tasks[j].pad = 0;
    if (tasks[j].term_cnt > max_terms) max_terms = tasks[j].term_cnt;
  }
  G1Affine* d_pts = (G1Affine*)c->buf(0, (total + 1) * sizeof(G1Affine));
  Fr* d_sc = (Fr*)c->buf(1, (total + 1) * sizeof(Fr));
  uint32_t* d_idx = (uint32_t*)c->buf(2, (total + 1) * sizeof(uint32_t));
  MsmTask* d_tasks = (MsmTask*)c->buf(3, k * sizeof(MsmTask));
  G1Affine* d_out = (G1Affine*)c->buf(4, k * sizeof(G1Affine));
  if (!d_pts || !d_sc || !d_idx || !d_tasks || !d_out) return c->fail(CDL_ERR_CUDA, "device allocation failed");
  if (total) {
    CDL_CUDA(c, cudaMemcpyAsync(d_pts, points, total * sizeof(G1Affine), cudaMemcpyHostToDevice, c->stream));
    CDL_CUDA(c, cudaMemcpyAsync(d_sc, scalars, total * sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
    k_iota<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(d_idx, (uint32_t)total);
  }
  CDL_CUDA(c, cudaMemcpyAsync(d_tasks, tasks.data(), k * sizeof(MsmTask), cudaMemcpyHostToDevice, c->stream));
  if ((int)k >= kMsmSplitThreshold || max_terms > kMsmSplitTerms) {
    std::vector<MsmSub> subs;
    std::vector<MsmTask2> tasks2;
    msm_build_subs(tasks.data(), k, msm_tp_pick_chunk(total, c->sm_count), subs, tasks2);
    MsmSub* d_subs = (MsmSub*)c->buf(5, subs.size() * sizeof(MsmSub));
    MsmTask2* d_t2 = (MsmTask2*)c->buf(6, tasks2.size() * sizeof(MsmTask2));
    void* d_scr = c->buf(7, msm_tp_scratch_bytes(total, subs.size(), k));
    if (!d_subs || !d_t2 || !d_scr) return c->fail(CDL_ERR_CUDA, "device allocation failed");
    CDL_CUDA(c, cudaMemcpyAsync(d_subs, subs.data(), subs.size() * sizeof(MsmSub), cudaMemcpyHostToDevice, c->stream));
    CDL_CUDA(c, cudaMemcpyAsync(d_t2, tasks2.data(), tasks2.size() * sizeof(MsmTask2), cudaMemcpyHostToDevice, c->stream));
    CDL_CUDA(c, cudaStreamSynchronize(c->stream));  // subs/tasks2 are pageable host vectors
    launch_msm_tp(d_pts, d_idx, d_sc, (int)total, d_subs, (int)subs.size(), d_t2, (int)k, d_out, nullptr, d_scr, c->stream);
  } else {
    launch_msm_small(d_pts, d_idx, d_sc, d_tasks, (int)k, max_terms, d_out, nullptr, c->stream);
  }
  CDL_CUDA(c, cudaGetLastError());
  CDL_CUDA(c, cudaMemcpyAsync(out, d_out, k * sizeof(G1Affine), cudaMemcpyDeviceToHost, c->stream));
  CDL_CUDA(c, cudaStreamSynchronize(c->stream));
  return CDL_OK;
}

// Batched MSM over a DEVICE-RESIDENT pool of bases: term t of MSM j is scalars[t] * d_pool[idx[t]]
// (bit 31 of idx[t] negates the base), t in [offsets[j], offsets[j+1]).  Indices, scalars and offsets are
// host arrays (what a Go-hosted orchestration computes between rounds); the bases - e.g. the folded
// vectors of innerproductargument.go:100-172, kept in HBM with cdl_g1_fold_device - never leave the
// device.  Result j goes to d_pool[out_slot[j]] (when out_slot != NULL; later MSMs / folds can use
// it as a base), to out[j] in gnark's affine layout (when out != NULL) and to out48 + 48*j as the
// compressed encoding the transcript hashes (when out48 != NULL).
int32_t cdl_g1_msm_batch_device(cdl_ctx* c, cdl_g1_affine* d_pool, const uint32_t* idx, const cdl_fr* scalars,
                                const uint32_t* offsets, size_t k, const uint32_t* out_slot, cdl_g1_affine* out,
                                uint8_t* out48) {
  if (!c || !d_pool || !offsets || (k && offsets[k] && (!idx || !scalars))) return CDL_ERR_INVALID_ARG;
  if (k == 0) return CDL_OK;
  std::lock_guard<std::mutex> lk(c->mu);
  CDL_CUDA(c, cudaSetDevice(c->device));
  const size_t total = offsets[k];
  size_t max_terms = 0;
  std::vector<MsmTask> tasks(k);
  for (size_t j = 0; j < k; j++) {
    if (offsets[j + 1] < offsets[j]) return c->fail(CDL_ERR_INVALID_ARG, "msm_batch_device: offsets not monotone");
    tasks[j].term_off = offsets[j];
    tasks[j].term_cnt = offsets[j + 1] - offsets[j];
    tasks[j].out_idx = (uint32_t)j;
    tasks[j].pad = 0;
    if (tasks[j].term_cnt > max_terms) max_terms = tasks[j].term_cnt;
  }
  // pool indices are 30 bits: bit 31 negates, bit 30 is the kernels' own flag
  for (size_t t = 0; t < total; t++)
    if (idx[t] & 0x40000000u) return c->fail(CDL_ERR_INVALID_ARG, "msm_batch_device: pool index of term %zu exceeds 2^30 - 1", t);
  Fr* d_sc = (Fr*)c->buf(1, (total + 1) * sizeof(Fr));
  uint32_t* d_idx = (uint32_t*)c->buf(2, (total + 1) * sizeof(uint32_t));
  MsmTask* d_tasks = (MsmTask*)c->buf(3, k * sizeof(MsmTask));
  G1Affine* d_out = (G1Affine*)c->buf(4, k * sizeof(G1Affine));
  uint8_t* d_c48 = (uint8_t*)c->buf(0, k * 48);
  if (!d_sc || !d_idx || !d_tasks || !d_out || !d_c48) return c->fail(CDL_ERR_CUDA, "device allocation failed");
  if (total) {
    CDL_CUDA(c, cudaMemcpyAsync(d_sc, scalars, total * sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
    CDL_CUDA(c, cudaMemcpyAsync(d_idx, idx, total * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  }
  CDL_CUDA(c, cudaMemcpyAsync(d_tasks, tasks.data(), k * sizeof(MsmTask), cudaMemcpyHostToDevice, c->stream));
  const G1Affine* pool = reinterpret_cast<const G1Affine*>(d_pool);
  if ((int)k >= kMsmSplitThreshold || max_terms > kMsmSplitTerms) {
    std::vector<MsmSub> subs;
    std::vector<MsmTask2> tasks2;
    msm_build_subs(tasks.data(), k, msm_tp_pick_chunk(total, c->sm_count), subs, tasks2);
    MsmSub* d_subs = (MsmSub*)c->buf(5, subs.size() * sizeof(MsmSub));
    MsmTask2* d_t2 = (MsmTask2*)c->buf(6, tasks2.size() * sizeof(MsmTask2));
    void* d_scr = c->buf(7, msm_tp_scratch_bytes(total, subs.size(), k));
    if (!d_subs || !d_t2 || !d_scr) return c->fail(CDL_ERR_CUDA, "device allocation failed");
    CDL_CUDA(c, cudaMemcpyAsync(d_subs, subs.data(), subs.size() * sizeof(MsmSub), cudaMemcpyHostToDevice, c->stream));
    CDL_CUDA(c, cudaMemcpyAsync(d_t2, tasks2.data(), tasks2.size() * sizeof(MsmTask2), cudaMemcpyHostToDevice, c->stream));
    CDL_CUDA(c, cudaStreamSynchronize(c->stream));  // subs/tasks2 are pageable host vectors
    launch_msm_tp(pool, d_idx, d_sc, (int)total, d_subs, (int)subs.size(), d_t2, (int)k, d_out, d_c48, d_scr, c->stream);
  } else {
    if (max_terms > kMsmMaxTerms) return c->fail(CDL_ERR_TOO_LARGE, "msm of %zu terms exceeds the small-MSM limit", max_terms);
    launch_msm_small(pool, d_idx, d_sc, d_tasks, (int)k, max_terms, d_out, d_c48, c->stream);
  }
  CDL_CUDA(c, cudaGetLastError());
  if (out_slot) {  // scatter the results into their pool slots (k small copies on the stream)
    for (size_t j = 0; j < k; j++)
      CDL_CUDA(c, cudaMemcpyAsync(reinterpret_cast<G1Affine*>(d_pool) + out_slot[j], d_out + j, sizeof(G1Affine),
                                  cudaMemcpyDeviceToDevice, c->stream));
  }
  if (out) CDL_CUDA(c, cudaMemcpyAsync(out, d_out, k * sizeof(G1Affine), cudaMemcpyDeviceToHost, c->stream));
  if (out48) CDL_CUDA(c, cudaMemcpyAsync(out48, d_c48, k * 48, cudaMemcpyDeviceToHost, c->stream));
  CDL_CUDA(c, cudaStreamSynchronize(c->stream));
  return CDL_OK;
}

int32_t cdl_g1_msm(cdl_ctx* c, const cdl_g1_affine* points, const cdl_fr* scalars, size_t n, cdl_g1_jac* out) {
  if (!c || !out || (n && (!points || !scalars))) return CDL_ERR_INVALID_ARG;
  if (n > kBigMsmThreshold) {  // Pippenger path (k_msm_big.cu); the small-MSM kernels serve Prove/Verify sizes
    std::lock_guard<std::mutex> lk(c->mu);
    CDL_CUDA(c, cudaSetDevice(c->device));
    return cdl_big_msm_host_(c, points, scalars, n, out);
  }
  uint32_t offs[2] = {0, (uint32_t)n};
  cdl_g1_affine a;
  int32_t rc = cdl_g1_msm_batch(c, points, scalars, offs, 1, &a);
  if (rc != CDL_OK) return rc;
  // lift to gnark's G1Jac: (x, y, 1) or (1, 1, 0) for infinity
  bool inf = true;
  for (int i = 0; i < 6; i++) inf = inf && a.x.l[i] == 0 && a.y.l[i] == 0;
  static const uint64_t one[6] = {0x760900000002fffdull, 0xebf4000bc40c0002ull, 0x5f48985753c758baull,
                                  0x77ce585370525745ull, 0x5c071a97a256ec6dull, 0x15f65ec3fa80e493ull};
  if (inf) {
    memcpy(out->x.l, one, 48);
    memcpy(out->y.l, one, 48);
    memset(out->z.l, 0, 48);
  } else {
    out->x = a.x;
    out->y = a.y;
    memcpy(out->z.l, one, 48);
  }
  return CDL_OK;
}

int32_t cdl_g1_sum_affine(cdl_ctx* c, const cdl_g1_affine* in, size_t n, cdl_g1_affine* out) {
  if (!c || !out || (n && !in)) return CDL_ERR_INVALID_ARG;
  static const uint64_t fr_one[4] = {0x00000001fffffffeull, 0x5884b7fa00034802ull, 0x998c4fefecbc4ff5ull, 0x1824b159acc5056full};
  std::vector<cdl_fr> ones(n);
  for (size_t i = 0; i < n; i++) memcpy(ones[i].l, fr_one, 32);
  if (n > kBigMsmThreshold) {
    cdl_g1_jac j;
    int32_t rc = cdl_g1_msm(c, in, ones.data(), n, &j);
    if (rc != CDL_OK) return rc;
    bool inf = true;
    for (int i = 0; i < 6; i++) inf = inf && j.z.l[i] == 0;
    if (inf) memset(out, 0, sizeof *out);
    else { out->x = j.x; out->y = j.y; }  // normalised: Z = 1
    return CDL_OK;
  }
  uint32_t offs[2] = {0, (uint32_t)n};
  return cdl_g1_msm_batch(c, in, ones.data(), offs, 1, out);
}

// --------------------------------------------------------------------- elementwise
static int32_t scalar_mul_impl(cdl_ctx* c, const cdl_g1_affine* in, const cdl_fr* s, size_t n, size_t stride,
                               const cdl_g1_affine* addend, cdl_g1_affine* out) {
  if (n == 0) return CDL_OK;
  std::lock_guard<std::mutex> lk(c->mu);
  CDL_CUDA(c, cudaSetDevice(c->device));
  size_t ns = stride ? n : 1;
  G1Affine* d_in = (G1Affine*)c->buf(0, n * sizeof(G1Affine));
  Fr* d_s = (Fr*)c->buf(1, ns * sizeof(Fr));
  G1Affine* d_add = addend ? (G1Affine*)c->buf(2, n * sizeof(G1Affine)) : nullptr;
  G1Affine* d_out = (G1Affine*)c->buf(4, n * sizeof(G1Affine));
  if (!d_in || !d_s || !d_out || (addend && !d_add)) return c->fail(CDL_ERR_CUDA, "device allocation failed");
  CDL_CUDA(c, cudaMemcpyAsync(d_in, in, n * sizeof(G1Affine), cudaMemcpyHostToDevice, c->stream));
  CDL_CUDA(c, cudaMemcpyAsync(d_s, s, ns * sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
  if (addend) CDL_CUDA(c, cudaMemcpyAsync(d_add, addend, n * sizeof(G1Affine), cudaMemcpyHostToDevice, c->stream));
  launch_scalar_mul(d_in, d_s, stride ? 1 : 0, d_add, d_out, (int)n, c->stream);
  CDL_CUDA(c, cudaGetLastError());
  CDL_CUDA(c, cudaMemcpyAsync(out, d_out, n * sizeof(G1Affine), cudaMemcpyDeviceToHost, c->stream));
  CDL_CUDA(c, cudaStreamSynchronize(c->stream));
  return CDL_OK;
}

int32_t cdl_g1_scalar_mul_affine(cdl_ctx* c, const cdl_g1_affine* in, const cdl_fr* s, size_t n,
                                 size_t scalar_stride, cdl_g1_affine* out) {
  if (!c || (n && (!in || !s || !out)) || scalar_stride > 1) return CDL_ERR_INVALID_ARG;
  return scalar_mul_impl(c, in, s, n, scalar_stride, nullptr, out);
}

int32_t cdl_g1_fold(cdl_ctx* c, cdl_g1_affine* L, const cdl_g1_affine* R, const cdl_fr* x, size_t n) {
  if (!c || (n && (!L || !R || !x))) return CDL_ERR_INVALID_ARG;
  return scalar_mul_impl(c, R, x, n, 0, L, L);
}

int32_t cdl_g1_batch_to_affine(cdl_ctx* c, const cdl_g1_jac* in, size_t n, cdl_g1_affine* out) {
  if (!c || (n && (!in || !out))) return CDL_ERR_INVALID_ARG;
  if (n == 0) return CDL_OK;
  std::lock_guard<std::mutex> lk(c->mu);
  CDL_CUDA(c, cudaSetDevice(c->device));
  G1Jac* d_in = (G1Jac*)c->buf(0, n * sizeof(G1Jac));
  G1Affine* d_out = (G1Affine*)c->buf(4, n * sizeof(G1Affine));
  if (!d_in || !d_out) return c->fail(CDL_ERR_CUDA, "device allocation failed");
  CDL_CUDA(c, cudaMemcpyAsync(d_in, in, n * sizeof(G1Jac), cudaMemcpyHostToDevice, c->stream));
  launch_jac_to_affine(d_in, d_out, (int)n, c->stream);
  CDL_CUDA(c, cudaGetLastError());
  CDL_CUDA(c, cudaMemcpyAsync(out, d_out, n * sizeof(G1Affine), cudaMemcpyDeviceToHost, c->stream));
  CDL_CUDA(c, cudaStreamSynchronize(c->stream));
  return CDL_OK;
}

// --------------------------------------------------------------------- codecs
int32_t cdl_g1_compress(cdl_ctx* c, const cdl_g1_affine* in, size_t n, uint8_t* out48) {
  if (!c || (n && (!in || !out48))) return CDL_ERR_INVALID_ARG;
  if (n == 0) return CDL_OK;
  std::lock_guard<std::mutex> lk(c->mu);
  CDL_CUDA(c, cudaSetDevice(c->device));
  G1Affine* d_in = (G1Affine*)c->buf(0, n * sizeof(G1Affine));
  uint8_t* d_out = (uint8_t*)c->buf(4, n * 48);
  if (!d_in || !d_out) return c->fail(CDL_ERR_CUDA, "device allocation failed");
  CDL_CUDA(c, cudaMemcpyAsync(d_in, in, n * sizeof(G1Affine), cudaMemcpyHostToDevice, c->stream));
  launch_compress(d_in, d_out, (int)n, c->stream);
  CDL_CUDA(c, cudaGetLastError());
  CDL_CUDA(c, cudaMemcpyAsync(out48, d_out, n * 48, cudaMemcpyDeviceToHost, c->stream));
  CDL_CUDA(c, cudaStreamSynchronize(c->stream));
  return CDL_OK;
}

int32_t cdl_g1_decompress(cdl_ctx* c, const uint8_t* in48, size_t n, cdl_g1_affine* out, uint8_t* status) {
  if (!c || (n && (!in48 || !out || !status))) return CDL_ERR_INVALID_ARG;
  if (n == 0) return CDL_OK;
  std::lock_guard<std::mutex> lk(c->mu);
  CDL_CUDA(c, cudaSetDevice(c->device));
  uint8_t* d_in = (uint8_t*)c->buf(0, n * 48);
  G1Affine* d_out = (G1Affine*)c->buf(4, n * sizeof(G1Affine));
  uint8_t* d_st = (uint8_t*)c->buf(3, n);
  if (!d_in || !d_out || !d_st) return c->fail(CDL_ERR_CUDA, "device allocation failed");
  CDL_CUDA(c, cudaMemcpyAsync(d_in, in48, n * 48, cudaMemcpyHostToDevice, c->stream));
  launch_decompress(d_in, d_out, d_st, (int)n, c->stream);
  CDL_CUDA(c, cudaGetLastError());
  CDL_CUDA(c, cudaMemcpyAsync(out, d_out, n * sizeof(G1Affine), cudaMemcpyDeviceToHost, c->stream));
  CDL_CUDA(c, cudaMemcpyAsync(status, d_st, n, cudaMemcpyDeviceToHost, c->stream));
  CDL_CUDA(c, cudaStreamSynchronize(c->stream));
  for (size_t i = 0; i < n; i++)
    if (status[i]) return c->fail(CDL_ERR_DECODE, "g1 decompress: point %zu rejected (reason %u)", i, (unsigned)status[i]);
  return CDL_OK;
}

// --------------------------------------------------------------------- diagnostics
int32_t cdl_fp_mul(cdl_ctx* c, const cdl_fp* a, const cdl_fp* b, size_t n, cdl_fp* out) {
  if (!c || (n && (!a || !b || !out))) return CDL_ERR_INVALID_ARG;
  if (n == 0) return CDL_OK;
  std::lock_guard<std::mutex> lk(c->mu);
  CDL_CUDA(c, cudaSetDevice(c->device));
  Fp* d_a = (Fp*)c->buf(0, n * sizeof(Fp));
  Fp* d_b = (Fp*)c->buf(1, n * sizeof(Fp));
  Fp* d_o = (Fp*)c->buf(4, n * sizeof(Fp));
  if (!d_a || !d_b || !d_o) return c->fail(CDL_ERR_CUDA, "device allocation failed");
  CDL_CUDA(c, cudaMemcpyAsync(d_a, a, n * sizeof(Fp), cudaMemcpyHostToDevice, c->stream));
  CDL_CUDA(c, cudaMemcpyAsync(d_b, b, n * sizeof(Fp), cudaMemcpyHostToDevice, c->stream));
  launch_fp_mul(d_a, d_b, d_o, (int)n, c->stream);
  CDL_CUDA(c, cudaGetLastError());
  CDL_CUDA(c, cudaMemcpyAsync(out, d_o, n * sizeof(Fp), cudaMemcpyDeviceToHost, c->stream));
  CDL_CUDA(c, cudaStreamSynchronize(c->stream));
  return CDL_OK;
}

int32_t cdl_int_peak(cdl_ctx* c, int kind, int iters, double* ops_per_s, double* ms_out) {
  return cdl_int_peak_cfg(c, kind, iters, kind >= 2 ? 2 : 8, 256, ops_per_s, ms_out);
}

int32_t cdl_int_peak_cfg(cdl_ctx* c, int kind, int iters, int blocks_per_sm, int tpb, double* ops_per_s,
                         double* ms_out) {
  if (!c || iters <= 0 || kind < 0 || kind > 4 || blocks_per_sm < 1 || blocks_per_sm > 32 || tpb < 32 || tpb > 1024)
    return CDL_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  CDL_CUDA(c, cudaSetDevice(c->device));
  const int blocks = c->sm_count * blocks_per_sm;
  void* d = c->buf(4, (size_t)blocks * tpb * sizeof(Fp));
  if (!d) return c->fail(CDL_ERR_CUDA, "device allocation failed");
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {  // rep 0 is the warm-up
    CDL_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    launch_peak(kind, d, blocks, tpb, iters, 12345u + rep, c->stream);
    CDL_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    CDL_CUDA(c, cudaStreamSynchronize(c->stream));
    CDL_CUDA(c, cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    if (rep > 0 && ms < best) best = ms;
  }
  double ops = (double)blocks * tpb * (double)iters * (kind == 2 ? 2.0 : kind == 3 ? 4.0 : kind == 4 ? 6.0 : 64.0);
  if (ops_per_s) *ops_per_s = ops / (best * 1e-3);
  if (ms_out) *ms_out = best;
  return CDL_OK;
}

}  // extern "C"
