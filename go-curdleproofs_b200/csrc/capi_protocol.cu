// Protocol-level C ABI: common.Rand, CRS, ShufflePermuteCommit, Prove, Verify and
// the Whisk wrappers (single and batched), on top of the batched engine.
#include <cstring>
#include <memory>
#include <algorithm>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "host/engine.hpp"

using cdlh::Engine;
using cdlh::Fr;
using cdlh::Layout;

static_assert(sizeof(Fr) == sizeof(cdl_fr), "host fr layout");

extern "C" cdl_ctx* cdl_lane_(cdl_ctx* root, size_t i);  // capi.cu

namespace {

// BLS12-381 G1 generator, affine Montgomery form (gnark bls12381.Generators())
const uint32_t kGenX[12] = {0xfd530c16u, 0x5cb38790u, 0x9976fff5u, 0x7817fc67u, 0x143ba1c1u, 0x154f95c7u,
                            0xf3d0e747u, 0xf0ae6acdu, 0x21dbf440u, 0xedce6eccu, 0x9e0bfb75u, 0x12017741u};
const uint32_t kGenY[12] = {0x0ce72271u, 0xbaac93d5u, 0x7918fd8eu, 0x8c22631au, 0x570725ceu, 0xdd595f13u,
                            0x50405194u, 0x51ac5829u, 0xad0059c0u, 0x0e1c8c3fu, 0x5008a26au, 0x0bbc3efcu};
const uint64_t kFpOne[6] = {0x760900000002fffdull, 0xebf4000bc40c0002ull, 0x5f48985753c758baull,
                            0x77ce585370525745ull, 0x5c071a97a256ec6dull, 0x15f65ec3fa80e493ull};

Engine* engine_of(cdl_ctx* c) {
  if (!c->engine) c->engine = new Engine(c);
  Engine* E = static_cast<Engine*>(c->engine);
  E->clear_fixed();  // begin_call selects the tables of its CRS; every other call runs without
  return E;
}

void affine_to_jac(cdl_g1_jac* out, const cdl_g1_affine& a) {
  bool inf = true;
  for (int i = 0; i < 6; i++) inf = inf && a.x.l[i] == 0 && a.y.l[i] == 0;
  if (inf) {
    memcpy(out->x.l, kFpOne, 48);
    memcpy(out->y.l, kFpOne, 48);
    memset(out->z.l, 0, 48);
  } else {
    out->x = a.x;
    out->y = a.y;
    memcpy(out->z.l, kFpOne, 48);
  }
}

// n draws a_i from r, points a_i * G (rand.go:72-95), written to pool[dst..]
int32_t rand_points_to_pool(Engine* E, cdl_rand* r, uint32_t gen_slot, uint32_t dst, size_t n) {
  if (!n) return CDL_OK;
  std::vector<Fr> sc(n);
  r->r.get_frs(sc.data(), n);
  std::vector<cdl::ElemOp> ops(n);
  for (size_t i = 0; i < n; i++) ops[i] = cdl::ElemOp{gen_slot, cdl::kNoPoint, (uint32_t)(dst + i), (uint32_t)i};
  return E->run_elem(ops, sc);
}

int32_t put_generator(Engine* E, uint32_t slot) {
  cdl_g1_affine g;
  memcpy(g.x.l, kGenX, 48);
  memcpy(g.y.l, kGenY, 48);
  return E->upload_points(slot, &g, 1);
}

// finish a CRS object from pool[0 .. ell+9): copies the image to its own device buffer, caches encodings
int32_t crs_finish(cdl_ctx* c, Engine* E, const Layout& L, cdl_crs** out) {
  std::unique_ptr<cdl_crs> crs(new cdl_crs());
  crs->ctx = c;
  crs->ell = L.ell;
  int32_t rc = E->set_infinity(L.INF, 1);
  if (rc) return rc;
  if (cudaMalloc(&crs->d_points, (size_t)L.crs_size * sizeof(cdl::G1Affine)) != cudaSuccess)
    return c->fail(CDL_ERR_CUDA, "crs allocation failed");
  std::vector<uint32_t> src(L.crs_size);
  for (uint32_t i = 0; i < L.crs_size; i++) src[i] = i;
  if ((rc = E->compress(src, crs->enc))) { cudaFree(crs->d_points); return rc; }
  std::vector<cdl_g1_affine> tmp(L.crs_size);
  if ((rc = E->download_points(0, tmp.data(), L.crs_size))) { cudaFree(crs->d_points); return rc; }
  if (cudaMemcpy(crs->d_points, tmp.data(), (size_t)L.crs_size * sizeof(cdl::G1Affine), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaFree(crs->d_points);
    return c->fail(CDL_ERR_CUDA, "crs upload failed");
  }
  *out = crs.release();
  return CDL_OK;
}

struct Prep {
  Engine* E;
  Layout L;
};

// common prologue of the protocol calls
int32_t begin_call(cdl_ctx* c, const cdl_crs* crs, size_t B, Engine** E, Layout* L) {
  if (crs->ctx != c->root()) return c->fail(CDL_ERR_INVALID_ARG, "crs belongs to another context");
  CDL_CUDA(c, cudaSetDevice(c->device));
  *E = engine_of(c);
  *L = Layout(crs->ell);
  c->spin_wait = B < 8 && c->parent == nullptr;  // latency regime: see cdl_ctx::sync_stream
  int32_t rc = (*E)->ensure_pool((size_t)L->crs_size + B * (size_t)L->inst_size);
  if (rc) return rc;
  (*E)->select_fixed(*L, crs, B);
  return (*E)->load_crs(*L, crs);
}

constexpr size_t kMinLaneBatch = 32;  // below this a sub-batch no longer fills its launches

// number of lanes a batch of B instances is cut into
size_t lanes_for(cdl_ctx* c, size_t B) {
  if (c->parent || c->n_lanes <= 1) return 1;
  return std::max<size_t>(1, std::min<size_t>((size_t)c->n_lanes, B / kMinLaneBatch));
}

// runs fn(lane ctx, lo, hi) for contiguous sub-batches [lo, hi) concurrently, one host thread per lane
int32_t run_lanes(cdl_ctx* c, size_t B, size_t G, const std::function<int32_t(cdl_ctx*, size_t, size_t)>& fn) {
  std::lock_guard<std::mutex> lk(c->mu);
  CDL_CUDA(c, cudaSetDevice(c->device));
  std::vector<cdl_ctx*> lane(G);
  for (size_t g = 0; g < G; g++)
    if (!(lane[g] = cdl_lane_(c, g))) return c->fail(CDL_ERR_CUDA, "lane context creation failed");
  std::vector<int32_t> rc(G, CDL_OK);
  std::vector<std::thread> th;
  for (size_t g = 0; g < G; g++) {
    size_t lo = B * g / G, hi = B * (g + 1) / G;
    th.emplace_back([&, g, lo, hi] { rc[g] = fn(lane[g], lo, hi); });
  }
  for (auto& t : th) t.join();
  for (size_t g = 0; g < G; g++)
    if (rc[g] != CDL_OK) return c->fail(rc[g], "%s", lane[g]->err.c_str());
  return CDL_OK;
}

template <class F>
void for_each_engine(cdl_ctx* c, F f) {
  if (c->engine) f(static_cast<Engine*>(c->engine));
  for (cdl_ctx* l : c->lanes)
    if (l->engine) f(static_cast<Engine*>(l->engine));
}

}  // namespace

extern "C" {

void cdl_engine_free_(void* engine) { delete static_cast<Engine*>(engine); }

uint64_t cdl_launch_count(cdl_ctx* c) {
  if (!c) return 0;
  std::lock_guard<std::mutex> lk(c->mu);
  uint64_t n = 0;
  for_each_engine(c, [&](Engine* E) { n += E->launches; });
  return n;
}

int32_t cdl_engine_stats(cdl_ctx* c, uint64_t* launches, double* ms, double* modmul, double* bytes, int reset) {
  if (!c) return CDL_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  for (int i = 0; i < 4; i++) {
    if (launches) launches[i] = 0;
    if (ms) ms[i] = 0;
    if (modmul) modmul[i] = 0;
    if (bytes) bytes[i] = 0;
  }
  for_each_engine(c, [&](Engine* E) {
    for (int i = 0; i < 4; i++) {
      if (launches) launches[i] += E->stats.n[i];
      if (ms) ms[i] += E->stats.ms[i];
      if (modmul) modmul[i] += E->stats.modmul[i];
      if (bytes) bytes[i] += E->stats.bytes[i];
    }
    if (reset) { E->stats = Engine::Stats(); E->intervals.clear(); }
  });
  return CDL_OK;
}

int32_t cdl_engine_busy_ms(cdl_ctx* c, double* busy_ms) {
  if (!c || !busy_ms) return CDL_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  std::vector<std::pair<float, float>> iv;
  for_each_engine(c, [&](Engine* E) { iv.insert(iv.end(), E->intervals.begin(), E->intervals.end()); });
  std::sort(iv.begin(), iv.end());
  double busy = 0, cur_lo = 0, cur_hi = -1;
  for (auto& p : iv) {
    if (cur_hi < 0 || p.first > cur_hi) {
      if (cur_hi >= 0) busy += cur_hi - cur_lo;
      cur_lo = p.first;
      cur_hi = p.second;
    } else if (p.second > cur_hi) {
      cur_hi = p.second;
    }
  }
  if (cur_hi >= 0) busy += cur_hi - cur_lo;
  *busy_ms = busy;
  return CDL_OK;
}

int32_t cdl_host_selftest(uint8_t* out32, const cdl_fr* a, const cdl_fr* b, cdl_fr* fr_out) {
  if (!out32 || !a || !b || !fr_out) return CDL_ERR_INVALID_ARG;
  cdlh::Transcript t("test protocol");
  t.append_message("some label", reinterpret_cast<const uint8_t*>("some data"), 9);
  t.challenge_bytes("challenge", out32, 32);
  Fr x, y;
  memcpy(&x, a, 32);
  memcpy(&y, b, 32);
  Fr t1 = cdlh::fr_sub(cdlh::fr_add(cdlh::fr_mul(x, y), x), y);
  Fr r = cdlh::fr_mul(cdlh::fr_inv(t1), cdlh::fr_pow_u64(x, 5));
  memcpy(fr_out, &r, 32);
  // the batched inversion must agree with the single one and keep zeros
  std::vector<Fr> bi = cdlh::fr_batch_inv({x, t1, cdlh::FR_ZERO, y, t1});
  if (!cdlh::fr_eq(bi[1], cdlh::fr_inv(t1)) || !cdlh::fr_eq(bi[4], bi[1]) || !cdlh::fr_is_zero(bi[2]) ||
      !cdlh::fr_eq(bi[0], cdlh::fr_inv(x)) || !cdlh::fr_eq(bi[3], cdlh::fr_inv(y)))
    return CDL_ERR_INTERNAL;
  if (!cdlh::fr_is_zero(t1) && !cdlh::fr_eq(cdlh::fr_mul(t1, cdlh::fr_inv(t1)), cdlh::FR_ONE)) return CDL_ERR_INTERNAL;
  return CDL_OK;
}

int32_t cdl_host_selftest_fibers(uint32_t n, uint32_t rounds, uint8_t* digest32, int32_t* used_x8) {
  if (!digest32 || !used_x8 || n == 0 || n > (1u << 16)) return CDL_ERR_INVALID_ARG;
  auto work = [&](size_t i, uint8_t* out) {  // out: rounds * 32 + 32 bytes
    cdlh::Transcript t("curdleproofs");
    cdlh::Rand rnd(1000 + i);
    uint8_t msg[200];
    for (uint32_t r = 0; r < rounds; r++) {
      size_t len = 1 + (i * 7 + r * 13) % sizeof msg;
      for (size_t k = 0; k < len; k++) msg[k] = (uint8_t)(i + 31 * r + k);
      t.append_message("selftest_msg", msg, len);
      Fr c = t.challenge("selftest_challenge");
      Fr x = rnd.get_fr();
      t.append_scalar("selftest_rand", x);
      cdlh::fr_to_bytes_be(out + 32 * r, cdlh::fr_add(c, x));
    }
    t.challenge_bytes("selftest_final", out + 32 * rounds, 32);
  };
  const size_t per = (size_t)rounds * 32 + 32;
  std::vector<uint8_t> plain(n * per), fib(n * per);
  for (size_t i = 0; i < n; i++) work(i, plain.data() + i * per);
  *used_x8 = cdlh::fibers_available() ? 1 : 0;
  cdlh::ThreadPool pool(3);
  std::function<void(size_t)> fn([&](size_t i) { work(i, fib.data() + i * per); });
  pool.parallel_for((n + 7) / 8, std::function<void(size_t)>([&](size_t g) {
    cdlh::run_fiber_group(fn, 8 * g, n - 8 * g < 8 ? n - 8 * g : 8);
  }));
  cdlh::Shake256 h;
  h.absorb(fib.data(), fib.size());
  h.read(digest32, 32);
  return plain == fib ? CDL_OK : CDL_ERR_INTERNAL;
}

// ------------------------------------------------------------------ Rand
int32_t cdl_rand_new(uint64_t seed, cdl_rand** out) {
  if (!out) return CDL_ERR_INVALID_ARG;
  *out = new cdl_rand(seed);
  return CDL_OK;
}
void cdl_rand_free(cdl_rand* r) { delete r; }
int32_t cdl_rand_get_frs(cdl_rand* r, size_t n, cdl_fr* out) {
  if (!r || (n && !out)) return CDL_ERR_INVALID_ARG;
  r->r.get_frs(reinterpret_cast<Fr*>(out), n);
  return CDL_OK;
}
int32_t cdl_rand_generate_permutation(cdl_rand* r, size_t n, uint32_t* out) {
  if (!r || (n && !out)) return CDL_ERR_INVALID_ARG;
  std::vector<uint32_t> p = r->r.generate_permutation(n);
  memcpy(out, p.data(), n * 4);
  return CDL_OK;
}
int32_t cdl_rand_get_g1_affines(cdl_ctx* c, cdl_rand* r, size_t n, cdl_g1_affine* out) {
  if (!c || !r || (n && !out)) return CDL_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  CDL_CUDA(c, cudaSetDevice(c->device));
  Engine* E = engine_of(c);
  int32_t rc;
  if ((rc = E->ensure_pool(n + 1)) || (rc = put_generator(E, 0)) || (rc = rand_points_to_pool(E, r, 0, 1, n))) return rc;
  return E->download_points(1, out, n);
}

// ------------------------------------------------------------------ CRS
int32_t cdl_crs_generate(cdl_ctx* c, size_t ell, cdl_rand* r, cdl_crs** out) {
  if (!c || !r || !out || ell == 0 || ell > (1u << 24)) return CDL_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  CDL_CUDA(c, cudaSetDevice(c->device));
  Engine* E = engine_of(c);
  Layout L((uint32_t)ell);
  int32_t rc;
  if ((rc = E->ensure_pool(L.crs_size + 1))) return rc;
  const uint32_t gen = L.crs_size;  // scratch slot after the CRS image
  if ((rc = put_generator(E, gen))) return rc;
  // draw order: Gs (ell), Hs (4), H, Gt, Gu   (crs.go:21-40); these are contiguous in the image
  if ((rc = rand_points_to_pool(E, r, gen, L.Gs, ell + 7))) return rc;
  // Gsum = sum Gs, Hsum = sum Hs   (crs.go:41-48)
  cdlh::MsmStage st;
  st.idx.resize(ell + 4);
  st.sc.assign(ell + 4, cdlh::FR_ONE);
  for (uint32_t i = 0; i < ell + 4; i++) st.idx[i] = i;
  st.tasks.push_back(cdl::MsmTask{0, (uint32_t)ell, L.Gsum, 0});
  st.tasks.push_back(cdl::MsmTask{(uint32_t)ell, 4, L.Hsum, 0});
  if ((rc = E->run_msm(st))) return rc;
  return crs_finish(c, E, L, out);
}

int32_t cdl_crs_from_points(cdl_ctx* c, size_t ell, const cdl_g1_affine* points, cdl_crs** out) {
  if (!c || !points || !out || ell == 0 || ell > (1u << 24)) return CDL_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  CDL_CUDA(c, cudaSetDevice(c->device));
  Engine* E = engine_of(c);
  Layout L((uint32_t)ell);
  int32_t rc;
  if ((rc = E->ensure_pool(L.crs_size + 1)) || (rc = E->upload_points(0, points, ell + 9))) return rc;
  return crs_finish(c, E, L, out);
}

int32_t cdl_crs_export(cdl_ctx* c, const cdl_crs* crs, cdl_g1_affine* points) {
  if (!c || !crs || !points) return CDL_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  CDL_CUDA(c, cudaSetDevice(c->device));
  CDL_CUDA(c, cudaMemcpy(points, crs->d_points, (size_t)(crs->ell + 9) * sizeof(cdl::G1Affine), cudaMemcpyDeviceToHost));
  return CDL_OK;
}

size_t cdl_crs_ell(const cdl_crs* crs) { return crs ? crs->ell : 0; }

int32_t cdl_set_fixed_base_min_batch(cdl_ctx* c, int32_t instances) {
  if (!c || instances < -1) return CDL_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  c->root()->fixed_min_batch = instances;
  return CDL_OK;
}

void cdl_crs_free(cdl_crs* crs) {
  if (!crs) return;
  if (crs->d_points) { cudaSetDevice(crs->ctx->device); cudaFree(crs->d_points); }
  if (crs->d_fixed) { cudaSetDevice(crs->ctx->device); cudaFree(crs->d_fixed); }
  delete crs;
}

// ------------------------------------------------------------------ ShufflePermuteCommit
int32_t cdl_shuffle_permute_commit(cdl_ctx* c, const cdl_crs* crs, const cdl_g1_affine* Rs, const cdl_g1_affine* Ss,
                                   const uint32_t* perm, const cdl_fr* k, cdl_rand* r, cdl_g1_affine* Ts,
                                   cdl_g1_affine* Us, cdl_g1_jac* M, cdl_fr* rs_m) {
  if (!c || !crs || !Rs || !Ss || !perm || !k || !r || !Ts || !Us || !M || !rs_m) return CDL_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  Engine* E;
  Layout L(1);
  int32_t rc = begin_call(c, crs, 1, &E, &L);
  if (rc) return rc;
  const uint32_t ell = L.ell, base = L.base(0);
  for (uint32_t i = 0; i < ell; i++)
    if (perm[i] >= ell) return c->fail(CDL_ERR_INVALID_ARG, "permutation entry out of range");
  if ((rc = E->upload_points(base + L.Rs, Rs, ell)) || (rc = E->upload_points(base + L.Ss, Ss, ell))) return rc;
  std::vector<std::vector<uint32_t>> perms(1, std::vector<uint32_t>(perm, perm + ell));
  std::vector<Fr> ks(1);
  memcpy(&ks[0], k, 32);
  std::vector<cdl_rand*> rands(1, r);
  std::vector<std::vector<Fr>> rsm;
  if ((rc = E->shuffle_permute_commit(L, 1, perms, ks, rands, rsm))) return rc;
  cdl_g1_affine m_aff;
  if ((rc = E->download_points(base + L.Ts, Ts, ell)) || (rc = E->download_points(base + L.Us, Us, ell)) ||
      (rc = E->download_points(base + L.M, &m_aff, 1)))
    return rc;
  affine_to_jac(M, m_aff);
  memcpy(rs_m, rsm[0].data(), 4 * 32);
  return CDL_OK;
}

// ------------------------------------------------------------------ Prove
int32_t cdl_prove(cdl_ctx* c, const cdl_crs* crs, const cdl_g1_affine* Rs, const cdl_g1_affine* Ss,
                  const cdl_g1_affine* Ts, const cdl_g1_affine* Us, const cdl_g1_jac* M, const uint32_t* perm,
                  const cdl_fr* k, const cdl_fr* rs_m, cdl_rand* r, uint8_t* proof, size_t proof_cap,
                  size_t* proof_len) {
  if (!c || !crs || !Rs || !Ss || !Ts || !Us || !M || !perm || !k || !rs_m || !r || !proof || !proof_len)
    return CDL_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  Engine* E;
  Layout L(1);
  int32_t rc = begin_call(c, crs, 1, &E, &L);
  if (rc) return rc;
  const uint32_t ell = L.ell, base = L.base(0);
  for (uint32_t i = 0; i < ell; i++)
    if (perm[i] >= ell) return c->fail(CDL_ERR_INVALID_ARG, "permutation entry out of range");
  if ((rc = E->upload_points(base + L.Rs, Rs, ell)) || (rc = E->upload_points(base + L.Ss, Ss, ell)) ||
      (rc = E->upload_points(base + L.Ts, Ts, ell)) || (rc = E->upload_points(base + L.Us, Us, ell)) ||
      (rc = E->upload_jac(base + L.M, M, 1)))
    return rc;
  std::vector<std::vector<uint32_t>> perms(1, std::vector<uint32_t>(perm, perm + ell));
  std::vector<Fr> ks(1);
  memcpy(&ks[0], k, 32);
  std::vector<std::vector<Fr>> rsm(1, std::vector<Fr>(4));
  memcpy(rsm[0].data(), rs_m, 4 * 32);
  std::vector<cdl_rand*> rands(1, r);
  std::vector<std::vector<uint8_t>> proofs;
  std::vector<int32_t> status;
  std::vector<std::string> errs;
  std::vector<uint8_t> inst_enc;
  if ((rc = E->prove(L, 1, crs, perms, ks, rsm, rands, proofs, status, errs, inst_enc, false))) return rc;
  if (status[0] != CDL_OK) return c->fail(status[0], "%s", errs[0].c_str());
  *proof_len = proofs[0].size();
  if (proofs[0].size() > proof_cap) return c->fail(CDL_ERR_INVALID_ARG, "proof buffer too small: need %zu bytes", proofs[0].size());
  memcpy(proof, proofs[0].data(), proofs[0].size());
  return CDL_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ Verify helpers
namespace {

// Decode B serialized proofs (+ optional leading M) and their instance encodings into
// the pool: fills pp[b] (indices, encodings, scalars) and marks decode failures.
// inst_enc[b] must hold Rs | Ss | Ts | Us encodings (4*ell*48 B) when decode_instance.
int32_t load_proofs(Engine* E, const Layout& L, uint32_t B, const uint8_t* const* proof_ptr, const size_t* proof_len,
                    bool with_m, const uint8_t* const* m_enc, const std::vector<const uint8_t*>& inst_enc,
                    bool decode_instance, std::vector<Engine::ParsedProof>& pp) {
  const uint32_t ell = L.ell;
  pp.assign(B, Engine::ParsedProof());
  std::vector<uint8_t> enc;
  std::vector<uint32_t> dst;
  std::vector<uint32_t> owner;  // instance of each queued point
  for (uint32_t b = 0; b < B; b++) {
    Engine::ParsedProof& q = pp[b];
    q.inst_enc = inst_enc[b];
    cdlh::WireProof w;
    size_t used = 0;
    std::string e = cdlh::parse_wire_proof(proof_ptr[b], proof_len[b], with_m, w, &used);
    if (!e.empty()) { q.err = "decoding proof: " + e; continue; }
    q.sc.resize(7);
    bool sc_ok = true;
    for (int i = 0; i < 7; i++) sc_ok = sc_ok && cdlh::fr_from_bytes_be_canonical(q.sc[i], w.scalars[i]);
    if (!sc_ok) { q.err = "decoding proof: scalar is not canonical"; continue; }
    memcpy(q.lens, w.lens, sizeof q.lens);
    if (w.points.size() + (with_m ? 0 : 1) > 19 + 10 * 32) { q.err = "decoding proof: too many points"; continue; }
    const uint32_t base = L.base(b);
    uint32_t slot = base + L.PP;
    if (!with_m) {  // M comes from the caller (already in the pool at L.M)
      q.pt.push_back(base + L.M);
      q.enc.push_back(m_enc[b]);
    }
    for (const uint8_t* p : w.points) {
      q.pt.push_back(slot);
      q.enc.push_back(p);
      enc.insert(enc.end(), p, p + 48);
      dst.push_back(slot++);
      owner.push_back(b);
    }
    if (decode_instance) {
      enc.insert(enc.end(), inst_enc[b], inst_enc[b] + (size_t)4 * ell * 48);
      for (uint32_t i = 0; i < 4 * ell; i++) { dst.push_back(base + L.Rs + i); owner.push_back(b); }
    }
    q.ok = true;
  }
  std::vector<uint8_t> st;
  int32_t rc = E->decompress(enc.data(), dst, st);
  if (rc) return rc;
  for (size_t i = 0; i < st.size(); i++)
    if (st[i] && pp[owner[i]].ok) {
      pp[owner[i]].ok = false;
      pp[owner[i]].err = "decoding point: rejected encoding (reason " + std::to_string((int)st[i]) + ")";
    }
  return CDL_OK;
}

}  // namespace

extern "C" {

int32_t cdl_verify(cdl_ctx* c, const cdl_crs* crs, const uint8_t* proof, size_t proof_len,
                   const cdl_g1_affine* Rs, const cdl_g1_affine* Ss, const cdl_g1_affine* Ts,
                   const cdl_g1_affine* Us, const cdl_g1_jac* M, cdl_rand* r, int32_t* ok) {
  if (!c || !crs || !proof || !Rs || !Ss || !Ts || !Us || !M || !r || !ok) return CDL_ERR_INVALID_ARG;
  *ok = 0;
  std::lock_guard<std::mutex> lk(c->mu);
  Engine* E;
  Layout L(1);
  int32_t rc = begin_call(c, crs, 1, &E, &L);
  if (rc) return rc;
  const uint32_t ell = L.ell, base = L.base(0);
  if ((rc = E->upload_points(base + L.Rs, Rs, ell)) || (rc = E->upload_points(base + L.Ss, Ss, ell)) ||
      (rc = E->upload_points(base + L.Ts, Ts, ell)) || (rc = E->upload_points(base + L.Us, Us, ell)) ||
      (rc = E->upload_jac(base + L.M, M, 1)))
    return rc;
  std::vector<uint32_t> src(4 * ell + 1);
  for (uint32_t i = 0; i < 4 * ell; i++) src[i] = base + L.Rs + i;
  src[4 * ell] = base + L.M;
  std::vector<uint8_t> inst;
  if ((rc = E->compress(src, inst))) return rc;
  const uint8_t* m_enc = inst.data() + (size_t)4 * ell * 48;
  std::vector<const uint8_t*> inst_ptr(1, inst.data());
  std::vector<Engine::ParsedProof> pp;
  if ((rc = load_proofs(E, L, 1, &proof, &proof_len, false, &m_enc, inst_ptr, false, pp))) return rc;
  if (!pp[0].ok) return c->fail(CDL_ERR_DECODE, "%s", pp[0].err.c_str());
  std::vector<cdl_rand*> rands(1, r);
  std::vector<int32_t> verdict, status;
  std::vector<std::string> errs;
  if ((rc = E->verify(L, 1, crs, pp, rands, verdict, status, errs))) return rc;
  if (status[0] != CDL_OK) return c->fail(status[0], "%s", errs[0].c_str());
  *ok = verdict[0];
  return CDL_OK;
}

// ------------------------------------------------------------------ Whisk
int32_t cdl_whisk_generate_shuffle_proof_batch(cdl_ctx* c, const cdl_crs* crs, size_t B, const uint8_t* pre_trackers,
                                               cdl_rand* const* rands_in, uint8_t* post_trackers, uint8_t* proofs_out,
                                               size_t proof_cap, int32_t* status_out) {
  if (!c || !crs || !pre_trackers || !rands_in || !post_trackers || !proofs_out || !status_out || B == 0)
    return CDL_ERR_INVALID_ARG;
  if (size_t G = lanes_for(c, B); G > 1) {
    const size_t tb = (size_t)crs->ell * 96;
    return run_lanes(c, B, G, [&](cdl_ctx* lane, size_t lo, size_t hi) {
      return cdl_whisk_generate_shuffle_proof_batch(lane, crs, hi - lo, pre_trackers + lo * tb, rands_in + lo,
                                                    post_trackers + lo * tb, proofs_out + lo * proof_cap, proof_cap,
                                                    status_out + lo);
    });
  }
  std::lock_guard<std::mutex> lk(c->mu);
  Engine* E;
  Layout L(1);
  int32_t rc = begin_call(c, crs, B, &E, &L);
  if (rc) return rc;
  const uint32_t ell = L.ell;
  std::vector<cdl_rand*> rands(rands_in, rands_in + B);
  // whisk.go:64-71: permutation, then k
  std::vector<std::vector<uint32_t>> perms(B);
  std::vector<Fr> ks(B);
  for (size_t b = 0; b < B; b++) {
    perms[b] = rands[b]->r.generate_permutation(ell);
    ks[b] = rands[b]->r.get_fr();
  }
  // whisk.go:73-80: trackers -> points (SetBytes with subgroup check), rG -> Rs, krG -> Ss
  std::vector<uint32_t> dst((size_t)B * 2 * ell);
  for (size_t b = 0; b < B; b++)
    for (uint32_t i = 0; i < ell; i++) {
      dst[(b * ell + i) * 2] = L.base((uint32_t)b) + L.Rs + i;
      dst[(b * ell + i) * 2 + 1] = L.base((uint32_t)b) + L.Ss + i;
    }
  std::vector<uint8_t> st;
  if ((rc = E->decompress(pre_trackers, dst, st))) return rc;
  bool any_bad = false;
  for (size_t b = 0; b < B; b++) {
    status_out[b] = CDL_OK;
    for (uint32_t i = 0; i < 2 * ell; i++)
      if (st[b * 2 * ell + i]) { status_out[b] = CDL_ERR_DECODE; any_bad = true; }
  }
  std::vector<std::vector<Fr>> rsm;
  if ((rc = E->shuffle_permute_commit(L, (uint32_t)B, perms, ks, rands, rsm))) return rc;
  std::vector<std::vector<uint8_t>> proofs;
  std::vector<int32_t> status;
  std::vector<std::string> errs;
  std::vector<uint8_t> inst_enc;
  if ((rc = E->prove(L, (uint32_t)B, crs, perms, ks, rsm, rands, proofs, status, errs, inst_enc, true))) return rc;
  const size_t per_inst = (size_t)(4 * ell + 1) * 48;
  for (size_t b = 0; b < B; b++) {
    uint8_t* po = proofs_out + b * proof_cap;
    memset(po, 0, proof_cap);
    uint8_t* post = post_trackers + b * (size_t)ell * 96;
    if (status_out[b] != CDL_OK) { memset(post, 0, (size_t)ell * 96); continue; }
    if (status[b] != CDL_OK) { status_out[b] = status[b]; c->fail(status[b], "generating proof: %s", errs[b].c_str()); any_bad = true; continue; }
    const uint8_t* ie = inst_enc.data() + b * per_inst;
    if (48 + proofs[b].size() > proof_cap) { status_out[b] = CDL_ERR_INVALID_ARG; any_bad = true; continue; }
    memcpy(po, ie + (size_t)4 * ell * 48, 48);  // M   (whisk/types.go:57-61)
    memcpy(po + 48, proofs[b].data(), proofs[b].size());
    for (uint32_t i = 0; i < ell; i++) {  // NewWhiskTracker(Ts[i], Us[i])  (whisk.go:108-111)
      memcpy(post + 96 * i, ie + ((size_t)2 * ell + i) * 48, 48);
      memcpy(post + 96 * i + 48, ie + ((size_t)3 * ell + i) * 48, 48);
    }
  }
  (void)any_bad;
  return CDL_OK;
}

int32_t cdl_whisk_generate_shuffle_proof(cdl_ctx* c, const cdl_crs* crs, const uint8_t* pre_trackers, cdl_rand* r,
                                         uint8_t* post_trackers, uint8_t* proof, size_t proof_cap) {
  int32_t status = CDL_OK;
  cdl_rand* rr[1] = {r};
  int32_t rc = cdl_whisk_generate_shuffle_proof_batch(c, crs, 1, pre_trackers, rr, post_trackers, proof, proof_cap, &status);
  if (rc) return rc;
  if (status == CDL_ERR_DECODE) return c->fail(status, "getting points: a pre-shuffle tracker failed to decode");
  if (status == CDL_ERR_INVALID_ARG) return c->fail(status, "proof buffer too small");
  return status;
}

int32_t cdl_whisk_is_valid_shuffle_proof_batch(cdl_ctx* c, const cdl_crs* crs, size_t B, const uint8_t* pre_trackers,
                                               const uint8_t* post_trackers, const uint8_t* proofs, size_t proof_len,
                                               cdl_rand* const* rands_in, int32_t* ok, int32_t* status_out) {
  if (!c || !crs || !pre_trackers || !post_trackers || !proofs || !rands_in || !ok || !status_out || B == 0)
    return CDL_ERR_INVALID_ARG;
  if (size_t G = lanes_for(c, B); G > 1) {
    const size_t tb = (size_t)crs->ell * 96;
    return run_lanes(c, B, G, [&](cdl_ctx* lane, size_t lo, size_t hi) {
      return cdl_whisk_is_valid_shuffle_proof_batch(lane, crs, hi - lo, pre_trackers + lo * tb, post_trackers + lo * tb,
                                                    proofs + lo * proof_len, proof_len, rands_in + lo, ok + lo,
                                                    status_out + lo);
    });
  }
  std::lock_guard<std::mutex> lk(c->mu);
  Engine* E;
  Layout L(1);
  int32_t rc = begin_call(c, crs, B, &E, &L);
  if (rc) return rc;
  const uint32_t ell = L.ell;
  // instance encodings in transcript order Rs | Ss | Ts | Us   (whisk.go:31-44)
  std::vector<uint8_t> inst((size_t)B * 4 * ell * 48);
  std::vector<const uint8_t*> inst_ptr(B), proof_ptr(B);
  std::vector<size_t> plen(B, proof_len);
  for (size_t b = 0; b < B; b++) {
    uint8_t* ie = inst.data() + b * (size_t)4 * ell * 48;
    const uint8_t* pre = pre_trackers + b * (size_t)ell * 96;
    const uint8_t* post = post_trackers + b * (size_t)ell * 96;
    for (uint32_t i = 0; i < ell; i++) {
      memcpy(ie + (size_t)i * 48, pre + 96 * i, 48);
      memcpy(ie + ((size_t)ell + i) * 48, pre + 96 * i + 48, 48);
      memcpy(ie + ((size_t)2 * ell + i) * 48, post + 96 * i, 48);
      memcpy(ie + ((size_t)3 * ell + i) * 48, post + 96 * i + 48, 48);
    }
    inst_ptr[b] = ie;
    proof_ptr[b] = proofs + b * proof_len;
  }
  std::vector<Engine::ParsedProof> pp;
  if ((rc = load_proofs(E, L, (uint32_t)B, proof_ptr.data(), plen.data(), true, nullptr, inst_ptr, true, pp))) return rc;
  std::vector<cdl_rand*> rands(rands_in, rands_in + B);
  std::vector<int32_t> verdict, status;
  std::vector<std::string> errs;
  // decode failures are the reference's (false, err) with a decode error
  std::vector<uint8_t> decode_failed(B, 0);
  for (size_t b = 0; b < B; b++) decode_failed[b] = !pp[b].ok;
  if ((rc = E->verify(L, (uint32_t)B, crs, pp, rands, verdict, status, errs))) return rc;
  for (size_t b = 0; b < B; b++) {
    ok[b] = verdict[b];
    status_out[b] = decode_failed[b] ? CDL_ERR_DECODE : status[b];
    if (status_out[b] != CDL_OK) { ok[b] = 0; c->fail(status_out[b], "%s", errs[b].c_str()); }
  }
  return CDL_OK;
}

int32_t cdl_whisk_is_valid_shuffle_proof(cdl_ctx* c, const cdl_crs* crs, const uint8_t* pre_trackers,
                                         const uint8_t* post_trackers, size_t n_pre, size_t n_post,
                                         const uint8_t* proof, size_t proof_len, cdl_rand* r, int32_t* ok) {
  if (!c || !crs || !ok) return CDL_ERR_INVALID_ARG;
  *ok = 0;
  if (n_pre != n_post) return c->fail(CDL_ERR_PROTOCOL, "pre and post shuffle trackers must be the same length");
  if (n_pre != crs->ell) return c->fail(CDL_ERR_INVALID_ARG, "tracker count %zu does not match the CRS (ell = %u)", n_pre, crs->ell);
  int32_t status = CDL_OK;
  cdl_rand* rr[1] = {r};
  int32_t rc = cdl_whisk_is_valid_shuffle_proof_batch(c, crs, 1, pre_trackers, post_trackers, proof, proof_len, rr, ok, &status);
  if (rc) return rc;
  return status;
}


// ------------------------------------------------------------------ Whisk tracker opening proofs
// whisk.GenerateWhiskTrackerProof / IsValidWhiskTrackerProof (whisk/whisk.go:116-176): a Schnorr-style
// proof that the tracker (rG, krG) and the commitment kG share k.  Constant work per tracker (3 / 4
// scalar multiplications), so it is only worth a launch when batched over validators; the layout in
// the point pool is [G | per instance: rG krG kG A B A' B'].
namespace {
constexpr uint32_t kTpStride = 7;
const uint8_t kGenEnc[48] = {0x97, 0xf1, 0xd3, 0xa7, 0x31, 0x97, 0xd7, 0x94, 0x26, 0x95, 0x63, 0x8c, 0x4f, 0xa9, 0xac, 0x0f,
                             0xc3, 0x68, 0x8c, 0x4f, 0x97, 0x74, 0xb9, 0x05, 0xa1, 0x4e, 0x3a, 0x3f, 0x17, 0x1b, 0xac, 0x58,
                             0x6c, 0x55, 0xe8, 0x3f, 0xf9, 0x7a, 0x1a, 0xef, 0xfb, 0x3a, 0xf0, 0x0a, 0xdb, 0x22, 0xc6, 0xbb};

cdlh::Fr tracker_challenge(const uint8_t* kG, const uint8_t* krG, const uint8_t* rG, const uint8_t* A, const uint8_t* B) {
  cdlh::Transcript tr("whisk_opening_proof");
  tr.append_points("tracker_opening_proof", kG, 1);
  tr.append_points("tracker_opening_proof", kGenEnc, 1);
  tr.append_points("tracker_opening_proof", krG, 1);
  tr.append_points("tracker_opening_proof", rG, 1);
  tr.append_points("tracker_opening_proof", A, 1);
  tr.append_points("tracker_opening_proof", B, 1);
  return tr.challenge("tracker_opening_proof_challenge");
}
}  // namespace

int32_t cdl_whisk_generate_tracker_proof_batch(cdl_ctx* c, size_t B, const uint8_t* trackers, const cdl_fr* ks,
                                               cdl_rand* const* rands, uint8_t* proofs, int32_t* status_out) {
  if (!c || !trackers || !ks || !rands || !proofs || !status_out || B == 0 || B > (1u << 24)) return CDL_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  CDL_CUDA(c, cudaSetDevice(c->device));
  Engine* E = engine_of(c);
  int32_t rc;
  if ((rc = E->ensure_pool(1 + B * kTpStride)) || (rc = put_generator(E, 0))) return rc;
  // tracker.getPoints(): rG, krG with curve + subgroup checks
  std::vector<uint32_t> dst(2 * B);
  for (size_t b = 0; b < B; b++) { dst[2 * b] = (uint32_t)(1 + b * kTpStride); dst[2 * b + 1] = dst[2 * b] + 1; }
  std::vector<uint8_t> st;
  if ((rc = E->decompress(trackers, dst, st))) return rc;
  // kG = k*G, A = blinder*G, B = blinder*rG
  std::vector<Fr> sc(2 * B);
  std::vector<cdl::ElemOp> ops(3 * B);
  for (size_t b = 0; b < B; b++) {
    status_out[b] = (st[2 * b] || st[2 * b + 1]) ? CDL_ERR_DECODE : CDL_OK;
    uint32_t base = (uint32_t)(1 + b * kTpStride);
    memcpy(&sc[2 * b], &ks[b], 32);
    sc[2 * b + 1] = rands[b]->r.get_fr();  // blinder (whisk.go:157)
    ops[3 * b] = cdl::ElemOp{0, cdl::kNoPoint, base + 2, (uint32_t)(2 * b)};
    ops[3 * b + 1] = cdl::ElemOp{0, cdl::kNoPoint, base + 3, (uint32_t)(2 * b + 1)};
    ops[3 * b + 2] = cdl::ElemOp{base, cdl::kNoPoint, base + 4, (uint32_t)(2 * b + 1)};
  }
  if ((rc = E->run_elem(ops, sc))) return rc;
  std::vector<uint32_t> src(3 * B);
  for (size_t b = 0; b < B; b++)
    for (uint32_t j = 0; j < 3; j++) src[3 * b + j] = (uint32_t)(1 + b * kTpStride + 2 + j);
  std::vector<uint8_t> enc;
  if ((rc = E->compress(src, enc))) return rc;
  E->par(B, [&](size_t b) {
    uint8_t* out = proofs + 128 * b;
    if (status_out[b] != CDL_OK) { memset(out, 0, 128); return; }
    const uint8_t* kG = enc.data() + 48 * (3 * b);
    const uint8_t* A = kG + 48;
    const uint8_t* Bp = kG + 96;
    const uint8_t* rG = trackers + 96 * b;
    Fr ch = tracker_challenge(kG, rG + 48, rG, A, Bp);
    Fr s = cdlh::fr_sub(sc[2 * b + 1], cdlh::fr_mul(ch, sc[2 * b]));  // s = blinder - challenge*k
    memcpy(out, A, 48);
    memcpy(out + 48, Bp, 48);
    cdlh::fr_to_bytes_be(out + 96, s);
  });
  return CDL_OK;
}

int32_t cdl_whisk_is_valid_tracker_proof_batch(cdl_ctx* c, size_t B, const uint8_t* trackers, const uint8_t* k_comms,
                                               const uint8_t* proofs, int32_t* ok, int32_t* status_out) {
  if (!c || !trackers || !k_comms || !proofs || !ok || !status_out || B == 0 || B > (1u << 24)) return CDL_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(c->mu);
  CDL_CUDA(c, cudaSetDevice(c->device));
  Engine* E = engine_of(c);
  int32_t rc;
  if ((rc = E->ensure_pool(1 + B * kTpStride)) || (rc = put_generator(E, 0))) return rc;
  // decode: proof (A, B, s), tracker (rG, krG), commitment kG
  std::vector<uint8_t> enc(B * 5 * 48);
  std::vector<uint32_t> dst(5 * B);
  std::vector<Fr> S(B);
  for (size_t b = 0; b < B; b++) {
    uint32_t base = (uint32_t)(1 + b * kTpStride);
    uint8_t* e = enc.data() + b * 5 * 48;
    memcpy(e, trackers + 96 * b, 96);          // rG, krG
    memcpy(e + 96, k_comms + 48 * b, 48);      // kG
    memcpy(e + 144, proofs + 128 * b, 96);     // A, B
    for (uint32_t j = 0; j < 5; j++) dst[5 * b + j] = base + j;
    status_out[b] = cdlh::fr_from_bytes_be_canonical(S[b], proofs + 128 * b + 96) ? CDL_OK : CDL_ERR_DECODE;
  }
  std::vector<uint8_t> st;
  if ((rc = E->decompress(enc.data(), dst, st))) return rc;
  cdlh::MsmStage stage;
  stage.idx.resize(4 * B);
  stage.sc.resize(4 * B);
  stage.tasks.resize(2 * B);
  E->par(B, [&](size_t b) {
    for (uint32_t j = 0; j < 5; j++)
      if (st[5 * b + j]) status_out[b] = CDL_ERR_DECODE;
    uint32_t base = (uint32_t)(1 + b * kTpStride);
    const uint8_t* e = enc.data() + b * 5 * 48;
    Fr ch = status_out[b] == CDL_OK ? tracker_challenge(e + 96, e + 48, e, e + 144, e + 192) : cdlh::FR_ZERO;
    // A' = s*G + c*kG ; B' = s*rG + c*krG
    stage.idx[4 * b] = 0;            stage.sc[4 * b] = S[b];
    stage.idx[4 * b + 1] = base + 2; stage.sc[4 * b + 1] = ch;
    stage.idx[4 * b + 2] = base;     stage.sc[4 * b + 2] = S[b];
    stage.idx[4 * b + 3] = base + 1; stage.sc[4 * b + 3] = ch;
    stage.tasks[2 * b] = cdl::MsmTask{(uint32_t)(4 * b), 2, base + 5, 0};
    stage.tasks[2 * b + 1] = cdl::MsmTask{(uint32_t)(4 * b + 2), 2, base + 6, 0};
  });
  if ((rc = E->run_msm(stage))) return rc;
  for (size_t b = 0; b < B; b++) {
    const uint8_t* e = enc.data() + b * 5 * 48;
    bool good = status_out[b] == CDL_OK && memcmp(stage.out48.data() + 48 * (2 * b), e + 144, 48) == 0 &&
                memcmp(stage.out48.data() + 48 * (2 * b + 1), e + 192, 48) == 0;
    ok[b] = good ? 1 : 0;
  }
  return CDL_OK;
}

}  // extern "C"
