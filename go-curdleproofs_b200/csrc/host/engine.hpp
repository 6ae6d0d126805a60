// Batched protocol engine: drives curdleproofs Prove / Verify / ShufflePermuteCommit
// and the Whisk wrappers for B independent instances in lock step.  The host
// keeps what the reference keeps on the host (Merlin transcript, Fiat-Shamir
// challenges, Fr vector arithmetic, RNG, wire format); every group operation is
// a GPU stage over a device-resident pool of affine points:
//   * an MSM stage   — many independent small MSMs in one launch (k_msm_small),
//   * an elementwise stage — pool[dst] = s * pool[src] (+ pool[add]) (k_elem_ops),
//   * (de)compression to/from 48-byte encodings.
// Group elements reach the host only as canonical 48-byte encodings, which is
// all the reference ever observes, so no Fp arithmetic exists on
// the host.
#pragma once
#include <cstdint>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "../context.cuh"
#include "../launch.h"
#include "fr.hpp"
#include "pool.hpp"
#include "transcript.hpp"

struct cdl_rand {
  cdlh::Rand r;
  explicit cdl_rand(uint64_t seed) : r(seed) {}
};

// crs.go:10-18.  Points are kept on the device (pool image) and as compressed
// encodings on the host (for transcript appends of H).
struct cdl_crs {
  cdl_ctx* ctx = nullptr;
  uint32_t ell = 0;
  cdl::G1Affine* d_points = nullptr;  // Gs[ell] Hs[4] H Gt Gu Gsum Hsum INF  (ell + 10 points)
  std::vector<uint8_t> enc;           // same order, 48 B each
  // fixed-base tables of those points (csrc/fixed_base.cuh), built by the first large batch that uses this CRS
  mutable std::mutex fixed_mu;
  mutable cdl::G1Affine* d_fixed = nullptr;
  mutable bool fixed_failed = false;
};

namespace cdlh {

constexpr uint32_t kBlinders = 4;  // common/constants.go:3

struct Layout {
  uint32_t ell, n, m;
  // shared CRS image at the start of the pool
  uint32_t Gs, Hs, H, Gt, Gu, Gsum, Hsum, INF, crs_size;
  // per-instance region (offsets relative to base(b))
  uint32_t Rs, Ss, Ts, Us, M, A, B, scratch, G, Gp, Gm, Tp, Up, PP, inst_size;
  static constexpr uint32_t kScratch = 24;
  explicit Layout(uint32_t ell_);
  uint32_t base(uint32_t b) const { return crs_size + b * inst_size; }
  uint32_t npp() const { return 19 + 10 * m; }  // points in a Whisk shuffle proof (M + 18 + 10m)
};

struct MsmStage {
  std::vector<uint32_t> idx;
  std::vector<Fr> sc;
  std::vector<cdl::MsmTask> tasks;
  std::vector<uint8_t> out48;  // filled by run: tasks.size() * 48
  // host-side accounting only (empty, or one entry per task): terms of the task that exist because one base of
  // the reference's MSM was written as several CRS points (lazy folding); not counted as algorithmic work
  std::vector<uint32_t> extra;
  void clear() { idx.clear(); sc.clear(); tasks.clear(); extra.clear(); }
};

// per-instance builder writing into a pre-sized slice of a stage
struct MsmSlice {
  uint32_t* idx;
  Fr* sc;
  cdl::MsmTask* tasks;
  uint32_t* extra = nullptr;  // MsmStage::extra of this slice's tasks
  uint32_t term_base;  // global index of idx[0]
  uint32_t nterm = 0, ntask = 0;
  void begin(uint32_t out_idx) {
    tasks[ntask].term_off = term_base + nterm;
    tasks[ntask].term_cnt = 0;
    tasks[ntask].out_idx = out_idx;
    tasks[ntask].pad = 0;
  }
  void term(uint32_t point, const Fr& s) {
    idx[nterm] = point;
    sc[nterm] = s;
    nterm++;
    tasks[ntask].term_cnt++;
  }
  // a further term of the base the previous term() call started (see MsmStage::extra)
  void term_more(uint32_t point, const Fr& s) {
    term(point, s);
    if (extra) extra[ntask]++;
  }
  void end() { ntask++; }
};

class Engine {
 public:
  explicit Engine(cdl_ctx* ctx);
  ~Engine();

  // --- API-level operations (all batched over B instances) -----------------
  // common.ShufflePermuteCommit (common/util.go:45-88) on instance points that
  // are already in the pool (Rs, Ss); writes Ts, Us, M into the pool.
  int32_t shuffle_permute_commit(const Layout& L, uint32_t B, const std::vector<std::vector<uint32_t>>& perms,
                                 const std::vector<Fr>& ks, std::vector<cdl_rand*>& rands,
                                 std::vector<std::vector<Fr>>& rs_m);
  // curdleproof.Prove (curdleproof.go:38-197); proofs[b] receives the serialized
  // proof (curdleproof.go:358-387).  status[b] = CDL_OK or the error.
  // inst_enc receives, per instance, the encodings of Rs | Ss | Ts | Us | M.
  // witness_is_ours: M and rs_m were produced by shuffle_permute_commit in the same call (the Whisk
  // wrapper), so the reference's msm(G', d) == D self-check (grandproductargument.go:171-177) cannot fail.
  int32_t prove(const Layout& L, uint32_t B, const cdl_crs* crs, const std::vector<std::vector<uint32_t>>& perms,
                const std::vector<Fr>& ks, const std::vector<std::vector<Fr>>& rs_m, std::vector<cdl_rand*>& rands,
                std::vector<std::vector<uint8_t>>& proofs, std::vector<int32_t>& status,
                std::vector<std::string>& errs, std::vector<uint8_t>& inst_enc, bool witness_is_ours);
  // curdleproof.Verify (curdleproof.go:199-318) for proofs whose points have
  // been decompressed into PP (see parse_and_load_proofs).
  struct ParsedProof {
    bool ok = false;          // decoded without error
    std::string err;
    uint32_t lens[10] = {};   // the ten slice lengths in wire order
    std::vector<uint32_t> pt; // pool index of each decoded point, wire order, M first
    std::vector<const uint8_t*> enc;  // 48-byte encoding of each of those points
    std::vector<Fr> sc;       // the 7 scalars in wire order: Rp c0 d0 Z_k Z_t Z_u x
    const uint8_t* inst_enc = nullptr;  // Rs | Ss | Ts | Us encodings, 4*ell*48 bytes
  };
  int32_t verify(const Layout& L, uint32_t B, const cdl_crs* crs, std::vector<ParsedProof>& pp,
                 std::vector<cdl_rand*>& rands, std::vector<int32_t>& verdict, std::vector<int32_t>& status,
                 std::vector<std::string>& errs);

  // --- pool plumbing ----------------------------------------------------------
  int32_t ensure_pool(size_t npoints);
  int32_t load_crs(const Layout& L, const cdl_crs* crs);
  // fixed-base tables for this call's CRS image (pool indices below crs_size); B = instances of the call
  void select_fixed(const Layout& L, const cdl_crs* crs, size_t B);
  void clear_fixed() { fixed_ = cdl::FixedTable(); }  // calls whose pool does not start with a CRS image
  int32_t upload_points(uint32_t dst, const void* host_affine, size_t count);
  int32_t upload_jac(uint32_t dst, const void* host_jac, size_t count);  // normalises to affine on the device
  int32_t download_points(uint32_t src, void* host_affine, size_t count);
  int32_t copy_points(uint32_t dst, uint32_t src, size_t count);
  int32_t set_infinity(uint32_t dst, size_t count);
  int32_t copy_ranges(const std::vector<cdl::CopyRange>& ranges);  // many pool-to-pool copies in one launch
  int32_t compress(const std::vector<uint32_t>& src, std::vector<uint8_t>& out48);
  int32_t decompress(const uint8_t* enc48, const std::vector<uint32_t>& dst, std::vector<uint8_t>& status);
  // `before` (optional) runs on the stream after the stage's indices / scalars reached the device and before
  // the MSM kernels: the verifier's scalar pipeline fills part of the scalar array there (d_scalars)
  int32_t run_msm(MsmStage& st, const std::function<int32_t(cdl::Fr* d_scalars)>& before = nullptr);
  int32_t run_elem(const std::vector<cdl::ElemOp>& ops, const std::vector<Fr>& sc);

  cdl_ctx* ctx() { return ctx_; }
  ThreadPool& threads() { return pool_; }
  // instrumentation for bench.py: per kernel class (0 msm, 1 elementwise, 2 decompress,
  // 3 compress / normalise): launches, CUDA-event time on the launching stream,
  // algorithmic modmul (SURVEY.md §8d conventions) and algorithmic bytes
  uint64_t launches = 0;
  // [start, end) of every timed kernel chain on the root context's timeline (ms since base_ev);
  // the union over all lanes is the device busy time
  std::vector<std::pair<float, float>> intervals;
  struct Stats {
    uint64_t n[4] = {};
    double ms[4] = {};
    double modmul[4] = {};
    double bytes[4] = {};
  } stats;
  static double msm_algorithmic_modmul(uint64_t terms);
  // CDL_PROFILE=1: wall-clock split of the protocol calls, printed when the engine is destroyed
  struct HostProf {
    double par = 0;    // inside parallel_for sections (per-proof host work)
    double gpu = 0;    // inside GPU stage calls (staging copies + launch + wait)
    double copy = 0;   // of which: filling the pinned staging buffers
    double total = 0;  // prove + verify + shuffle_permute_commit calls
  } prof;
  // Per-proof host work of a stage.  Eight proofs at a time share a pool thread as cooperating
  // fibers so that their Keccak permutations run as one eight-way call (host/fiber.hpp).
  template <class F>
  void par(size_t n, F&& f) {
    double t0 = now();
    std::function<void(size_t)> fn(std::forward<F>(f));
    if (n >= 4 && fibers_available()) {
      const size_t groups = (n + 7) / 8;
      pool_.parallel_for(groups, std::function<void(size_t)>([&](size_t g) {
        run_fiber_group(fn, 8 * g, n - 8 * g < 8 ? n - 8 * g : 8);
      }));
    } else {
      pool_.parallel_for(n, fn);
    }
    prof.par += now() - t0;
  }
  static double now();
  static unsigned host_threads(bool lane);

 private:
  cdl_ctx* ctx_;
  ThreadPool pool_;
  cdl::G1Affine* d_pool_ = nullptr;
  cdl::FixedTable fixed_;  // tables of the current call's CRS, or empty
  size_t pool_cap_ = 0;
  void* d_win_ = nullptr;  // window sums of the two-kernel MSM path
  size_t win_cap_ = 0;
  // pinned host staging + device scratch, grow-only
  struct Staging {
    void* h = nullptr;
    void* d = nullptr;
    size_t cap = 0;
  };
  Staging s_idx_, s_sc_, s_task_, s_out_, s_ops_, s_enc_, s_st_, s_jac_, s_sub_, s_t2_, s_cr_, s_fsub_, s_vs_, s_as_;
  int32_t reserve(Staging& s, size_t bytes);
  void tick();                                   // event before a kernel
  void tock(int cls, double modmul, double bytes);  // event after; call finish_timing() after the sync
  int pending_cls_ = -1;
  void finish_timing();
};

// Whisk proof wire walker (whisk/types.go:39-72, curdleproof.go:320-387):
// splits a serialized proof into point encodings, slice lengths and scalars.
struct WireProof {
  std::vector<const uint8_t*> points;  // 48-byte compressed encodings, wire order
  uint32_t lens[10];
  const uint8_t* scalars[7];
};
// returns empty string on success, else the decode error
std::string parse_wire_proof(const uint8_t* buf, size_t len, bool with_m, WireProof& out, size_t* used);

}  // namespace cdlh
