// Elementwise batched point kernels: scalar multiplication / base folding and
// Jacobian -> affine.
#define CDL_FP_MUL_CALL 1  // one shared product body: the hot loops fit the instruction caches (mont.cuh)
#include <algorithm>
#include <cuda_runtime.h>
#include "quad.cuh"
#include "launch.h"

namespace cdl {

// ---------------------------------------------------------------- elementwise
// out[i] = s[i*stride] * P[i] (+ L[i] when L != nullptr), affine result.
// stride 0: one shared scalar (Whisk rescale, IPA / SameMSM folds) — every lane
// runs the identical digit schedule, no divergence.  Scalars arrive in gnark's
// Montgomery fr.Element form and are brought to canonical form here.
// Shared body of the two scalar-multiplication kernels (k_elem_ops below explains the schedule):
// a persistent grid, thread t takes items t, t + T, t + 2T, ..; every kEB consecutive items of a
// thread are normalised with ONE inversion.  load(i, p, k, add, has_add) fetches item i.
constexpr int kEB = 8;
#ifndef CDL_ELEM_MINB
#define CDL_ELEM_MINB 4  // resident 64-thread CTAs per SM the register budget is sized for (238 registers, no spills;
                         // 5 or 6 per SM cap the kernel at 168 registers and spill 400 B in the hot loop: 7 % slower)
#endif

template <class Load, class Store>
__device__ __forceinline__ void scalar_mul_batched(int n, const FixedTable& ft, Load load, Store store) {
  const int T = gridDim.x * blockDim.x;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  G1Jac res[kEB];  // local memory (indexed in rolled loops): 144 B per pending point
  Fp pre[kEB];     // pre[e] = z_0 * .. * z_e over the finite results of the chunk
#pragma unroll 1
  for (int base = t; base < n; base += kEB * T) {
    Fp run;
    FpM::set_one(run);
    int cnt = 0;
#pragma unroll 1
    for (int e = 0; e < kEB; e++) {
      const int i = base + e * T;
      if (i >= n) break;
      G1Affine p, l;
      Fr km, k;
      bool has_add;
      int fb = -1;  // index of a tabulated base (fixed_base.cuh), or -1
      load(i, p, km, l, has_add, fb);
      FrM::from_mont(k, km);
      G1Jac r;
      if (fb >= 0) {  // uniform across a warp in practice: a stage's ops are all on CRS points or none is
        G1Xyzz x;
        xyzz_set_inf(x);
        fixed_base_accumulate<false>(x, ft, (uint32_t)fb, k.v, false);
        xyzz_to_jac(r, x);
      } else {
        jac_scalar_mul_glv(r, p, k.v);
      }
      if (has_add) jac_add_mixed(r, r, l);
      if (!jac_is_inf(r)) FpM::mul(run, run, r.z);
      res[e] = r;
      pre[e] = run;
      cnt++;
    }
    Fp inv;
    fp_inv(inv, run);
#pragma unroll 1
    for (int e = cnt - 1; e >= 0; e--) {
      G1Jac r = res[e];
      G1Affine a;
      if (jac_is_inf(r)) {
        aff_set_inf(a);
      } else {
        Fp zi;
        if (e > 0) FpM::mul(zi, inv, pre[e - 1]); else zi = inv;   // 1 / z_e
        FpM::mul(inv, inv, r.z);                                   // 1 / (z_0 .. z_{e-1})
        jac_to_affine_with_zinv(a, r, zi);
      }
      store(base + e * T, a);
    }
  }
}

__global__ void __launch_bounds__(64, CDL_ELEM_MINB)
k_scalar_mul(const G1Affine* __restrict__ P, const Fr* __restrict__ s, int stride,
             const G1Affine* __restrict__ L, G1Affine* __restrict__ out, int n) {
  scalar_mul_batched(
      n, FixedTable(),
      [&](int i, G1Affine& p, Fr& km, G1Affine& l, bool& has_add, int& fb) {
        (void)fb;
        km = s[(size_t)i * stride];
        p = P[i];
        has_add = L != nullptr;
        if (has_add) l = L[i];
      },
      [&](int i, const G1Affine& a) { out[i] = a; });
}

// bls12381.BatchJacobianToAffineG1 (transcript/transcript.go:26): every thread walks E points
// (strided, so a warp reads consecutive points), multiplies their Z coordinates up, inverts once and
// walks back (Montgomery's trick: 3 products + 1/E inversion per point); Z = 0 maps to (0, 0).
template <int E>
__global__ void __launch_bounds__(64)
k_jac_to_affine(const G1Jac* __restrict__ in, G1Affine* __restrict__ out, int n) {
  const int T = gridDim.x * blockDim.x;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  Fp pre[E];
  Fp run;
  FpM::set_one(run);
#pragma unroll 1
  for (int e = 0; e < E; e++) {
    const int i = e * T + t;
    if (i >= n) break;
    Fp z = in[i].z;
    if (!FpM::is_zero(z)) FpM::mul(run, run, z);
    pre[e] = run;
  }
  Fp inv;
  fp_inv(inv, run);
#pragma unroll 1
  for (int e = E - 1; e >= 0; e--) {
    const int i = e * T + t;
    if (i >= n) continue;
    G1Jac p = in[i];
    G1Affine a;
    if (jac_is_inf(p)) {
      aff_set_inf(a);
    } else {
      Fp zi;
      if (e > 0) FpM::mul(zi, inv, pre[e - 1]); else zi = inv;
      FpM::mul(inv, inv, p.z);
      jac_to_affine_with_zinv(a, p, zi);
    }
    out[i] = a;
  }
}


// Descriptor-driven form used by the protocol engine: every op reads and writes
// the device-resident point pool, so folded bases never leave HBM between
// rounds (the reference mutates its slices in place the same way,
// innerproductargument.go:157-171).  A launch never has dst aliasing another
// op's src/add, so ops are independent.
// 64-thread CTAs, four per SM (two warps per scheduler saturate the carry-chain pipe, profiles/r1_ilp_probe.txt).
//
// Normalisation: a field inversion costs as much as ~210 products (fields.cuh), 12 % of the whole
// scalar multiplication, and a warp pays for it once whether 1 or 32 lanes invert - so sharing one
// inversion across the lanes of a warp saves nothing.  Instead the grid is persistent (one wave of
// resident CTAs), every thread works through its items t, t + T, .. (a warp still covers 32
// consecutive ops per pass, i.e. one shared scalar and a uniform digit schedule), parks up to kEB
// Jacobian results in local memory and inverts the product of their denominators once (Montgomery's
// trick, bls12381.BatchJacobianToAffineG1 in the reference, transcript/transcript.go:26): 3 products
// + 1/kEB inversion per point.  The persistent grid also removes the tail of the last wave: every
// thread of a launch does floor or ceil of n / T equal-cost items.
__global__ void __launch_bounds__(64, CDL_ELEM_MINB)
k_elem_ops(G1Affine* __restrict__ pool, const ElemOp* __restrict__ ops, const Fr* __restrict__ scalars, int n,
           const FixedTable ft) {
  scalar_mul_batched(
      n, ft,
      [&](int i, G1Affine& p, Fr& km, G1Affine& l, bool& has_add, int& fb) {
        const ElemOp op = ops[i];
        km = scalars[op.sc];
        if (op.src < ft.nbase) fb = (int)op.src;
        p = pool[op.src];
        has_add = op.add != kNoPoint;
        if (has_add) l = pool[op.add];
      },
      [&](int i, const G1Affine& a) { pool[ops[i].dst] = a; });
}

// Latency form for small launches (one proof at a time: 126 .. 1 016 ops per stage): FOUR lanes per
// op (quad.cuh).  With one thread per point a stage is one 1 950-product dependent chain (1.75 ms); the
// quad evaluates the independent products of every doubling / addition in parallel (3 and 4 levels
// instead of 9 and 14 products) and finishes in about a third of that.  It costs ~1.8x the issue
// slots, so launches that fill the machine keep the thread-per-point kernel above.
constexpr int kQuadOpsPerCta = 8;  // 32 threads, 24 KB of window tables

template <class Load, class Store>
__device__ __forceinline__ void scalar_mul_quad(int n, Load load, Store store) {
  __shared__ G1Xyzz tab[kQuadOpsPerCta][16];
  const Quad q;
  const int slot = threadIdx.x >> 2;
  const int i = blockIdx.x * kQuadOpsPerCta + slot;
  if (i >= n) return;  // the whole quad
  G1Affine p, l;
  Fr km, k;
  bool has_add;
  load(i, p, km, l, has_add);
  FrM::from_mont(k, km);
  G1Xyzz r;
  qxyzz_scalar_mul_glv(q, r, p, k.v, tab[slot]);
  if (has_add) qxyzz_add_mixed(q, r, r, l);
  G1Affine a;
  qxyzz_to_affine(q, a, r);
  if (q.lane == 0) store(i, a);
}

__global__ void __launch_bounds__(4 * kQuadOpsPerCta)
k_elem_ops_quad(G1Affine* __restrict__ pool, const ElemOp* __restrict__ ops, const Fr* __restrict__ scalars, int n) {
  scalar_mul_quad(
      n,
      [&](int i, G1Affine& p, Fr& km, G1Affine& l, bool& has_add) {
        const ElemOp op = ops[i];
        km = scalars[op.sc];
        p = pool[op.src];
        has_add = op.add != kNoPoint;
        if (has_add) l = pool[op.add];
      },
      [&](int i, const G1Affine& a) { pool[ops[i].dst] = a; });
}

__global__ void __launch_bounds__(4 * kQuadOpsPerCta)
k_scalar_mul_quad(const G1Affine* __restrict__ P, const Fr* __restrict__ s, int stride,
                  const G1Affine* __restrict__ L, G1Affine* __restrict__ out, int n) {
  scalar_mul_quad(
      n,
      [&](int i, G1Affine& p, Fr& km, G1Affine& l, bool& has_add) {
        km = s[(size_t)i * stride];
        p = P[i];
        has_add = L != nullptr;
        if (has_add) l = L[i];
      },
      [&](int i, const G1Affine& a) { out[i] = a; });
}

constexpr int kQuadMaxOps = 8192;  // above this the thread-per-point kernel has enough warps to hide its chains

// resident CTAs of a kernel on the current device (persistent-grid size)
template <class K>
static int resident_ctas(K kernel, int tpb) {
  int dev = 0, sms = 0, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, tpb, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
  return sms * per_sm;
}

// Grid of a launch over n equal-cost items: with at most `resident` CTAs in flight a thread takes
// k = ceil(n / (resident * tpb)) items; the grid is then shrunk to ceil(n / k) threads so that every
// thread does k items (k - 1 for a few) instead of some doing k and the rest idling through the last pass.
static int balanced_blocks(int n, int tpb, int resident) {
  const long long tmax = (long long)resident * tpb;
  const int k = (int)((n + tmax - 1) / tmax);
  const int threads = (n + k - 1) / k;
  return (threads + tpb - 1) / tpb;
}

void launch_elem_ops(G1Affine* pool, const ElemOp* ops, const Fr* scalars, int n, cudaStream_t st, FixedTable ft) {
  if (n <= 0) return;
  if (n <= kQuadMaxOps) {
    k_elem_ops_quad<<<(n + kQuadOpsPerCta - 1) / kQuadOpsPerCta, 4 * kQuadOpsPerCta, 0, st>>>(pool, ops, scalars, n);
    return;
  }
  const int tpb = 64;
  static thread_local int resident = 0;  // one device per context thread; re-queried per thread
  if (!resident) resident = resident_ctas(k_elem_ops, tpb);
  if (!ft.tab) ft.nbase = 0;
  k_elem_ops<<<balanced_blocks(n, tpb, resident), tpb, 0, st>>>(pool, ops, scalars, n, ft);
}

void launch_scalar_mul(const G1Affine* P, const Fr* s, int stride, const G1Affine* L, G1Affine* out, int n,
                       cudaStream_t st) {
  if (n <= 0) return;
  if (n <= kQuadMaxOps) {
    k_scalar_mul_quad<<<(n + kQuadOpsPerCta - 1) / kQuadOpsPerCta, 4 * kQuadOpsPerCta, 0, st>>>(P, s, stride, L, out, n);
    return;
  }
  const int tpb = 64;
  static thread_local int resident = 0;
  if (!resident) resident = resident_ctas(k_scalar_mul, tpb);
  k_scalar_mul<<<balanced_blocks(n, tpb, resident), tpb, 0, st>>>(P, s, stride, L, out, n);
}
// pool[dst .. dst+count) = pool[src .. src+count) (src == kNoPoint: the point at infinity) for many ranges
// in one launch: the working vectors of a batch (G, G', T', U' of every instance) are set up by a
// handful of launches instead of thousands of tiny device-to-device copies.
__global__ void __launch_bounds__(128)
k_copy_ranges(G1Affine* __restrict__ pool, const CopyRange* __restrict__ ranges, int n) {
  const CopyRange r = ranges[blockIdx.x];
  uint4* d = reinterpret_cast<uint4*>(pool + r.dst);
  const int granules = (int)r.count * (int)(sizeof(G1Affine) / 16);
  if (r.src == kNoPoint) {
    for (int i = threadIdx.x; i < granules; i += blockDim.x) d[i] = make_uint4(0, 0, 0, 0);
  } else {
    const uint4* s = reinterpret_cast<const uint4*>(pool + r.src);
    for (int i = threadIdx.x; i < granules; i += blockDim.x) d[i] = s[i];
  }
  (void)n;
}
void launch_copy_ranges(G1Affine* pool, const CopyRange* ranges, int n, cudaStream_t st) {
  if (n > 0) k_copy_ranges<<<n, 128, 0, st>>>(pool, ranges, n);
}

void launch_jac_to_affine(const G1Jac* in, G1Affine* out, int n, cudaStream_t st) {
  if (n <= 0) return;
  if (n >= 8 * 148 * 64) {
    const int threads = (n + 7) / 8;
    k_jac_to_affine<8><<<(threads + 63) / 64, 64, 0, st>>>(in, out, n);
  } else if (n >= 2 * 148 * 64) {
    const int threads = (n + 1) / 2;
    k_jac_to_affine<2><<<(threads + 63) / 64, 64, 0, st>>>(in, out, n);
  } else {
    k_jac_to_affine<1><<<(n + 63) / 64, 64, 0, st>>>(in, out, n);
  }
}

}  // namespace cdl
