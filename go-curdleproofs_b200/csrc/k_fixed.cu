// Builds the fixed-base tables of a CRS (fixed_base.cuh) on the device.
#define CDL_FP_MUL_CALL 1
#include "fixed_base.cuh"
#include "launch.h"

namespace cdl {

constexpr int kFbSeg = 8;                  // a (base, window) row is built by 8 threads, 256 entries each
constexpr int kFbSegLen = kFbM / kFbSeg;
constexpr int kFbNorm = 8;                 // entries normalised with one inversion

// thread (b, w, s): Q = 2^(12 w) P_b, then the entries (256 s + 1) Q .. (256 s + 256) Q by repeated addition
// of Q; every 8 consecutive entries share one inversion (Montgomery's trick)
__global__ void __launch_bounds__(64)
k_fixed_build(const G1Affine* __restrict__ bases, uint32_t nbase, G1Affine* __restrict__ tab) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nbase * (uint32_t)(kFbW * kFbSeg)) return;
  const uint32_t b = t / (kFbW * kFbSeg), r = t - b * (kFbW * kFbSeg);
  const int w = (int)(r / kFbSeg), s = (int)(r - (uint32_t)w * kFbSeg);
  G1Affine* row = tab + ((size_t)b * kFbW + w) * kFbM + (size_t)s * kFbSegLen;
  const G1Affine p = bases[b];
  G1Affine q;
  {
    G1Jac j;
    jac_from_affine(j, p);
#pragma unroll 1
    for (int i = 0; i < w * kFbC; i++) jac_dbl(j, j);
    jac_to_affine(q, j);  // infinity stays infinity: the whole row is infinity then
  }
  // acc = (256 s) Q
  G1Xyzz acc;
  xyzz_set_inf(acc);
  {
    const uint32_t e = (uint32_t)s * kFbSegLen;
    if (e) {
      const int top = 31 - __clz(e);
#pragma unroll 1
      for (int i = top; i >= 0; i--) {
        xyzz_dbl(acc, acc);
        if ((e >> i) & 1u) xyzz_add_mixed(acc, acc, q);
      }
    }
  }
  G1Xyzz pend[kFbNorm];
  Fp pre[kFbNorm];
#pragma unroll 1
  for (int base = 0; base < kFbSegLen; base += kFbNorm) {
    Fp run;
    FpM::set_one(run);
#pragma unroll 1
    for (int e = 0; e < kFbNorm; e++) {
      xyzz_add_mixed(acc, acc, q);
      if (!xyzz_is_inf(acc)) FpM::mul(run, run, acc.zzz);
      pend[e] = acc;
      pre[e] = run;
    }
    Fp inv;
    fp_inv(inv, run);
#pragma unroll 1
    for (int e = kFbNorm - 1; e >= 0; e--) {
      const G1Xyzz v = pend[e];
      G1Affine a;
      if (xyzz_is_inf(v)) {
        aff_set_inf(a);
      } else {
        Fp zi;  // 1 / zzz
        if (e > 0) FpM::mul(zi, inv, pre[e - 1]); else zi = inv;
        FpM::mul(inv, inv, v.zzz);
        // x = X / ZZ = X * ZZZ^-2 * ... : with ZZ^3 = ZZZ^2, 1/ZZ = (ZZ * (1/ZZZ))^2
        Fp u, izz;
        FpM::mul(u, v.zz, zi);  // ZZ / ZZZ = 1 / Z
        FpM::sqr(izz, u);       // 1 / ZZ
        FpM::mul(a.x, v.x, izz);
        FpM::mul(a.y, v.y, zi);
      }
      row[base + e] = a;
    }
  }
}

size_t fixed_table_bytes(uint32_t nbase) { return (size_t)nbase * kFbW * kFbM * sizeof(G1Affine); }

void launch_fixed_build(const G1Affine* bases, uint32_t nbase, G1Affine* tab, cudaStream_t st) {
  const uint32_t threads = nbase * (uint32_t)(kFbW * kFbSeg);
  if (threads) k_fixed_build<<<(threads + 63) / 64, 64, 0, st>>>(bases, nbase, tab);
}

}  // namespace cdl
