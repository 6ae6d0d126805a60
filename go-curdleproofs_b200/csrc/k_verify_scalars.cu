// Verifier scalar pipeline on the device.
//
// curdleproof.Verify defers its eight checks into one MSM whose scalars the reference builds on the host:
// the unfolded IPA vectors s, s' (innerproductargument.go:223-234), the unfolded same-multiscalar vector
// (samemultiscalarargument.go:267-277), the powers of beta^-1 (grandproductargument.go:234-242) and the
// accumulator's merge by base with one random weight per check (msmaccumulator.go:23-47).  For the 5*ell + 7
// bases every proof shares structurally (Gs, Hs, H, Gt, Gu, Ts, Us, Rs, Ss) those scalars are pure functions
// of ~60 field elements per proof (challenges, weights, proof scalars) and of the ell vector challenges, so
// the host uploads that block and this kernel writes the scalars straight into the MSM stage's scalar array:
//   s_Gs[i] = a1*beta_sp + a2*c0*s[i] + a3*d0*s'[i]*u[i] + a4*x*t[i]      s_Ts[i] = a5*x*t[i]
//   s_Us[i] = a6*x*t[i]        s_Rs[i] = a7*as[i]        s_Ss[i] = a8*as[i]
// with s[i] = prod_{bit j of i} gamma[m-1-j], s'[i] the same over gamma^-1, t[i] over the same-multiscalar
// challenges and u[i] = beta^-(min(i, ell) + 1); the four Hs, and H / Gt / Gu, take the entries ell .. ell+3.
// One thread per (proof, i < n): ~ 3m + 2 log2(ell) + 12 products of the 255-bit field.
#include <cuda_runtime.h>
#include "fields.cuh"
#include "launch.h"

namespace cdl {

__global__ void __launch_bounds__(128)
k_verify_scalars(const VsParams* __restrict__ params, const Fr* __restrict__ as, Fr* __restrict__ sc, uint32_t ell,
                 uint32_t n, uint32_t m) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const VsParams& P = params[blockIdx.y];
  Fr sv, svp, t;
  FrM::set_one(sv);
  FrM::set_one(svp);
  FrM::set_one(t);
#pragma unroll 1
  for (uint32_t j = 0; j < m; j++)
    if (i & (1u << j)) {
      FrM::mul(sv, sv, P.gamma[m - 1 - j]);
      FrM::mul(svp, svp, P.gamma_inv[m - 1 - j]);
      FrM::mul(t, t, P.ch[m - 1 - j]);
    }
  // u = beta^-(e), e = min(i, ell) + 1, MSB-first square and multiply
  const uint32_t e = (i < ell ? i : ell) + 1;
  Fr u = P.beta_inv;
#pragma unroll 1
  for (int b = 30 - __clz((int)e); b >= 0; b--) {
    FrM::sqr(u, u);
    if ((e >> b) & 1u) FrM::mul(u, u, P.beta_inv);
  }
  Fr ipa, tmp, xt4;
  FrM::mul(ipa, P.a2c0, sv);
  FrM::mul(tmp, svp, u);
  FrM::mul(tmp, tmp, P.a3d0);
  FrM::add(ipa, ipa, tmp);  // a2*c0*s + a3*d0*s'*u
  FrM::mul(xt4, P.a4xf, t);
  Fr* out = sc + P.sc_base;
  if (i < ell) {
    Fr g, a;
    FrM::add(g, ipa, P.a1b);
    FrM::add(g, g, xt4);
    out[i] = g;
    FrM::mul(tmp, P.a5xf, t);
    out[ell + 7 + i] = tmp;
    FrM::mul(tmp, P.a6xf, t);
    out[2 * ell + 7 + i] = tmp;
    a = as[P.as_base + i];
    FrM::mul(tmp, P.a7, a);
    out[3 * ell + 7 + i] = tmp;
    FrM::mul(tmp, P.a8, a);
    out[4 * ell + 7 + i] = tmp;
  } else {
    const uint32_t j = i - ell;  // 0..3: Hs[j]; the same-multiscalar vector is Gs || Hs[0..2) || Gt || Gu
    Fr h = ipa;
    if (j < 2) FrM::add(h, h, xt4);
    out[ell + j] = h;
    if (j == 2) out[ell + 5] = xt4;  // Gt
    if (j == 3) {
      out[ell + 6] = xt4;            // Gu
      // H: host part + a5*x*t[ell+2] + a6*x*t[ell+3]; t[ell+2] is recomputed here (the product over the bits of i - 1)
      Fr t2, sH;
      FrM::set_one(t2);
#pragma unroll 1
      for (uint32_t jj = 0; jj < m; jj++)
        if ((i - 1) & (1u << jj)) FrM::mul(t2, t2, P.ch[m - 1 - jj]);
      FrM::mul(sH, P.a5xf, t2);
      FrM::mul(tmp, P.a6xf, t);
      FrM::add(sH, sH, tmp);
      FrM::add(sH, sH, P.sH_host);
      out[ell + 4] = sH;
    }
  }
}

void launch_verify_scalars(const VsParams* params, const Fr* as, Fr* sc, uint32_t nproofs, uint32_t ell, uint32_t n,
                           uint32_t m, cudaStream_t st) {
  if (!nproofs) return;
  dim3 grid((n + 127) / 128, nproofs);
  k_verify_scalars<<<grid, 128, 0, st>>>(params, as, sc, ell, n, m);
}

}  // namespace cdl
