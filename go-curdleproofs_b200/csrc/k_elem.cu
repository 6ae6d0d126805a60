// Elementwise batched point kernels: scalar multiplication / base folding and
// Jacobian -> affine.
#define CDL_FP_MUL_CALL 1  // one shared product body: the hot loops fit the instruction caches (mont.cuh)
#include <cuda_runtime.h>
#include "g1.cuh"
#include "launch.h"

namespace cdl {

// ---------------------------------------------------------------- elementwise
// out[i] = s[i*stride] * P[i] (+ L[i] when L != nullptr), affine result.
// stride 0: one shared scalar (Whisk rescale, IPA / SameMSM folds) — every lane
// runs the identical digit schedule, no divergence.  Scalars arrive in gnark's
// Montgomery fr.Element form and are brought to canonical form here.
// E points per thread share one inversion (see k_elem_ops below).
template <int E>
__global__ void __launch_bounds__(64, 5)
k_scalar_mul(const G1Affine* __restrict__ P, const Fr* __restrict__ s, int stride,
             const G1Affine* __restrict__ L, G1Affine* __restrict__ out, int n) {
  const int T = gridDim.x * blockDim.x;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  G1Jac res[E];
  Fp pre[E];
  Fp run;
  FpM::set_one(run);
#pragma unroll 1
  for (int e = 0; e < E; e++) {
    const int i = e * T + t;
    if (i >= n) break;
    Fr km = s[(size_t)i * stride], k;
    FrM::from_mont(k, km);
    G1Affine p = P[i];
    G1Jac r;
    jac_scalar_mul_glv(r, p, k.v);
    if (L != nullptr) {
      G1Affine l = L[i];
      jac_add_mixed(r, r, l);
    }
    if (E == 1) {
      G1Affine a;
      jac_to_affine(a, r);
      out[i] = a;
      return;
    }
    if (!jac_is_inf(r)) FpM::mul(run, run, r.z);
    res[e] = r;
    pre[e] = run;
  }
  Fp inv;
  fp_inv(inv, run);
#pragma unroll 1
  for (int e = E - 1; e >= 0; e--) {
    const int i = e * T + t;
    if (i >= n) continue;
    G1Jac r = res[e];
    G1Affine a;
    if (jac_is_inf(r)) {
      aff_set_inf(a);
    } else {
      Fp zi;
      if (e > 0) FpM::mul(zi, inv, pre[e - 1]); else zi = inv;
      FpM::mul(inv, inv, r.z);
      jac_to_affine_with_zinv(a, r, zi);
    }
    out[i] = a;
  }
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
// 64-thread CTAs, five per SM (<= 192 registers): 2.5 warps per scheduler keep the integer pipe fed.
//
// Normalisation: a field inversion costs as much as ~210 products (fields.cuh), a tenth of the whole
// scalar multiplication, and a warp pays for it once whether 1 or 32 lanes invert - so sharing one
// inversion across the lanes of a warp saves nothing.  Instead every thread works through E ops one
// after the other (op i = e*T + t: a warp still covers 32 consecutive ops, i.e. one shared scalar and
// a uniform digit schedule per pass), parks the Jacobian results in local memory and inverts the
// product of its E denominators once (Montgomery's trick, bls12381.BatchJacobianToAffineG1 in the
// reference, transcript/transcript.go:26): 3 products + 1/E inversion per point.
template <int E>
__global__ void __launch_bounds__(64, 5)
k_elem_ops(G1Affine* __restrict__ pool, const ElemOp* __restrict__ ops, const Fr* __restrict__ scalars, int n) {
  const int T = gridDim.x * blockDim.x;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  G1Jac res[E];   // local memory (indexed in a rolled loop); 144 B per pending point
  Fp pre[E];      // pre[e] = z_0 * .. * z_e over the finite results
  Fp run;
  FpM::set_one(run);
#pragma unroll 1
  for (int e = 0; e < E; e++) {
    const int i = e * T + t;
    if (i >= n) break;
    ElemOp op = ops[i];
    Fr km = scalars[op.sc], k;
    FrM::from_mont(k, km);
    G1Affine p = pool[op.src];
    G1Jac r;
    jac_scalar_mul_glv(r, p, k.v);
    if (op.add != kNoPoint) {
      G1Affine l = pool[op.add];
      jac_add_mixed(r, r, l);
    }
    if (E == 1) {
      G1Affine a;
      jac_to_affine(a, r);
      pool[op.dst] = a;
      return;
    }
    if (!jac_is_inf(r)) FpM::mul(run, run, r.z);
    res[e] = r;
    pre[e] = run;
  }
  Fp inv;
  fp_inv(inv, run);
#pragma unroll 1
  for (int e = E - 1; e >= 0; e--) {
    const int i = e * T + t;
    if (i >= n) continue;
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
    pool[ops[i].dst] = a;
  }
}

void launch_elem_ops(G1Affine* pool, const ElemOp* ops, const Fr* scalars, int n, cudaStream_t st) {
  const int tpb = 64;
  // enough ops to fill the machine several times over: four per thread share one inversion
  if (n >= 4 * 148 * 5 * tpb) {
    const int threads = (n + 3) / 4;
    k_elem_ops<4><<<(threads + tpb - 1) / tpb, tpb, 0, st>>>(pool, ops, scalars, n);
  } else {
    k_elem_ops<1><<<(n + tpb - 1) / tpb, tpb, 0, st>>>(pool, ops, scalars, n);
  }
}

void launch_scalar_mul(const G1Affine* P, const Fr* s, int stride, const G1Affine* L, G1Affine* out, int n,
                       cudaStream_t st) {
  const int tpb = 64;
  if (n >= 4 * 148 * 5 * tpb) {
    const int threads = (n + 3) / 4;
    k_scalar_mul<4><<<(threads + tpb - 1) / tpb, tpb, 0, st>>>(P, s, stride, L, out, n);
  } else {
    k_scalar_mul<1><<<(n + tpb - 1) / tpb, tpb, 0, st>>>(P, s, stride, L, out, n);
  }
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
