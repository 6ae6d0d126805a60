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
__global__ void __launch_bounds__(64, 5)
k_scalar_mul(const G1Affine* __restrict__ P, const Fr* __restrict__ s, int stride,
             const G1Affine* __restrict__ L, G1Affine* __restrict__ out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr km = s[(size_t)i * stride], k;
  FrM::from_mont(k, km);
  G1Affine p = P[i];
  G1Jac r;
  jac_scalar_mul_glv(r, p, k.v);
  if (L != nullptr) {
    G1Affine l = L[i];
    jac_add_mixed(r, r, l);
  }
  G1Affine a;
  jac_to_affine(a, r);
  out[i] = a;
}

__global__ void __launch_bounds__(128)
k_jac_to_affine(const G1Jac* __restrict__ in, G1Affine* __restrict__ out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1Jac p = in[i];
  G1Affine a;
  jac_to_affine(a, p);
  out[i] = a;
}


// Descriptor-driven form used by the protocol engine: every op reads and writes
// the device-resident point pool, so folded bases never leave HBM between
// rounds (the reference mutates its slices in place the same way,
// innerproductargument.go:157-171).  A launch never has dst aliasing another
// op's src/add, so ops are independent.
// 64-thread CTAs, five per SM (<= 192 registers): 2.5 warps per scheduler keep the integer pipe fed
__global__ void __launch_bounds__(64, 5)
k_elem_ops(G1Affine* __restrict__ pool, const ElemOp* __restrict__ ops, const Fr* __restrict__ scalars, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
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
  G1Affine a;
  jac_to_affine(a, r);
  pool[op.dst] = a;
}

void launch_elem_ops(G1Affine* pool, const ElemOp* ops, const Fr* scalars, int n, cudaStream_t st) {
  const int tpb = 64;
  k_elem_ops<<<(n + tpb - 1) / tpb, tpb, 0, st>>>(pool, ops, scalars, n);
}

void launch_scalar_mul(const G1Affine* P, const Fr* s, int stride, const G1Affine* L, G1Affine* out, int n,
                       cudaStream_t st) {
  const int tpb = 64;
  k_scalar_mul<<<(n + tpb - 1) / tpb, tpb, 0, st>>>(P, s, stride, L, out, n);
}
void launch_jac_to_affine(const G1Jac* in, G1Affine* out, int n, cudaStream_t st) {
  k_jac_to_affine<<<(n + 63) / 64, 64, 0, st>>>(in, out, n);
}

}  // namespace cdl
