// Host build of the product's field / curve headers (the carry-chain primitives
// are emulated bit-exactly on the CPU, see csrc/carry.cuh).  TEST HARNESS ONLY:
// lets tests/ drive fp/fr/g1 code paths against the oracle without a GPU.  It is
// not part of the shipped library and is never used as a fallback.
#include <cstring>
#include "../../go-curdleproofs_b200/csrc/g1.cuh"
#include "../../go-curdleproofs_b200/csrc/batch_affine.cuh"
#include "../../go-curdleproofs_b200/csrc/fixed_base.cuh"
#include <vector>

using namespace cdl;

extern "C" {

void hc_fp_mul(const Fp* a, const Fp* b, Fp* r, int n) { for (int i = 0; i < n; i++) FpM::mul(r[i], a[i], b[i]); }
void hc_fp_sqr(const Fp* a, Fp* r, int n) { for (int i = 0; i < n; i++) FpM::sqr(r[i], a[i]); }
void hc_fr_sqr(const Fr* a, Fr* r, int n) { for (int i = 0; i < n; i++) FrM::sqr(r[i], a[i]); }
void hc_fp_add(const Fp* a, const Fp* b, Fp* r, int n) { for (int i = 0; i < n; i++) FpM::add(r[i], a[i], b[i]); }
void hc_fp_sub(const Fp* a, const Fp* b, Fp* r, int n) { for (int i = 0; i < n; i++) FpM::sub(r[i], a[i], b[i]); }
void hc_fp_neg(const Fp* a, Fp* r, int n) { for (int i = 0; i < n; i++) FpM::neg(r[i], a[i]); }
void hc_fp_inv(const Fp* a, Fp* r, int n) { for (int i = 0; i < n; i++) fp_inv(r[i], a[i]); }
void hc_fp_inv_eea(const Fp* a, Fp* r, int n) { for (int i = 0; i < n; i++) fp_inv_eea(r[i], a[i]); }
void hc_fp_sqrt(const Fp* a, Fp* r, int* ok, int n) { for (int i = 0; i < n; i++) ok[i] = fp_sqrt(r[i], a[i]); }
void hc_fp_from_mont(const Fp* a, Fp* r, int n) { for (int i = 0; i < n; i++) FpM::from_mont(r[i], a[i]); }
void hc_fp_to_mont(const Fp* a, Fp* r, int n) { for (int i = 0; i < n; i++) FpM::to_mont(r[i], a[i]); }
void hc_fp_lex(const Fp* a, int* r, int n) { for (int i = 0; i < n; i++) r[i] = fp_lex_largest(a[i]); }
void hc_fr_mul(const Fr* a, const Fr* b, Fr* r, int n) { for (int i = 0; i < n; i++) FrM::mul(r[i], a[i], b[i]); }
void hc_fr_from_mont(const Fr* a, Fr* r, int n) { for (int i = 0; i < n; i++) FrM::from_mont(r[i], a[i]); }

// op: 0 jac mixed add, 1 jac full add, 2 xyzz mixed add, 3 xyzz full add, 4 jac dbl, 5 xyzz dbl
void hc_g1_binop(int op, const G1Affine* a, const G1Affine* b, G1Affine* r, int n) {
  for (int i = 0; i < n; i++) {
    G1Jac ja, jb, jr;
    G1Xyzz xa, xb, xr;
    jac_from_affine(ja, a[i]);
    jac_from_affine(jb, b[i]);
    // de-normalise so that Z != 1 paths are exercised: double then use as-is where possible
    switch (op) {
      case 0: jac_dbl(ja, ja); jac_add_mixed(jr, ja, b[i]); break;                 // 2a + b
      case 1: jac_dbl(ja, ja); jac_dbl(jb, jb); jac_add(jr, ja, jb); break;       // 2a + 2b
      case 2: xyzz_from_affine(xa, a[i]); xyzz_dbl(xa, xa); xyzz_add_mixed(xr, xa, b[i]); xyzz_to_jac(jr, xr); break;
      case 3: xyzz_from_affine(xa, a[i]); xyzz_from_affine(xb, b[i]); xyzz_dbl(xa, xa); xyzz_dbl(xb, xb);
              xyzz_add(xr, xa, xb); xyzz_to_jac(jr, xr); break;
      case 4: jac_dbl(jr, ja); break;                                               // 2a
      case 5: xyzz_from_affine(xa, a[i]); xyzz_dbl(xr, xa); xyzz_to_jac(jr, xr); break;
      case 6: jac_add_mixed(jr, ja, b[i]); break;                                   // a + b (exceptional cases)
      case 7: jac_add(jr, ja, jb); break;
      case 8: xyzz_from_affine(xa, a[i]); xyzz_add_mixed(xr, xa, b[i]); xyzz_to_jac(jr, xr); break;
      case 9: xyzz_from_affine(xa, a[i]); xyzz_from_affine(xb, b[i]); xyzz_add(xr, xa, xb); xyzz_to_jac(jr, xr); break;
      default: jac_set_inf(jr);
    }
    jac_to_affine(r[i], jr);
  }
}

// scalars: canonical little-endian 8 x u32
void hc_g1_scalar_mul(const G1Affine* p, const uint32_t* k, G1Affine* r, int n) {
  for (int i = 0; i < n; i++) {
    G1Jac j;
    jac_scalar_mul(j, p[i], k + 8 * i);
    jac_to_affine(r[i], j);
  }
}

void hc_g1_scalar_mul_glv(const G1Affine* p, const uint32_t* k, G1Affine* r, int n) {
  for (int i = 0; i < n; i++) {
    G1Jac j;
    jac_scalar_mul_glv(j, p[i], k + 8 * i);
    jac_to_affine(r[i], j);
  }
}

// k (8 words) -> |k1| (4 words), k2 (4 words), neg1, neg2
void hc_glv_decompose(const uint32_t* k, uint32_t* k1, uint32_t* k2, int* neg1, int* neg2, int n) {
  for (int i = 0; i < n; i++) {
    Glv g;
    glv_decompose(g, k + 8 * i);
    for (int j = 0; j < 4; j++) { k1[4 * i + j] = g.k1[j]; k2[4 * i + j] = g.k2[j]; }
    neg1[i] = g.neg1;
    neg2[i] = g.neg2;
  }
}

void hc_g1_in_subgroup(const G1Affine* p, int* out, int n) {
  for (int i = 0; i < n; i++) out[i] = g1_in_subgroup_endo(p[i]);
}

int hc_on_curve(const G1Affine* p) { return aff_on_curve(*p); }

// r[i] = (neg[i] ? -k[i] : k[i]) * p through a fixed-base table of p built here the way k_fixed_build
// (csrc/k_fixed.cu) builds it — row w holds d * 2^(12 w) * p for d = 1 .. 2048 — and read by the product's
// own fixed_base_accumulate (csrc/fixed_base.cuh).  k: canonical little-endian 8 x u32.
void hc_fixed_base(const G1Affine* p, const uint32_t* k, const int* neg, G1Affine* r, int n) {
  std::vector<G1Affine> tab((size_t)kFbW * kFbM);
  G1Jac q;
  jac_from_affine(q, *p);
  for (int w = 0; w < kFbW; w++) {
    G1Affine qa;
    jac_to_affine(qa, q);
    G1Xyzz acc;
    xyzz_set_inf(acc);
    for (int d = 0; d < kFbM; d++) {
      xyzz_add_mixed(acc, acc, qa);
      G1Jac j;
      xyzz_to_jac(j, acc);
      jac_to_affine(tab[(size_t)w * kFbM + d], j);
    }
    for (int i = 0; i < kFbC; i++) jac_dbl(q, q);
  }
  FixedTable ft;
  ft.tab = tab.data();
  ft.nbase = 1;
  for (int i = 0; i < n; i++) {
    G1Xyzz acc;
    xyzz_set_inf(acc);
    fixed_base_accumulate(acc, ft, 0, k + 8 * i, neg[i] != 0);
    G1Jac j;
    xyzz_to_jac(j, acc);
    jac_to_affine(r[i], j);
  }
}

// r[i] = a[i] + b[i] as ONE batch-affine run (forward products, one inversion, backward peel): the
// per-thread schedule of k_ba_fwd / k_ba_bwd (csrc/k_msm_big.cu) on the host
void hc_batch_affine(const G1Affine* a, const G1Affine* b, G1Affine* r, Fp* pre, int n) {
  Fp run;
  FpM::set_one(run);
  for (int j = 0; j < n; j++) {
    Fp d;
    ba_denominator(d, a[j], b[j]);
    if (j == 0) run = d;
    else FpM::mul(run, run, d);
    pre[j] = run;
  }
  Fp inv;
  fp_inv(inv, run);
  for (int j = n - 1; j >= 0; j--) {
    Fp invj = inv;
    if (j > 0) {
      Fp d;
      ba_denominator(d, a[j], b[j]);
      FpM::mul(invj, inv, pre[j - 1]);
      FpM::mul(inv, inv, d);
    }
    ba_pair_sum(r[j], a[j], b[j], invj);
  }
}

}  // extern "C"
