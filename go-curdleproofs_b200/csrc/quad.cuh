// Warp-cooperative group law for the LATENCY-bound pieces (one proof at a time, the tail of an MSM):
// four lanes ("a quad") work on ONE point.
//
// A dependent chain of field products runs at ~0.9 us per product when a warp is alone on its
// scheduler, whatever the lane count, so a 124-doubling chain costs 124 x 7 products of latency with
// one thread per point.  Here every lane of a quad holds an identical copy of the point, the
// independent products of one level of the formula's dependency graph are computed one per lane,
// and the results are exchanged with warp shuffles: an XYZZ doubling is 3 levels instead of 9
// sequential products, a mixed or full addition 4 levels instead of 10 / 14.  Additions,
// subtractions and selections are cheap and simply replicated on all four lanes.
//
// Shuffles name only the quad's own lanes in their mask, so quads of one warp may diverge
// (exceptional cases of the group law are uniform WITHIN a quad because its lanes hold the same data).
// Device only: the CPU tier checks the formulas through the ordinary xyzz_* functions, the GPU tier
// through every single-proof parity test (the quad kernels serve the small launches).
#pragma once
#include "g1.cuh"

namespace cdl {

#if defined(__CUDACC__)

struct Quad {
  int lane;       // 0..3 inside the quad
  unsigned mask;  // the quad's four lanes
  __device__ __forceinline__ Quad() {
    const unsigned l = threadIdx.x & 31u;
    lane = (int)(l & 3u);
    mask = 0xfu << (l & ~3u);
  }
};

__device__ __forceinline__ Fp qsel(const Quad& q, const Fp& a0, const Fp& a1, const Fp& a2, const Fp& a3) {
  Fp r;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    uint32_t lo = q.lane & 1 ? a1.v[i] : a0.v[i];
    uint32_t hi = q.lane & 1 ? a3.v[i] : a2.v[i];
    r.v[i] = q.lane & 2 ? hi : lo;
  }
  return r;
}

__device__ __forceinline__ Fp qbcast(const Quad& q, const Fp& mine, int src) {
  Fp r;
#pragma unroll
  for (int i = 0; i < 12; i++) r.v[i] = __shfl_sync(q.mask, mine.v[i], src, 4);
  return r;
}

// r_i = a_i * b_i for i < N (N <= 4), one product latency; every lane receives every result
template <int N>
__device__ __forceinline__ void qmul(const Quad& q, Fp* r, const Fp* a, const Fp* b) {
  const Fp& a3 = N > 3 ? a[3] : a[0];
  const Fp& b3 = N > 3 ? b[3] : b[0];
  const Fp& a2 = N > 2 ? a[2] : a[0];
  const Fp& b2 = N > 2 ? b[2] : b[0];
  const Fp& a1 = N > 1 ? a[1] : a[0];
  const Fp& b1 = N > 1 ? b[1] : b[0];
  Fp x = qsel(q, a[0], a1, a2, a3), y = qsel(q, b[0], b1, b2, b3), p;
  FpM::mul(p, x, y);
#pragma unroll
  for (int i = 0; i < N; i++) r[i] = qbcast(q, p, i);
}

// dbl-2008-s-1 in three levels (2 + 4 + 3 products).  r may alias p.
__device__ __forceinline__ void qxyzz_dbl(const Quad& q, G1Xyzz& r, const G1Xyzz& p) {
  if (xyzz_is_inf(p)) { r = p; return; }
  Fp u, a[4], b[4], o[4];
  FpM::dbl(u, p.y);
  a[0] = u; b[0] = u;        // V = U^2
  a[1] = p.x; b[1] = p.x;    // X^2
  qmul<2>(q, o, a, b);
  const Fp v = o[0];
  Fp m, t;
  FpM::dbl(t, o[1]);
  FpM::add(m, o[1], t);      // M = 3 X^2
  a[0] = u; b[0] = v;        // W = U * V
  a[1] = p.x; b[1] = v;      // S = X * V
  a[2] = m; b[2] = m;        // M^2
  a[3] = v; b[3] = p.zz;     // ZZ3
  qmul<4>(q, o, a, b);
  const Fp w = o[0], s = o[1];
  Fp x3;
  FpM::sub(x3, o[2], s);
  FpM::sub(x3, x3, s);
  const Fp zz3 = o[3];
  FpM::sub(t, s, x3);
  a[0] = m; b[0] = t;        // M * (S - X3)
  a[1] = w; b[1] = p.y;      // W * Y
  a[2] = w; b[2] = p.zzz;    // ZZZ3
  qmul<3>(q, o, a, b);
  r.x = x3;
  FpM::sub(r.y, o[0], o[1]);
  r.zz = zz3;
  r.zzz = o[2];
}

// doubling of an affine point (mdbl-2008-s-1), two levels
__device__ __forceinline__ void qxyzz_dbl_affine(const Quad& q, G1Xyzz& r, const G1Affine& p) {
  Fp u, a[4], b[4], o[4];
  FpM::dbl(u, p.y);
  a[0] = u; b[0] = u;
  a[1] = p.x; b[1] = p.x;
  qmul<2>(q, o, a, b);
  const Fp v = o[0];
  Fp m, t;
  FpM::dbl(t, o[1]);
  FpM::add(m, o[1], t);
  a[0] = u; b[0] = v;        // W
  a[1] = p.x; b[1] = v;      // S
  a[2] = m; b[2] = m;        // M^2
  qmul<3>(q, o, a, b);
  const Fp w = o[0], s = o[1];
  Fp x3;
  FpM::sub(x3, o[2], s);
  FpM::sub(x3, x3, s);
  FpM::sub(t, s, x3);
  a[0] = m; b[0] = t;
  a[1] = w; b[1] = p.y;
  qmul<2>(q, o, a, b);
  r.x = x3;
  FpM::sub(r.y, o[0], o[1]);
  r.zz = v;
  r.zzz = w;
}

// madd-2008-s in four levels (2 + 2 + 3 + 3 products).  r may alias p.
__device__ __forceinline__ void qxyzz_add_mixed(const Quad& q, G1Xyzz& r, const G1Xyzz& p, const G1Affine& s) {
  if (aff_is_inf(s)) { r = p; return; }
  if (xyzz_is_inf(p)) { r.x = s.x; r.y = s.y; FpM::set_one(r.zz); FpM::set_one(r.zzz); return; }
  Fp a[4], b[4], o[4];
  a[0] = s.x; b[0] = p.zz;   // U2
  a[1] = s.y; b[1] = p.zzz;  // S2
  qmul<2>(q, o, a, b);
  Fp pp_, rr;
  FpM::sub(pp_, o[0], p.x);  // P
  FpM::sub(rr, o[1], p.y);   // R
  if (FpM::is_zero(pp_)) {
    if (FpM::is_zero(rr)) { qxyzz_dbl_affine(q, r, s); return; }
    xyzz_set_inf(r);
    return;
  }
  a[0] = pp_; b[0] = pp_;    // PP
  a[1] = rr; b[1] = rr;      // R^2
  qmul<2>(q, o, a, b);
  const Fp pp = o[0], r2 = o[1];
  a[0] = pp; b[0] = pp_;     // PPP
  a[1] = p.x; b[1] = pp;     // Q
  a[2] = p.zz; b[2] = pp;    // ZZ3
  qmul<3>(q, o, a, b);
  const Fp ppp = o[0], qq = o[1], zz3 = o[2];
  Fp x3, t;
  FpM::sub(x3, r2, ppp);
  FpM::sub(x3, x3, qq);
  FpM::sub(x3, x3, qq);
  FpM::sub(t, qq, x3);
  a[0] = rr; b[0] = t;       // R * (Q - X3)
  a[1] = p.y; b[1] = ppp;    // Y1 * PPP
  a[2] = p.zzz; b[2] = ppp;  // ZZZ3
  qmul<3>(q, o, a, b);
  r.x = x3;
  FpM::sub(r.y, o[0], o[1]);
  r.zz = zz3;
  r.zzz = o[2];
}

// add-2008-s in four levels (4 + 4 + 3 + 3 products).  r may alias p or s.
__device__ __forceinline__ void qxyzz_add(const Quad& q, G1Xyzz& r, const G1Xyzz& p, const G1Xyzz& s) {
  if (xyzz_is_inf(s)) { r = p; return; }
  if (xyzz_is_inf(p)) { r = s; return; }
  Fp a[4], b[4], o[4];
  a[0] = p.x; b[0] = s.zz;   // U1
  a[1] = s.x; b[1] = p.zz;   // U2
  a[2] = p.y; b[2] = s.zzz;  // S1
  a[3] = s.y; b[3] = p.zzz;  // S2
  qmul<4>(q, o, a, b);
  const Fp u1 = o[0], s1 = o[2];
  Fp pp_, rr;
  FpM::sub(pp_, o[1], u1);
  FpM::sub(rr, o[3], s1);
  if (FpM::is_zero(pp_)) {
    if (FpM::is_zero(rr)) { qxyzz_dbl(q, r, p); return; }
    xyzz_set_inf(r);
    return;
  }
  a[0] = pp_; b[0] = pp_;      // PP
  a[1] = rr; b[1] = rr;        // R^2
  a[2] = p.zz; b[2] = s.zz;    // ZZ1 * ZZ2
  a[3] = p.zzz; b[3] = s.zzz;  // ZZZ1 * ZZZ2
  qmul<4>(q, o, a, b);
  const Fp pp = o[0], r2 = o[1], zzp = o[2], zzzp = o[3];
  a[0] = pp; b[0] = pp_;       // PPP
  a[1] = u1; b[1] = pp;        // Q
  a[2] = zzp; b[2] = pp;       // ZZ3
  qmul<3>(q, o, a, b);
  const Fp ppp = o[0], qq = o[1], zz3 = o[2];
  Fp x3, t;
  FpM::sub(x3, r2, ppp);
  FpM::sub(x3, x3, qq);
  FpM::sub(x3, x3, qq);
  FpM::sub(t, qq, x3);
  a[0] = rr; b[0] = t;
  a[1] = s1; b[1] = ppp;
  a[2] = zzzp; b[2] = ppp;     // ZZZ3
  qmul<3>(q, o, a, b);
  r.x = x3;
  FpM::sub(r.y, o[0], o[1]);
  r.zz = zz3;
  r.zzz = o[2];
}

// XYZZ -> affine: one inversion (replicated on the four lanes) + one level of products
__device__ __forceinline__ void qxyzz_to_affine(const Quad& q, G1Affine& r, const G1Xyzz& p) {
  if (xyzz_is_inf(p)) { aff_set_inf(r); return; }
  Fp a[4], b[4], o[4];
  a[0] = p.zz; b[0] = p.zzz;
  qmul<1>(q, o, a, b);           // ZZ * ZZZ
  Fp ti;
  fp_inv(ti, o[0]);              // 1 / (ZZ * ZZZ)
  a[0] = ti; b[0] = p.zzz;       // 1 / ZZ
  a[1] = ti; b[1] = p.zz;        // 1 / ZZZ
  qmul<2>(q, o, a, b);
  a[0] = p.x; b[0] = o[0];
  a[1] = p.y; b[1] = o[1];
  qmul<2>(q, o, a, b);
  r.x = o[0];
  r.y = o[1];
}

// r = k * p with the GLV split and signed 4-bit windows (the schedule of jac_scalar_mul_glv), a quad per
// point.  tab: 16 XYZZ entries of shared memory owned by this quad ({1..8}P, then {1..8}phi(P)).
__device__ __forceinline__ void qxyzz_scalar_mul_glv(const Quad& q, G1Xyzz& r, const G1Affine& p, const uint32_t* k,
                                                     G1Xyzz* tab) {
  if (aff_is_inf(p)) { xyzz_set_inf(r); return; }
  Glv g;
  glv_decompose(g, k);
  {
    G1Xyzz t1, t2, t3, t4, t;
    xyzz_from_affine(t1, p);
    qxyzz_dbl_affine(q, t2, p);
    qxyzz_add_mixed(q, t3, t2, p);
    qxyzz_dbl(q, t4, t2);
    if (q.lane == 0) { tab[0] = t1; tab[1] = t2; tab[2] = t3; tab[3] = t4; }
    qxyzz_add_mixed(q, t, t4, p);   // 5P
    if (q.lane == 0) tab[4] = t;
    qxyzz_dbl(q, t, t3);            // 6P
    if (q.lane == 0) tab[5] = t;
    qxyzz_add_mixed(q, t, t, p);    // 7P
    if (q.lane == 0) tab[6] = t;
    qxyzz_dbl(q, t, t4);            // 8P
    if (q.lane == 0) tab[7] = t;
    __syncwarp(q.mask);
    // phi(x, y, zz, zzz) = (beta * x, y, zz, zzz): eight products, two levels
    Fp beta, a[4], b[4], o[4];
    fp_set_beta(beta);
#pragma unroll 1
    for (int h = 0; h < 2; h++) {
#pragma unroll
      for (int i = 0; i < 4; i++) { a[i] = tab[4 * h + i].x; b[i] = beta; }
      qmul<4>(q, o, a, b);
      if (q.lane == 0) {
#pragma unroll 1
        for (int i = 0; i < 4; i++) {
          G1Xyzz e = tab[4 * h + i];
          e.x = o[i];
          tab[8 + 4 * h + i] = e;
        }
      }
    }
    __syncwarp(q.mask);
  }
  int8_t dg[2][32];
  recode_w4_128(dg[0], g.k1);
  recode_w4_128(dg[1], g.k2);
  xyzz_set_inf(r);
#pragma unroll 1
  for (int i = 31; i >= 0; i--) {
    if (i != 31) {
#pragma unroll 1
      for (int j = 0; j < 4; j++) qxyzz_dbl(q, r, r);
    }
#pragma unroll 1
    for (int h = 0; h < 2; h++) {
      const int d = dg[h][i];
      if (d != 0) {
        const int a = d < 0 ? -d : d;
        G1Xyzz t = tab[8 * h + a - 1];
        const bool neg = (d < 0) != (h == 0 ? g.neg1 : g.neg2);
        if (neg) FpM::neg(t.y, t.y);
        qxyzz_add(q, r, r, t);
      }
    }
  }
}

#endif  // __CUDACC__

}  // namespace cdl
