// Batch-affine addition: the pair sums of one bucket-accumulation round share field
// inversions through Montgomery's trick (gnark-crypto's MultiExp does the same for its
// large windows — the batch-affine buckets behind msmaccumulator.go:59 and
// common/util.go:75; BatchJacobianToAffineG1, transcript/transcript.go:26, is the same
// trick on Z coordinates).
//
// An affine addition P1 + P2 needs 1 / d with d = x2 - x1 (or 2*y1 for a doubling).  For a
// run of pairs j = 0 .. b-1 the forward pass stores the running products pre_j = d_0 .. d_j,
// ONE inversion gives 1 / pre_{b-1}, and the backward pass peels the factors off again:
//     1 / d_j = inv_run * pre_{j-1},   inv_run *= d_j
// so an addition costs 1 (forward) + 2 (peel) + 3 (lambda, lambda^2, y3) = 6 field products
// plus its share of the inversion, against 10 for an XYZZ mixed addition.
//
// Both passes must see the SAME denominator for a pair, including the exceptional pairs, so
// the denominator is a pure function of the two operands:
//     x1 != x2                 d = x2 - x1     (also when one operand is infinity = (0, 0))
//     x1 == x2, y1 == y2 != 0  d = 2 * y1      (doubling)
//     otherwise                d = 1           (P - P, infinity + infinity, y = 0)
// d is never zero, so a run's product is always invertible.
#pragma once
#include "g1.cuh"

namespace cdl {

// denominator of a pair whose x coordinates are equal
CDL_FN void ba_denominator_equal_x(Fp& d, const Fp& y1, const Fp& y2) {
  if (FpM::eq(y1, y2) && !FpM::is_zero(y1)) FpM::dbl(d, y1);
  else FpM::set_one(d);
}

CDL_FN void ba_denominator(Fp& d, const G1Affine& p1, const G1Affine& p2) {
  FpM::sub(d, p2.x, p1.x);
  if (FpM::is_zero(d)) ba_denominator_equal_x(d, p1.y, p2.y);
}

// r = p1 + p2 given inv = 1 / ba_denominator(p1, p2).  r may alias an operand.
CDL_FN void ba_pair_sum(G1Affine& r, const G1Affine& p1, const G1Affine& p2, const Fp& inv) {
  if (aff_is_inf(p1)) { r = p2; return; }
  if (aff_is_inf(p2)) { r = p1; return; }
  Fp lam, t, x3;
  FpM::sub(t, p2.x, p1.x);
  if (FpM::is_zero(t)) {
    if (!FpM::eq(p1.y, p2.y) || FpM::is_zero(p1.y)) { aff_set_inf(r); return; }
    FpM::sqr(t, p1.x);  // lambda = 3 x1^2 / (2 y1)
    FpM::dbl(lam, t);
    FpM::add(lam, lam, t);
    FpM::mul(lam, lam, inv);
  } else {
    FpM::sub(t, p2.y, p1.y);
    FpM::mul(lam, t, inv);
  }
  FpM::sqr(x3, lam);
  FpM::sub(x3, x3, p1.x);
  FpM::sub(x3, x3, p2.x);
  FpM::sub(t, p1.x, x3);
  FpM::mul(t, t, lam);
  FpM::sub(r.y, t, p1.y);
  r.x = x3;
}

}  // namespace cdl
