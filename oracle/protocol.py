"""CPU restatement of the curdleproofs protocol orchestration around the G1 hot
path: CRS, ShufflePermuteCommit, Prove / Verify and every sub-argument, the
MSM accumulator and the wire format.  TEST INFRASTRUCTURE ONLY (see
oracle/__init__.py); PARITY unpinned vs a Go run.

Each function cites the /root/reference file:line it follows.  Points are
``None`` (infinity) or affine ``(x, y)``; scalars are ints mod r.  Group
arithmetic goes through a *backend* (pure Python by default, or the C oracle
in oracle/c/ for speed); both are CPU restatements and are cross-checked in
tests/.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

from . import bls12381 as bls
from .merlin import Transcript
from .rand import Rand

R = bls.R
N_BLINDERS = 4  # common/constants.go:3


class ProofError(Exception):
    """Mirrors the reference returning a non-nil error."""


# ----------------------------------------------------------------------------
# group-arithmetic backend
# ----------------------------------------------------------------------------
class PyBackend:
    name = "python-bigint"

    def msm(self, points, scalars):
        return bls.g1_msm(points, scalars)

    def mul(self, pt, k):
        return bls.g1_mul(pt, k)

    def mul_batch(self, pts, ks):
        return bls.g1_mul_batch(pts, ks)

    def add(self, a, b):
        return bls.g1_add(a, b)

    def sub(self, a, b):
        return bls.g1_sub(a, b)

    def fold(self, L, Rr, x):
        """L[i] + x·R[i]  (innerproductargument.go:155-166)."""
        t = self.mul_batch(Rr, [x] * len(Rr))
        return [bls.g1_add(a, b) for a, b in zip(L, t)]


_BACKEND = PyBackend()


def set_backend(b) -> None:
    global _BACKEND
    _BACKEND = b


def get_backend():
    return _BACKEND


def ipa(a, b) -> int:  # common/util.go:26-35
    if len(a) != len(b):
        raise ProofError("IPA: len(a) != len(b)")
    return sum(x * y for x, y in zip(a, b)) % R


def permute(vs, perm):  # common/util.go:37-43
    return [vs[p] for p in perm]


# ----------------------------------------------------------------------------
# CRS  (crs.go:10-59)
# ----------------------------------------------------------------------------
@dataclass
class CRS:
    Gs: list
    Hs: list
    H: object
    Gt: object
    Gu: object
    Gsum: object
    Hsum: object


def generate_crs(size: int, rand: Rand) -> CRS:  # crs.go:20-59
    gs = rand.get_g1_affines(size)
    hs = rand.get_g1_affines(N_BLINDERS)
    h = rand.get_g1_jac()
    gt = rand.get_g1_jac()
    gu = rand.get_g1_jac()
    return CRS(gs, hs, h, gt, gu, bls.g1_sum(gs), bls.g1_sum(hs))


# ----------------------------------------------------------------------------
# ShufflePermuteCommit  (common/util.go:45-88)
# ----------------------------------------------------------------------------
def shuffle_permute_commit(crs_gs, crs_hs, Rs, Ss, perm, k, rand: Rand):
    be = _BACKEND
    Ts = be.mul_batch(Rs, [k] * len(Rs))  # util.go:55-58
    Us = be.mul_batch(Ss, [k] * len(Ss))  # util.go:60-63
    Ts = permute(Ts, perm)
    Us = permute(Us, perm)
    range_frs = [0] * len(crs_gs)
    for i in range(len(perm)):
        range_frs[i] = i
    perm_range = permute(range_frs, perm)
    M = be.msm(crs_gs, perm_range)  # util.go:75
    rs_m = rand.get_frs(N_BLINDERS)
    M2 = be.msm(crs_hs, rs_m)  # util.go:82
    M = be.add(M, M2)
    return Ts, Us, M, rs_m


# ----------------------------------------------------------------------------
# group commitment  (groupcommitment/groupcommitment.go)
# ----------------------------------------------------------------------------
@dataclass
class GroupCommitment:
    T_1: object = None
    T_2: object = None

    @staticmethod
    def new(crsG, crsH, T, r):  # groupcommitment.go:17-31
        be = _BACKEND
        return GroupCommitment(be.mul(crsG, r), be.add(T, be.mul(crsH, r)))

    def add(self, o):  # :33-39
        be = _BACKEND
        return GroupCommitment(be.add(self.T_1, o.T_1), be.add(self.T_2, o.T_2))

    def mul(self, s):  # :41-48
        be = _BACKEND
        return GroupCommitment(be.mul(self.T_1, s), be.mul(self.T_2, s))

    def eq(self, o) -> bool:  # :50-52
        return self.T_1 == o.T_1 and self.T_2 == o.T_2

    def serialize(self, e: bls.Encoder):  # :71-81
        e.point(self.T_1)
        e.point(self.T_2)

    @staticmethod
    def from_reader(d: bls.Decoder):  # :54-69
        return GroupCommitment(d.point(), d.point())


# ----------------------------------------------------------------------------
# MSM accumulator  (msmaccumulator/msmaccumulator.go)
# ----------------------------------------------------------------------------
class MsmAccumulator:
    def __init__(self):  # :16-21
        self.A_c = None
        self.base_scalar = {}

    def accumulate_check(self, C, x, v, rand: Rand):  # :23-47
        if len(v) != len(x):
            raise ProofError("x and v must have the same length")
        alpha = rand.get_fr()
        m = self.base_scalar
        for xi, vi in zip(x, v):
            m[vi] = (m.get(vi, 0) + alpha * xi) % R
        self.A_c = _BACKEND.add(self.A_c, _BACKEND.mul(C, alpha))

    def final_msm_inputs(self):
        pts = list(self.base_scalar.keys())
        return pts, [self.base_scalar[p] for p in pts]

    def verify(self) -> bool:  # :49-64
        pts, xs = self.final_msm_inputs()
        return _BACKEND.msm(pts, xs) == self.A_c


# ----------------------------------------------------------------------------
# inner product argument  (innerproductargument/innerproductargument.go)
# ----------------------------------------------------------------------------
@dataclass
class IPAProof:
    B_c: object = None
    B_d: object = None
    L_Cs: list = field(default_factory=list)
    R_Cs: list = field(default_factory=list)
    L_Ds: list = field(default_factory=list)
    R_Ds: list = field(default_factory=list)
    c0: int = 0
    d0: int = 0

    def serialize(self, e: bls.Encoder):  # :427-460
        e.point(self.B_c)
        e.point(self.B_d)
        e.points(self.L_Cs)
        e.points(self.R_Cs)
        e.points(self.L_Ds)
        e.points(self.R_Ds)
        e.scalar(self.c0)
        e.scalar(self.d0)

    @staticmethod
    def from_reader(d: bls.Decoder):  # :393-425
        p = IPAProof()
        p.B_c = d.point()
        p.B_d = d.point()
        p.L_Cs = d.points()
        p.R_Cs = d.points()
        p.L_Ds = d.points()
        p.R_Ds = d.points()
        p.c0 = d.scalar()
        p.d0 = d.scalar()
        return p


def generate_ipa_blinders(rand: Rand, cs, ds):  # innerproductargument.go:299-391
    n = len(cs)
    rs = rand.get_frs(n)
    zs = rand.get_frs(n - 2)
    omega = (ipa(rs, ds) + ipa(zs[:n - 2], cs[:n - 2])) % R
    delta = ipa(rs[:n - 2], zs[:n - 2])
    inv_c = bls.fr_inv(cs[n - 2])
    t1 = (rs[n - 2] * inv_c % R * omega - delta) % R
    t2 = ((-rs[n - 2]) % R * inv_c % R * cs[n - 1] + rs[n - 1]) % R
    if t2 == 0:
        raise ProofError("last_z_term2 is zero")
    last_z = t1 * bls.fr_inv(t2) % R
    pen_z = (-inv_c) % R * ((last_z * cs[n - 1] + omega) % R) % R
    zs = zs + [pen_z, last_z]
    if (ipa(rs, ds) + ipa(zs, cs)) % R != 0 or ipa(rs, zs) != 0:
        raise ProofError("failed to generate IPA blinders: constraints not satisfied")
    return rs, zs


def ipa_prove(Gs, Gs_prime, Hpt, C, D, z, cs, ds, tr: Transcript, rand: Rand) -> IPAProof:
    """innerproductargument.Prove (:42-188).  Gs, Gs_prime, cs, ds are copied."""
    be = _BACKEND
    Gs, Gs_prime, cs, ds = list(Gs), list(Gs_prime), list(cs), list(ds)
    if len(cs) != len(ds):
        raise ProofError("cs and ds are not the same length")
    if len(cs) & (len(cs) - 1):
        raise ProofError("cs and ds are not a power of two")
    rs_c, rs_d = generate_ipa_blinders(rand, cs, ds)
    B_c = be.msm(Gs, rs_c)  # :65
    B_d = be.msm(Gs_prime, rs_d)  # :69
    tr.append_points(b"ipa_step1", C, D)
    tr.append_scalars(b"ipa_step1", z)
    tr.append_points(b"ipa_step1", B_c, B_d)
    alpha = tr.get_and_append_challenge(b"ipa_alpha")
    beta = tr.get_and_append_challenge(b"ipa_beta")
    n = len(cs)
    cs = [(rs_c[i] + alpha * cs[i]) % R for i in range(n)]
    ds = [(rs_d[i] + alpha * ds[i]) % R for i in range(n)]
    H = be.mul(Hpt, beta)  # :90-91
    proof = IPAProof(B_c=B_c, B_d=B_d)
    while len(cs) > 1:  # :100-172
        n //= 2
        c_L, c_R = cs[:n], cs[n:]
        d_L, d_R = ds[:n], ds[n:]
        G_L, G_R = Gs[:n], Gs[n:]
        Gp_L, Gp_R = Gs_prime[:n], Gs_prime[n:]
        L_C = be.add(be.msm(G_R, c_L), be.mul(H, ipa(c_L, d_R)))  # :108-118
        L_D = be.msm(Gp_L, d_R)  # :121
        R_C = be.add(be.msm(G_L, c_R), be.mul(H, ipa(c_R, d_L)))  # :125-135
        R_D = be.msm(Gp_R, d_L)  # :138
        proof.L_Cs.append(L_C)
        proof.L_Ds.append(L_D)
        proof.R_Cs.append(R_C)
        proof.R_Ds.append(R_D)
        tr.append_points(b"ipa_loop", L_C, L_D, R_C, R_D)
        gamma = tr.get_and_append_challenge(b"ipa_gamma")
        if gamma == 0:
            raise ProofError("ipa gamma challenge is zero")
        gamma_inv = bls.fr_inv(gamma)
        cs = [(c_L[i] + gamma_inv * c_R[i]) % R for i in range(n)]  # :155-166
        ds = [(d_L[i] + gamma * d_R[i]) % R for i in range(n)]
        Gs = be.fold(G_L, G_R, gamma)
        Gs_prime = be.fold(Gp_L, Gp_R, gamma_inv)
    proof.c0 = cs[0]
    proof.d0 = ds[0]
    return proof


def ipa_verify(proof: IPAProof, Gs, Hpt, C, D, z, us, tr: Transcript, acc: MsmAccumulator, rand: Rand) -> bool:
    """innerproductargument.Verify (:190-297)."""
    be = _BACKEND
    tr.append_points(b"ipa_step1", C, D)
    tr.append_scalars(b"ipa_step1", z)
    tr.append_points(b"ipa_step1", proof.B_c, proof.B_d)
    alpha = tr.get_and_append_challenge(b"ipa_alpha")
    beta = tr.get_and_append_challenge(b"ipa_beta")
    n = len(Gs)
    if n & (n - 1):
        raise ProofError("ipa n is not a power of two")
    m = n.bit_length() - 1
    # the reference indexes proof.L_Cs[i] for i<m unchecked (:217, would panic);
    # the boundary defines that as an error.
    if min(len(proof.L_Cs), len(proof.L_Ds), len(proof.R_Cs), len(proof.R_Ds)) < m:
        raise ProofError("ipa proof has too few rounds")
    gamma = []
    for i in range(m):
        tr.append_points(b"ipa_loop", proof.L_Cs[i], proof.L_Ds[i], proof.R_Cs[i], proof.R_Ds[i])
        gamma.append(tr.get_and_append_challenge(b"ipa_gamma"))
    gamma_inv = bls.fr_batch_inv(gamma)
    s = [1] * n
    s_prime = [1] * n
    for i in range(n):  # :223-234
        for j in range(m):
            if i & (1 << j):
                s[i] = s[i] * gamma[m - j - 1] % R
                s_prime[i] = s_prime[i] * gamma_inv[m - j - 1] % R
    # accumulate check 1 (:237-271).  gnark MultiExp errors when len(points) !=
    # len(scalars) — reachable when the proof carries more than m rounds.
    if len(proof.L_Cs) != m or len(proof.R_Cs) != m or len(proof.L_Ds) != m or len(proof.R_Ds) != m:
        raise ProofError("ipa multiexp: length mismatch")
    AC1 = be.msm(proof.L_Cs, gamma)
    AC1 = be.add(AC1, proof.B_c)
    AC1 = be.add(AC1, be.mul(C, alpha))
    betaH = be.mul(Hpt, beta)
    AC1 = be.add(AC1, be.mul(betaH, alpha * alpha % R * z % R))
    AC1 = be.add(AC1, be.msm(proof.R_Cs, gamma_inv))
    GplusH = list(Gs) + [Hpt]
    scalars = [si * proof.c0 % R for si in s] + [beta * proof.d0 % R * proof.c0 % R]
    acc.accumulate_check(AC1, scalars, GplusH, rand)
    # accumulate check 2 (:274-294)
    AC2 = be.msm(proof.L_Ds, gamma)
    AC2 = be.add(AC2, proof.B_d)
    AC2 = be.add(AC2, be.mul(D, alpha))
    AC2 = be.add(AC2, be.msm(proof.R_Ds, gamma_inv))
    scalars = [s_prime[i] * us[i] % R * proof.d0 % R for i in range(n)]
    acc.accumulate_check(AC2, scalars, list(Gs), rand)
    return True


# ----------------------------------------------------------------------------
# grand product argument  (grandproductargument/grandproductargument.go)
# ----------------------------------------------------------------------------
@dataclass
class GPAProof:
    C: object = None
    Rp: int = 0
    ipa: IPAProof = None

    def serialize(self, e):  # :304-318
        e.point(self.C)
        e.scalar(self.Rp)
        self.ipa.serialize(e)

    @staticmethod
    def from_reader(d):  # :288-302
        p = GPAProof()
        p.C = d.point()
        p.Rp = d.scalar()
        p.ipa = IPAProof.from_reader(d)
        return p


def gpa_prove(crs_gs, crs_hs, crs_h, B, result, bs, r_bs, tr: Transcript, rand: Rand) -> GPAProof:
    """grandproductargument.Prove (:42-204)."""
    be = _BACKEND
    ell = len(crs_gs)
    tr.append_points(b"gprod_step1", B)
    tr.append_scalars(b"gprod_step1", result)
    alpha = tr.get_and_append_challenge(b"gprod_alpha")
    cs = [1] * ell
    for i in range(1, ell):
        cs[i] = cs[i - 1] * bs[i - 1] % R
    r_cs = rand.get_frs(len(r_bs))
    C = be.add(be.msm(crs_gs, cs), be.msm(crs_hs, r_cs))  # :66-73
    r_b_plus_alpha = [(x + alpha) % R for x in r_bs]
    r_p = ipa(r_b_plus_alpha, r_cs)
    tr.append_points(b"gprod_step2", C)
    tr.append_scalars(b"gprod_step2", r_p)
    beta = tr.get_and_append_challenge(b"gprod_beta")
    if beta == 0:
        raise ProofError("beta is zero")
    beta_inv = bls.fr_inv(beta)
    # Gs'[i] = beta^-(i+1) Gs[i];  Hs'[i] = beta^-(ell+1) Hs[i]   (:94-103)
    pw = []
    t = beta_inv
    for _ in range(ell):
        pw.append(t)
        t = t * beta_inv % R
    pw += [t] * len(crs_hs)
    prime = be.mul_batch(list(crs_gs) + list(crs_hs), pw)
    Gs_prime, Hs_prime = prime[:ell], prime[ell:]
    bs_prime = []
    tb = beta
    for i in range(ell):
        bs_prime.append(bs[i] * tb % R)
        tb = tb * beta % R
    ds = []
    beta_powers = []
    tb = 1
    for i in range(ell):
        ds.append((bs_prime[i] - tb) % R)
        beta_powers.append(tb)
        tb = tb * beta % R
    # after the loop tb == beta^ell  (used for z_R at :151)
    beta_l1 = pow(beta, ell + 1, R)
    r_ds = [beta_l1 * x % R for x in r_b_plus_alpha]
    ab = [alpha * beta_l1 % R] * len(r_bs)
    D = be.add(be.sub(B, be.msm(Gs_prime, beta_powers)), be.msm(Hs_prime, ab))  # :131-138
    Gs = list(crs_gs) + list(crs_hs)
    Gs_prime = Gs_prime + Hs_prime
    z = (r_p * beta_l1 + result * tb - 1) % R
    cs = cs + r_cs
    ds = ds + r_ds
    if ipa(cs, ds) != z:
        raise ProofError("IPA(C, D) != z")
    if be.msm(Gs, cs) != C:  # :164-170
        raise ProofError("msm(G, c) != C")
    if be.msm(Gs_prime, ds) != D:  # :171-177
        raise ProofError("msm(G', d) != D")
    ipa_proof = ipa_prove(Gs, Gs_prime, crs_h, C, D, z, cs, ds, tr, rand)
    return GPAProof(C, r_p, ipa_proof)


def gpa_verify(proof: GPAProof, crs_gs, crs_hs, crs_h, Gsum, Hsum, B, result, num_blinders, tr, acc, rand) -> bool:
    """grandproductargument.Verify (:206-286)."""
    be = _BACKEND
    ell = len(crs_gs)
    tr.append_points(b"gprod_step1", B)
    tr.append_scalars(b"gprod_step1", result)
    alpha = tr.get_and_append_challenge(b"gprod_alpha")
    tr.append_points(b"gprod_step2", proof.C)
    tr.append_scalars(b"gprod_step2", proof.Rp)
    beta = tr.get_and_append_challenge(b"gprod_beta")
    if beta == 0:
        raise ProofError("beta is zero")
    beta_inv = bls.fr_inv(beta)
    us = []
    t = beta_inv
    for _ in range(ell):
        us.append(t)
        t = t * beta_inv % R
    us += [t] * num_blinders
    D_M = be.mul(Gsum, beta_inv)  # :244
    D_R = be.mul(Hsum, alpha)  # :245
    D = be.add(be.sub(B, D_M), D_R)  # :246
    Gs = list(crs_gs) + list(crs_hs)
    beta_l = pow(beta, ell, R)
    beta_l1 = beta_l * beta % R
    z = (result * beta_l + proof.Rp * beta_l1 - 1) % R
    return ipa_verify(proof.ipa, Gs, crs_h, proof.C, D, z, us, tr, acc, rand)


# ----------------------------------------------------------------------------
# same permutation argument  (samepermutationargument/samepermutationargument.go)
# ----------------------------------------------------------------------------
@dataclass
class SamePermProof:
    B: object = None
    gpa: GPAProof = None

    def serialize(self, e):  # :180-192
        e.point(self.B)
        self.gpa.serialize(e)

    @staticmethod
    def from_reader(d):  # :166-178
        p = SamePermProof()
        p.B = d.point()
        p.gpa = GPAProof.from_reader(d)
        return p


def sameperm_prove(crs_gs, crs_hs, crs_h, A, M, as_, perm, rs_a, rs_m, tr, rand) -> SamePermProof:
    """samepermutationargument.Prove (:32-101)."""
    be = _BACKEND
    tr.append_points(b"same_perm_step1", A, M)
    tr.append_scalars(b"same_perm_step1", *as_)
    alpha = tr.get_and_append_challenge(b"same_perm_alpha")
    beta = tr.get_and_append_challenge(b"same_perm_beta")
    permuted_as = permute(as_, perm)
    bs = []
    p = 1
    for i in range(len(permuted_as)):
        b = (alpha * perm[i] + permuted_as[i] + beta) % R
        bs.append(b)
        p = p * b % R
    msm_betas = be.msm(crs_gs, [beta] * len(crs_gs))  # :62-69
    B = be.add(be.add(A, be.mul(M, alpha)), msm_betas)  # :70-73
    rs_b = [(alpha * rs_m[i] + rs_a[i]) % R for i in range(len(rs_a))]
    gpa = gpa_prove(crs_gs, crs_hs, crs_h, B, p, bs, rs_b, tr, rand)
    return SamePermProof(B, gpa)


def sameperm_verify(proof: SamePermProof, crs_gs, crs_hs, crs_h, Gsum, Hsum, A, M, as_, num_blinders, tr, acc, rand) -> bool:
    """samepermutationargument.Verify (:103-164)."""
    be = _BACKEND
    tr.append_points(b"same_perm_step1", A, M)
    tr.append_scalars(b"same_perm_step1", *as_)
    alpha = tr.get_and_append_challenge(b"same_perm_alpha")
    beta = tr.get_and_append_challenge(b"same_perm_beta")
    p = 1
    for i in range(len(as_)):
        p = p * ((i * alpha + beta + as_[i]) % R) % R
    betas = [beta] * len(crs_gs)
    C = be.sub(be.sub(proof.B, A), be.mul(M, alpha))  # :136-138
    acc.accumulate_check(C, betas, list(crs_gs), rand)  # :140
    return gpa_verify(proof.gpa, crs_gs, crs_hs, crs_h, Gsum, Hsum, proof.B, p, num_blinders, tr, acc, rand)


# ----------------------------------------------------------------------------
# same scalar argument  (samescalarargument/samescalarargument.go)
# ----------------------------------------------------------------------------
@dataclass
class SameScalarProof:
    A: GroupCommitment = None
    B: GroupCommitment = None
    Z_k: int = 0
    Z_t: int = 0
    Z_u: int = 0

    def serialize(self, e):  # :121-140
        self.A.serialize(e)
        self.B.serialize(e)
        e.scalar(self.Z_k)
        e.scalar(self.Z_t)
        e.scalar(self.Z_u)

    @staticmethod
    def from_reader(d):  # :102-119
        p = SameScalarProof()
        p.A = GroupCommitment.from_reader(d)
        p.B = GroupCommitment.from_reader(d)
        p.Z_k = d.scalar()
        p.Z_t = d.scalar()
        p.Z_u = d.scalar()
        return p


def samescalar_prove(Gt, Gu, H, Rp, Sp, T: GroupCommitment, U: GroupCommitment, k, r_t, r_u, tr, rand) -> SameScalarProof:
    """samescalarargument.Prove (:34-81)."""
    be = _BACKEND
    r_a = rand.get_fr()
    r_b = rand.get_fr()
    r_k = rand.get_fr()
    A = GroupCommitment.new(Gt, H, be.mul(Rp, r_k), r_a)
    B = GroupCommitment.new(Gu, H, be.mul(Sp, r_k), r_b)
    tr.append_points(b"sameexp_points", Rp, Sp, T.T_1, T.T_2, U.T_1, U.T_2, A.T_1, A.T_2, B.T_1, B.T_2)
    alpha = tr.get_and_append_challenge(b"sameexp_alpha")
    return SameScalarProof(A, B, (r_k + k * alpha) % R, (r_a + r_t * alpha) % R, (r_b + r_u * alpha) % R)


def samescalar_verify(proof: SameScalarProof, Gt, Gu, H, Rp, Sp, T, U, tr) -> bool:
    """samescalarargument.Verify (:83-100)."""
    be = _BACKEND
    tr.append_points(b"sameexp_points", Rp, Sp, T.T_1, T.T_2, U.T_1, U.T_2,
                     proof.A.T_1, proof.A.T_2, proof.B.T_1, proof.B.T_2)
    alpha = tr.get_and_append_challenge(b"sameexp_alpha")
    e1 = GroupCommitment.new(Gt, H, be.mul(Rp, proof.Z_k), proof.Z_t)
    e2 = GroupCommitment.new(Gu, H, be.mul(Sp, proof.Z_k), proof.Z_u)
    return proof.A.add(T.mul(alpha)).eq(e1) and proof.B.add(U.mul(alpha)).eq(e2)


# ----------------------------------------------------------------------------
# same multiscalar argument  (samemultiscalarargument/samemultiscalarargument.go)
# ----------------------------------------------------------------------------
MAX_RECURSIVE_STEPS = 32  # :237


@dataclass
class SameMSMProof:
    B_a: object = None
    B_t: object = None
    B_u: object = None
    L_A: list = field(default_factory=list)
    L_T: list = field(default_factory=list)
    L_U: list = field(default_factory=list)
    R_A: list = field(default_factory=list)
    R_T: list = field(default_factory=list)
    R_U: list = field(default_factory=list)
    x: int = 0

    def serialize(self, e):  # :323-365
        e.point(self.B_a)
        e.point(self.B_t)
        e.point(self.B_u)
        for v in (self.L_A, self.L_T, self.L_U, self.R_A, self.R_T, self.R_U):
            e.points(v)
        e.scalar(self.x)

    @staticmethod
    def from_reader(d):  # :282-321
        p = SameMSMProof()
        p.B_a = d.point()
        p.B_t = d.point()
        p.B_u = d.point()
        p.L_A = d.points()
        p.L_T = d.points()
        p.L_U = d.points()
        p.R_A = d.points()
        p.R_T = d.points()
        p.R_U = d.points()
        p.x = d.scalar()
        return p


def samemsm_prove(G, A, Z_t, Z_u, T, U, x, tr, rand) -> SameMSMProof:
    """samemultiscalarargument.Prove (:37-157).  G, T, U, x are copied."""
    be = _BACKEND
    G, T, U, x = list(G), list(T), list(U), list(x)
    n = len(x)
    r = rand.get_frs(n)
    proof = SameMSMProof()
    proof.B_a = be.msm(G, r)  # :63-72
    proof.B_t = be.msm(T, r)
    proof.B_u = be.msm(U, r)
    tr.append_points(b"same_msm_step1", A, Z_t, Z_u)
    tr.append_points(b"same_msm_step1", *T)
    tr.append_points(b"same_msm_step1", *U)
    tr.append_points(b"same_msm_step1", proof.B_a, proof.B_t, proof.B_u)
    alpha = tr.get_and_append_challenge(b"same_msm_alpha")
    x = [(r[i] + x[i] * alpha) % R for i in range(n)]
    while len(x) > 1:  # :85-140
        n //= 2
        x_L, x_R = x[:n], x[n:]
        T_L, T_R = T[:n], T[n:]
        U_L, U_R = U[:n], U[n:]
        G_L, G_R = G[:n], G[n:]
        L_A = be.msm(G_R, x_L)
        L_T = be.msm(T_R, x_L)
        L_U = be.msm(U_R, x_L)
        R_A = be.msm(G_L, x_R)
        R_T = be.msm(T_L, x_R)
        R_U = be.msm(U_L, x_R)
        proof.L_A.append(L_A)
        proof.L_T.append(L_T)
        proof.L_U.append(L_U)
        proof.R_A.append(R_A)
        proof.R_T.append(R_T)
        proof.R_U.append(R_U)
        tr.append_points(b"same_msm_loop", L_A, L_T, L_U, R_A, R_T, R_U)
        gamma = tr.get_and_append_challenge(b"same_msm_gamma")
        if gamma == 0:
            raise ProofError("gamma is zero")
        gamma_inv = bls.fr_inv(gamma)
        x = [(x_L[i] + gamma_inv * x_R[i]) % R for i in range(n)]  # :129-135
        T = be.fold(T_L, T_R, gamma)
        U = be.fold(U_L, U_R, gamma)
        G = be.fold(G_L, G_R, gamma)
    proof.x = x[0]
    return proof


def _unfolded_scalars(proof: SameMSMProof, n: int, tr):  # :239-280
    lg_n = len(proof.L_A)
    if lg_n >= MAX_RECURSIVE_STEPS:
        raise ProofError("recursive steps greater than expected")
    if n != (1 << lg_n):
        raise ProofError("must by log2(L_a)")
    # the reference indexes the other five slices unchecked (:255-262)
    if any(len(v) != lg_n for v in (proof.L_T, proof.L_U, proof.R_A, proof.R_T, proof.R_U)):
        raise ProofError("same msm proof: inconsistent round counts")
    ch = []
    for i in range(lg_n):
        tr.append_points(b"same_msm_loop", proof.L_A[i], proof.L_T[i], proof.L_U[i],
                         proof.R_A[i], proof.R_T[i], proof.R_U[i])
        ch.append(tr.get_and_append_challenge(b"same_msm_gamma"))
    ss = []
    for i in range(n):
        t = 1
        for k in range(lg_n - 1, -1, -1):
            if i & (1 << (lg_n - k - 1)):
                t = t * ch[k] % R
        ss.append(t)
    return ch, bls.fr_batch_inv(ch), ss


def samemsm_verify(proof: SameMSMProof, G, A, Z_t, Z_u, T, U, tr, acc, rand) -> bool:
    """samemultiscalarargument.Verify (:159-235)."""
    be = _BACKEND
    n = len(T)
    tr.append_points(b"same_msm_step1", A, Z_t, Z_u)
    tr.append_points(b"same_msm_step1", *T)
    tr.append_points(b"same_msm_step1", *U)
    tr.append_points(b"same_msm_step1", proof.B_a, proof.B_t, proof.B_u)
    alpha = tr.get_and_append_challenge(b"same_msm_alpha")
    gamma, gamma_inv, s = _unfolded_scalars(proof, n, tr)
    xs = [proof.x * si % R for si in s]
    A_a = be.add(proof.B_a, be.mul(A, alpha))
    Zt_a = be.add(proof.B_t, be.mul(Z_t, alpha))
    Zu_a = be.add(proof.B_u, be.mul(Z_u, alpha))
    p = be.add(be.add(A_a, be.msm(proof.L_A, gamma)), be.msm(proof.R_A, gamma_inv))
    acc.accumulate_check(p, xs, list(G), rand)  # :206
    p = be.add(be.add(Zt_a, be.msm(proof.L_T, gamma)), be.msm(proof.R_T, gamma_inv))
    acc.accumulate_check(p, xs, list(T), rand)  # :218
    p = be.add(be.add(Zu_a, be.msm(proof.L_U, gamma)), be.msm(proof.R_U, gamma_inv))
    acc.accumulate_check(p, xs, list(U), rand)  # :231
    return True


# ----------------------------------------------------------------------------
# top level  (curdleproof.go)
# ----------------------------------------------------------------------------
@dataclass
class Proof:
    A: object = None
    T: GroupCommitment = None
    U: GroupCommitment = None
    R: object = None
    S: object = None
    same_perm: SamePermProof = None
    same_scalar: SameScalarProof = None
    same_msm: SameMSMProof = None

    def serialize(self) -> bytes:  # curdleproof.go:358-387
        e = bls.Encoder()
        self.serialize_into(e)
        return e.bytes()

    def serialize_into(self, e):
        e.point(self.A)
        self.T.serialize(e)
        self.U.serialize(e)
        e.point(self.R)
        e.point(self.S)
        self.same_perm.serialize(e)
        self.same_scalar.serialize(e)
        self.same_msm.serialize(e)

    @staticmethod
    def from_reader(d: bls.Decoder):  # curdleproof.go:320-356
        p = Proof()
        p.A = d.point()
        p.T = GroupCommitment.from_reader(d)
        p.U = GroupCommitment.from_reader(d)
        p.R = d.point()
        p.S = d.point()
        p.same_perm = SamePermProof.from_reader(d)
        p.same_scalar = SameScalarProof.from_reader(d)
        p.same_msm = SameMSMProof.from_reader(d)
        return p

    @staticmethod
    def deserialize(b: bytes):
        return Proof.from_reader(bls.Decoder(b))


def prove(crs: CRS, Rs, Ss, Ts, Us, M, perm, k, rs_m, rand: Rand) -> Proof:
    """curdleproof.Prove (curdleproof.go:38-197)."""
    be = _BACKEND
    tr = Transcript(b"curdleproofs")
    tr.append_points(b"curdleproofs_step1", *Rs)
    tr.append_points(b"curdleproofs_step1", *Ss)
    tr.append_points(b"curdleproofs_step1", *Ts)
    tr.append_points(b"curdleproofs_step1", *Us)
    tr.append_points(b"curdleproofs_step1", M)
    as_ = tr.get_and_append_challenges(b"curdleproofs_vec_a", len(Rs))
    rs_a = rand.get_frs(N_BLINDERS - 2)
    rs_a_prime = rs_a + [0, 0]
    perm_as = permute(as_, perm)
    A = be.add(be.msm(crs.Gs, perm_as), be.msm(crs.Hs, rs_a_prime))  # :72-79
    same_perm = sameperm_prove(crs.Gs, crs.Hs, crs.H, A, M, as_, perm, rs_a_prime, rs_m, tr, rand)
    r_t = rand.get_fr()
    r_u = rand.get_fr()
    Rp = be.msm(Rs, as_)  # :109-116
    Sp = be.msm(Ss, as_)
    T = GroupCommitment.new(crs.Gt, crs.H, be.mul(Rp, k), r_t)  # :118-122
    U = GroupCommitment.new(crs.Gu, crs.H, be.mul(Sp, k), r_u)
    same_scalar = samescalar_prove(crs.Gt, crs.Gu, crs.H, Rp, Sp, T, U, k, r_t, r_u, tr, rand)
    A_prime = be.add(be.add(A, T.T_1), U.T_1)  # :146-148
    G = list(crs.Gs) + list(crs.Hs[:N_BLINDERS - 2]) + [crs.Gt, crs.Gu]
    T_prime = list(Ts) + [None, None, crs.H, None]
    U_prime = list(Us) + [None, None, None, crs.H]
    x = perm_as + rs_a + [r_t, r_u]
    same_msm = samemsm_prove(G, A_prime, T.T_2, U.T_2, T_prime, U_prime, x, tr, rand)
    return Proof(A, T, U, Rp, Sp, same_perm, same_scalar, same_msm)


def verify_with_accumulator(proof: Proof, crs: CRS, Rs, Ss, Ts, Us, M, rand: Rand):
    """curdleproof.Verify up to (not including) the final accumulator MSM.

    Returns (early_verdict, accumulator): early_verdict is False when a
    sub-argument already rejected (accumulator then None)."""
    be = _BACKEND
    tr = Transcript(b"curdleproofs")
    acc = MsmAccumulator()
    if len(Ts) == 0 or Ts[0] is None:  # :213-215
        raise ProofError("randomizer is zero")
    tr.append_points(b"curdleproofs_step1", *Rs)
    tr.append_points(b"curdleproofs_step1", *Ss)
    tr.append_points(b"curdleproofs_step1", *Ts)
    tr.append_points(b"curdleproofs_step1", *Us)
    tr.append_points(b"curdleproofs_step1", M)
    as_ = tr.get_and_append_challenges(b"curdleproofs_vec_a", len(Rs))
    if not sameperm_verify(proof.same_perm, crs.Gs, crs.Hs, crs.H, crs.Gsum, crs.Hsum,
                           proof.A, M, as_, N_BLINDERS, tr, acc, rand):
        return False, None
    if not samescalar_verify(proof.same_scalar, crs.Gt, crs.Gu, crs.H, proof.R, proof.S, proof.T, proof.U, tr):
        return False, None
    A_prime = be.add(be.add(proof.A, proof.T.T_1), proof.U.T_1)
    Gs = list(crs.Gs) + list(crs.Hs[:N_BLINDERS - 2]) + [crs.Gt, crs.Gu]
    T_prime = list(Ts) + [None, None, crs.H, None]
    U_prime = list(Us) + [None, None, None, crs.H]
    if not samemsm_verify(proof.same_msm, Gs, A_prime, proof.T.T_2, proof.U.T_2, T_prime, U_prime, tr, acc, rand):
        return False, None
    acc.accumulate_check(proof.R, as_, list(Rs), rand)  # :306
    acc.accumulate_check(proof.S, as_, list(Ss), rand)  # :309
    return True, acc


def verify(proof: Proof, crs: CRS, Rs, Ss, Ts, Us, M, rand: Rand) -> bool:
    """curdleproof.Verify (curdleproof.go:199-318).  Returns the verdict;
    raises ProofError where the reference returns (false, err)."""
    ok, acc = verify_with_accumulator(proof, crs, Rs, Ss, Ts, Us, M, rand)
    if not ok:
        return False
    return acc.verify()  # :313-317
