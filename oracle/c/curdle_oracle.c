/* CPU oracle, plain C: BLS12-381 G1 arithmetic on 6 x 64-bit Montgomery limbs.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker for the CUDA
 * path at sizes pure Python cannot reach, and the timed "port" CPU baseline in
 * bench.py.  It is never linked into or called by the product library.
 * PARITY: unpinned vs a Go run (the reference's arithmetic lives in
 * gnark-crypto v0.11.0, go.mod:6, absent from /root/reference); pinned to the
 * pure-Python restatement oracle/bls12381.py (itself pinned to public KATs) by
 * tests/test_oracle_c.py.
 *
 * Restates, independently of the CUDA code (different limb width, plain
 * Jacobian formulas, unsigned-window Pippenger):
 *   G1Jac.MultiExp                     -> co_g1_msm       (e.g. msmaccumulator.go:59)
 *   G1Affine.ScalarMultiplication xN   -> co_g1_mul_batch (common/util.go:55-63,
 *                                                          grandproductargument.go:94-103)
 *   L[i] += x*R[i]                     -> co_g1_fold      (innerproductargument.go:155-166,
 *                                                          samemultiscalarargument.go:129-135)
 *   G1Affine.Bytes / SetBytes          -> co_g1_compress / co_g1_decompress
 *                                                         (whisk/types.go:79-95)
 *   Keccak-f[1600]                     -> co_keccak_f1600 (under jsign/merlin, go.mod:7)
 *
 * Interface: coordinates and scalars are canonical little-endian byte strings
 * (48 / 32 bytes); a point is x || y (96 bytes), infinity is 96 zero bytes.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[6]; } fp;
typedef struct { fp x, y, z; } jac;   /* z == 0: infinity */
typedef struct { fp x, y; } aff;      /* (0,0): infinity */

static const fp FP_P = {{0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull,
                         0x64774b84f38512bfull, 0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull}};
static const fp FP_ONE = {{0x760900000002fffdull, 0xebf4000bc40c0002ull, 0x5f48985753c758baull,
                           0x77ce585370525745ull, 0x5c071a97a256ec6dull, 0x15f65ec3fa80e493ull}};
static const fp FP_R2 = {{0xf4df1f341c341746ull, 0x0a76e6a609d104f1ull, 0x8de5476c4c95b6d5ull,
                          0x67eb88a9939d83c0ull, 0x9a793e85b519952dull, 0x11988fe592cae3aaull}};
static const uint64_t FP_INV = 0x89f3fffcfffcfffdull; /* -p^-1 mod 2^64 */
static const uint64_t FR_R[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull,
                                 0x73eda753299d7d48ull};

static int fp_is_zero(const fp* a) {
  uint64_t o = 0;
  for (int i = 0; i < 6; i++) o |= a->l[i];
  return o == 0;
}
static int fp_eq(const fp* a, const fp* b) { return memcmp(a, b, sizeof(fp)) == 0; }
static int fp_geq_p(const fp* a) {
  for (int i = 5; i >= 0; i--) {
    if (a->l[i] > FP_P.l[i]) return 1;
    if (a->l[i] < FP_P.l[i]) return 0;
  }
  return 1;
}
static void fp_sub_p(fp* a) {
  uint64_t borrow = 0;
  for (int i = 0; i < 6; i++) {
    u128 t = (u128)a->l[i] - FP_P.l[i] - borrow;
    a->l[i] = (uint64_t)t;
    borrow = (uint64_t)(t >> 64) & 1;
  }
}
static void fp_add(fp* r, const fp* a, const fp* b) {
  uint64_t c = 0;
  for (int i = 0; i < 6; i++) {
    u128 t = (u128)a->l[i] + b->l[i] + c;
    r->l[i] = (uint64_t)t;
    c = (uint64_t)(t >> 64);
  }
  if (fp_geq_p(r)) fp_sub_p(r);
}
static void fp_sub(fp* r, const fp* a, const fp* b) {
  uint64_t borrow = 0;
  for (int i = 0; i < 6; i++) {
    u128 t = (u128)a->l[i] - b->l[i] - borrow;
    r->l[i] = (uint64_t)t;
    borrow = (uint64_t)(t >> 64) & 1;
  }
  if (borrow) {
    uint64_t c = 0;
    for (int i = 0; i < 6; i++) {
      u128 t = (u128)r->l[i] + FP_P.l[i] + c;
      r->l[i] = (uint64_t)t;
      c = (uint64_t)(t >> 64);
    }
  }
}
static void fp_neg(fp* r, const fp* a) {
  if (fp_is_zero(a)) { *r = *a; return; }
  fp z = {{0}};
  fp_sub(r, &z, a);
}
/* Montgomery product: operand-scanning CIOS on 64-bit words, fully unrolled and
 * register resident.  p < 2^381 leaves the top word with spare bits, so the
 * running sum never needs a 8th word ("no-carry" variant). */
#define CO_MAC(lo, hi, x, y, add1, add2)                      \
  do {                                                        \
    u128 t__ = (u128)(x) * (y) + (add1) + (add2);             \
    (lo) = (uint64_t)t__;                                     \
    (hi) = (uint64_t)(t__ >> 64);                             \
  } while (0)
#define CO_ROW(bi_)                                           \
  do {                                                        \
    uint64_t bi = (bi_), c, m, d, junk;                       \
    CO_MAC(t0, c, a0, bi, t0, 0);                             \
    CO_MAC(t1, c, a1, bi, t1, c);                             \
    CO_MAC(t2, c, a2, bi, t2, c);                             \
    CO_MAC(t3, c, a3, bi, t3, c);                             \
    CO_MAC(t4, c, a4, bi, t4, c);                             \
    CO_MAC(t5, c, a5, bi, t5, c);                             \
    t6 = c;                                                   \
    m = t0 * FP_INV;                                          \
    CO_MAC(junk, d, m, p0, t0, 0);                            \
    (void)junk;                                               \
    CO_MAC(t0, d, m, p1, t1, d);                              \
    CO_MAC(t1, d, m, p2, t2, d);                              \
    CO_MAC(t2, d, m, p3, t3, d);                              \
    CO_MAC(t3, d, m, p4, t4, d);                              \
    CO_MAC(t4, d, m, p5, t5, d);                              \
    t5 = t6 + d;                                              \
  } while (0)
static void fp_mul(fp* r, const fp* a, const fp* b) {
  const uint64_t a0 = a->l[0], a1 = a->l[1], a2 = a->l[2], a3 = a->l[3], a4 = a->l[4], a5 = a->l[5];
  const uint64_t p0 = FP_P.l[0], p1 = FP_P.l[1], p2 = FP_P.l[2], p3 = FP_P.l[3], p4 = FP_P.l[4], p5 = FP_P.l[5];
  uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0, t6 = 0;
  CO_ROW(b->l[0]);
  CO_ROW(b->l[1]);
  CO_ROW(b->l[2]);
  CO_ROW(b->l[3]);
  CO_ROW(b->l[4]);
  CO_ROW(b->l[5]);
  r->l[0] = t0; r->l[1] = t1; r->l[2] = t2; r->l[3] = t3; r->l[4] = t4; r->l[5] = t5;
  if (fp_geq_p(r)) fp_sub_p(r);
}
static void fp_sqr(fp* r, const fp* a) { fp_mul(r, a, a); }
static void fp_pow(fp* r, const fp* a, const uint64_t* e, int nwords) {
  fp acc = FP_ONE, base = *a;
  int started = 0;
  for (int w = nwords - 1; w >= 0; w--)
    for (int bit = 63; bit >= 0; bit--) {
      if (started) fp_sqr(&acc, &acc);
      if ((e[w] >> bit) & 1) {
        if (started) fp_mul(&acc, &acc, &base); else { acc = base; started = 1; }
      }
    }
  *r = acc;
}
static void fp_inv(fp* r, const fp* a) {
  uint64_t e[6];
  memcpy(e, FP_P.l, sizeof e);
  e[0] -= 2;
  fp_pow(r, a, e, 6);
}
static void fp_to_mont(fp* r, const fp* a) { fp_mul(r, a, &FP_R2); }
static void fp_from_mont(fp* r, const fp* a) {
  fp one = {{1, 0, 0, 0, 0, 0}};
  fp_mul(r, a, &one);
}
static int fp_sqrt(fp* r, const fp* a) { /* p = 3 mod 4 */
  uint64_t e[6];
  /* (p+1)/4 */
  uint64_t c = 1;
  for (int i = 0; i < 6; i++) { u128 t = (u128)FP_P.l[i] + c; e[i] = (uint64_t)t; c = (uint64_t)(t >> 64); }
  for (int i = 0; i < 6; i++) e[i] = (e[i] >> 2) | (i < 5 ? e[i + 1] << 62 : 0);
  fp s, t;
  fp_pow(&s, a, e, 6);
  fp_sqr(&t, &s);
  *r = s;
  return fp_eq(&t, a);
}

/* ---------------------------------------------------------------- group */
static void jac_set_inf(jac* p) { p->x = FP_ONE; p->y = FP_ONE; memset(&p->z, 0, sizeof(fp)); }
static int jac_is_inf(const jac* p) { return fp_is_zero(&p->z); }
static int aff_is_inf(const aff* p) { return fp_is_zero(&p->x) && fp_is_zero(&p->y); }

static void jac_dbl(jac* r, const jac* p) {
  if (jac_is_inf(p)) { *r = *p; return; }
  fp a, b, c, d, e, f, t;
  fp_sqr(&a, &p->x);
  fp_sqr(&b, &p->y);
  fp_sqr(&c, &b);
  fp_add(&t, &p->x, &b);
  fp_sqr(&t, &t);
  fp_sub(&t, &t, &a);
  fp_sub(&t, &t, &c);
  fp_add(&d, &t, &t);
  fp_add(&e, &a, &a);
  fp_add(&e, &e, &a);
  fp_sqr(&f, &e);
  fp z3;
  fp_mul(&z3, &p->y, &p->z);
  fp_add(&z3, &z3, &z3);
  fp x3;
  fp_sub(&x3, &f, &d);
  fp_sub(&x3, &x3, &d);
  fp_sub(&t, &d, &x3);
  fp_mul(&t, &t, &e);
  fp_add(&c, &c, &c);
  fp_add(&c, &c, &c);
  fp_add(&c, &c, &c);
  fp_sub(&r->y, &t, &c);
  r->x = x3;
  r->z = z3;
}
static void jac_add(jac* r, const jac* p, const jac* q) {
  if (jac_is_inf(p)) { *r = *q; return; }
  if (jac_is_inf(q)) { *r = *p; return; }
  fp z1z1, z2z2, u1, u2, s1, s2, h, rr, hh, hhh, v, t;
  fp_sqr(&z1z1, &p->z);
  fp_sqr(&z2z2, &q->z);
  fp_mul(&u1, &p->x, &z2z2);
  fp_mul(&u2, &q->x, &z1z1);
  fp_mul(&s1, &p->y, &q->z);
  fp_mul(&s1, &s1, &z2z2);
  fp_mul(&s2, &q->y, &p->z);
  fp_mul(&s2, &s2, &z1z1);
  if (fp_eq(&u1, &u2)) {
    if (fp_eq(&s1, &s2)) { jac_dbl(r, p); return; }
    jac_set_inf(r);
    return;
  }
  fp_sub(&h, &u2, &u1);
  fp_sub(&rr, &s2, &s1);
  fp_sqr(&hh, &h);
  fp_mul(&hhh, &hh, &h);
  fp_mul(&v, &u1, &hh);
  fp x3, y3, z3;
  fp_sqr(&x3, &rr);
  fp_sub(&x3, &x3, &hhh);
  fp_sub(&x3, &x3, &v);
  fp_sub(&x3, &x3, &v);
  fp_sub(&t, &v, &x3);
  fp_mul(&y3, &rr, &t);
  fp_mul(&t, &s1, &hhh);
  fp_sub(&y3, &y3, &t);
  fp_mul(&z3, &p->z, &q->z);
  fp_mul(&z3, &z3, &h);
  r->x = x3; r->y = y3; r->z = z3;
}
static void jac_from_aff(jac* r, const aff* p) {
  if (aff_is_inf(p)) { jac_set_inf(r); return; }
  r->x = p->x; r->y = p->y; r->z = FP_ONE;
}
/* r = p + q, q affine (8M + 3S) */
static void jac_add_aff(jac* r, const jac* p, const aff* q) {
  if (aff_is_inf(q)) { *r = *p; return; }
  if (jac_is_inf(p)) { jac_from_aff(r, q); return; }
  fp z1z1, u2, s2, h, rr, hh, hhh, v, t;
  fp_sqr(&z1z1, &p->z);
  fp_mul(&u2, &q->x, &z1z1);
  fp_mul(&s2, &q->y, &p->z);
  fp_mul(&s2, &s2, &z1z1);
  if (fp_eq(&u2, &p->x)) {
    if (fp_eq(&s2, &p->y)) { jac_dbl(r, p); return; }
    jac_set_inf(r);
    return;
  }
  fp_sub(&h, &u2, &p->x);
  fp_sub(&rr, &s2, &p->y);
  fp_sqr(&hh, &h);
  fp_mul(&hhh, &hh, &h);
  fp_mul(&v, &p->x, &hh);
  fp x3, y3, z3;
  fp_sqr(&x3, &rr);
  fp_sub(&x3, &x3, &hhh);
  fp_sub(&x3, &x3, &v);
  fp_sub(&x3, &x3, &v);
  fp_sub(&t, &v, &x3);
  fp_mul(&y3, &rr, &t);
  fp_mul(&t, &p->y, &hhh);
  fp_sub(&y3, &y3, &t);
  fp_mul(&z3, &p->z, &h);
  r->x = x3; r->y = y3; r->z = z3;
}
/* batch Jacobian -> affine (Montgomery's trick) */
static void jac_batch_to_aff(aff* out, const jac* in, size_t n) {
  if (n == 0) return;
  fp* pref = (fp*)malloc((n + 1) * sizeof(fp));
  pref[0] = FP_ONE;
  for (size_t i = 0; i < n; i++) {
    if (jac_is_inf(&in[i])) pref[i + 1] = pref[i]; else fp_mul(&pref[i + 1], &pref[i], &in[i].z);
  }
  fp inv;
  fp_inv(&inv, &pref[n]);
  for (size_t i = n; i-- > 0;) {
    if (jac_is_inf(&in[i])) { memset(&out[i], 0, sizeof(aff)); continue; }
    fp zi, zi2, zi3;
    fp_mul(&zi, &inv, &pref[i]);
    fp_mul(&inv, &inv, &in[i].z);
    fp_sqr(&zi2, &zi);
    fp_mul(&zi3, &zi2, &zi);
    fp_mul(&out[i].x, &in[i].x, &zi2);
    fp_mul(&out[i].y, &in[i].y, &zi3);
  }
  free(pref);
}
static int scalar_bit(const uint64_t* k, int i) { return (int)((k[i >> 6] >> (i & 63)) & 1); }
static unsigned scalar_window(const uint64_t* k, int lo, int c) {
  unsigned v = 0;
  for (int b = 0; b < c; b++) {
    int i = lo + b;
    if (i < 256) v |= (unsigned)scalar_bit(k, i) << b;
  }
  return v;
}
/* r = k*p, 4-bit unsigned fixed windows; nbits: scalar length to scan */
static void jac_mul(jac* r, const aff* p, const uint64_t* k, int nbits) {
  jac tab[16];
  jac_set_inf(&tab[0]);
  jac_from_aff(&tab[1], p);
  for (int i = 2; i < 16; i++) jac_add_aff(&tab[i], &tab[i - 1], p);
  jac acc;
  jac_set_inf(&acc);
  int nw = (nbits + 3) / 4;
  for (int w = nw - 1; w >= 0; w--) {
    for (int d = 0; d < 4; d++) jac_dbl(&acc, &acc);
    unsigned v = scalar_window(k, 4 * w, 4);
    if (v) jac_add(&acc, &acc, &tab[v]);
  }
  *r = acc;
}

/* ---------------------------------------------------------------- I/O helpers */
static void load_fp_le(fp* r, const uint8_t* b) { /* canonical LE bytes -> Montgomery */
  fp c;
  memcpy(c.l, b, 48);
  fp_to_mont(r, &c);
}
static void store_fp_le(uint8_t* b, const fp* a) {
  fp c;
  fp_from_mont(&c, a);
  memcpy(b, c.l, 48);
}
static void load_aff(aff* r, const uint8_t* b) { load_fp_le(&r->x, b); load_fp_le(&r->y, b + 48); }
static void store_aff(uint8_t* b, const aff* a) { store_fp_le(b, &a->x); store_fp_le(b + 48, &a->y); }

/* ---------------------------------------------------------------- exported */
void co_g1_mul_batch(const uint8_t* points, const uint8_t* scalars, size_t n, size_t scalar_stride, uint8_t* out) {
  jac* res = (jac*)malloc((n ? n : 1) * sizeof(jac));
  aff* ao = (aff*)malloc((n ? n : 1) * sizeof(aff));
  for (size_t i = 0; i < n; i++) {
    aff p;
    uint64_t k[4];
    load_aff(&p, points + 96 * i);
    memcpy(k, scalars + 32 * i * scalar_stride, 32);
    jac_mul(&res[i], &p, k, 256);
  }
  jac_batch_to_aff(ao, res, n);
  for (size_t i = 0; i < n; i++) store_aff(out + 96 * i, &ao[i]);
  free(res);
  free(ao);
}

void co_g1_fold(const uint8_t* L, const uint8_t* R, const uint8_t* x, size_t n, uint8_t* out) {
  jac* res = (jac*)malloc((n ? n : 1) * sizeof(jac));
  aff* ao = (aff*)malloc((n ? n : 1) * sizeof(aff));
  uint64_t k[4];
  memcpy(k, x, 32);
  for (size_t i = 0; i < n; i++) {
    aff l, r;
    load_aff(&l, L + 96 * i);
    load_aff(&r, R + 96 * i);
    jac t;
    jac_mul(&t, &r, k, 256);
    jac_add_aff(&res[i], &t, &l);
  }
  jac_batch_to_aff(ao, res, n);
  for (size_t i = 0; i < n; i++) store_aff(out + 96 * i, &ao[i]);
  free(res);
  free(ao);
}

typedef struct {
  const aff* pts;
  const uint64_t* sc; /* 4 words per scalar */
  size_t n;
  int c, w_begin, w_end;
  jac* win; /* per-window sums */
} msm_job;

static void* msm_worker(void* arg) {
  msm_job* j = (msm_job*)arg;
  int nb = 1 << j->c;
  jac* buckets = (jac*)malloc((size_t)nb * sizeof(jac));
  for (int w = j->w_begin; w < j->w_end; w++) {
    for (int b = 0; b < nb; b++) jac_set_inf(&buckets[b]);
    for (size_t i = 0; i < j->n; i++) {
      unsigned v = scalar_window(j->sc + 4 * i, w * j->c, j->c);
      if (v) jac_add_aff(&buckets[v], &buckets[v], &j->pts[i]);
    }
    jac run, acc;
    jac_set_inf(&run);
    jac_set_inf(&acc);
    for (int b = nb - 1; b >= 1; b--) {
      jac_add(&run, &run, &buckets[b]);
      jac_add(&acc, &acc, &run);
    }
    j->win[w] = acc;
  }
  free(buckets);
  return NULL;
}

/* out = sum scalars[i]*points[i]; bucket method, windows spread over nthreads */
void co_g1_msm(const uint8_t* points, const uint8_t* scalars, size_t n, uint8_t* out, int nthreads) {
  aff res;
  memset(&res, 0, sizeof res);
  if (n == 0) { store_aff(out, &res); memset(out, 0, 96); return; }
  aff* pts = (aff*)malloc(n * sizeof(aff));
  uint64_t* sc = (uint64_t*)malloc(n * 32);
  for (size_t i = 0; i < n; i++) load_aff(&pts[i], points + 96 * i);
  memcpy(sc, scalars, n * 32);
  int c = 2;
  if (n >= 8) c = 3;
  if (n >= 32) c = 5;
  if (n >= 128) c = 6;
  if (n >= 512) c = 8;
  if (n >= 4096) c = 10;
  if (n >= 32768) c = 12;
  if (n >= 262144) c = 14;
  if (n >= 1048576) c = 16;
  int nw = (255 + c - 1) / c;
  jac* win = (jac*)malloc((size_t)nw * sizeof(jac));
  if (nthreads < 1) nthreads = 1;
  if (nthreads > nw) nthreads = nw;
  if (n < 256) nthreads = 1;
  msm_job* jobs = (msm_job*)malloc((size_t)nthreads * sizeof(msm_job));
  pthread_t* th = (pthread_t*)malloc((size_t)nthreads * sizeof(pthread_t));
  for (int t = 0; t < nthreads; t++) {
    jobs[t].pts = pts; jobs[t].sc = sc; jobs[t].n = n; jobs[t].c = c; jobs[t].win = win;
    jobs[t].w_begin = (int)((long)nw * t / nthreads);
    jobs[t].w_end = (int)((long)nw * (t + 1) / nthreads);
  }
  if (nthreads == 1) msm_worker(&jobs[0]);
  else {
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, msm_worker, &jobs[t]);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  }
  jac total;
  jac_set_inf(&total);
  for (int w = nw - 1; w >= 0; w--) {
    for (int d = 0; d < c; d++) jac_dbl(&total, &total);
    jac_add(&total, &total, &win[w]);
  }
  jac_batch_to_aff(&res, &total, 1);
  store_aff(out, &res);
  if (aff_is_inf(&res)) memset(out, 0, 96);
  free(pts); free(sc); free(win); free(jobs); free(th);
}

/* sum of points */
void co_g1_sum(const uint8_t* points, size_t n, uint8_t* out) {
  jac acc;
  jac_set_inf(&acc);
  for (size_t i = 0; i < n; i++) {
    aff p;
    load_aff(&p, points + 96 * i);
    jac_add_aff(&acc, &acc, &p);
  }
  aff r;
  jac_batch_to_aff(&r, &acc, 1);
  store_aff(out, &r);
  if (aff_is_inf(&r)) memset(out, 0, 96);
}

static int lex_largest(const fp* y_mont) { /* canonical y > (p-1)/2 */
  fp c;
  fp_from_mont(&c, y_mont);
  /* 2y > p-1  <=> 2y >= p (p odd) */
  uint64_t carry = 0;
  fp d;
  for (int i = 0; i < 6; i++) { d.l[i] = (c.l[i] << 1) | carry; carry = c.l[i] >> 63; }
  return carry || fp_geq_p(&d);
}

void co_g1_compress(const uint8_t* points, size_t n, uint8_t* out48) {
  for (size_t i = 0; i < n; i++) {
    const uint8_t* p = points + 96 * i;
    uint8_t* o = out48 + 48 * i;
    int inf = 1;
    for (int k = 0; k < 96; k++) if (p[k]) { inf = 0; break; }
    if (inf) { memset(o, 0, 48); o[0] = 0xc0; continue; }
    for (int k = 0; k < 48; k++) o[k] = p[47 - k];
    fp y;
    load_fp_le(&y, p + 48);
    o[0] |= lex_largest(&y) ? 0xa0 : 0x80;
  }
}

static int in_subgroup(const aff* p) {
  jac t;
  jac_mul(&t, p, FR_R, 256);
  return jac_is_inf(&t);
}

/* status: 0 ok, 1 flags, 2 x >= p, 3 no sqrt, 4 subgroup, 5 infinity padding */
void co_g1_decompress(const uint8_t* in48, size_t n, uint8_t* out, uint8_t* status) {
  for (size_t i = 0; i < n; i++) {
    const uint8_t* b = in48 + 48 * i;
    uint8_t* o = out + 96 * i;
    memset(o, 0, 96);
    unsigned flags = b[0] & 0xe0;
    if (!(flags & 0x80) || flags == 0xe0) { status[i] = 1; continue; }
    if (flags & 0x40) {
      unsigned acc = b[0] & 0x1f;
      for (int k = 1; k < 48; k++) acc |= b[k];
      status[i] = acc ? 5 : 0;
      continue;
    }
    uint8_t le[48];
    for (int k = 0; k < 48; k++) le[k] = b[47 - k];
    le[47] &= 0x1f;
    fp xc;
    memcpy(xc.l, le, 48);
    if (fp_geq_p(&xc)) { status[i] = 2; continue; }
    aff p;
    fp_to_mont(&p.x, &xc);
    fp y2, four;
    fp_sqr(&y2, &p.x);
    fp_mul(&y2, &y2, &p.x);
    fp_add(&four, &FP_ONE, &FP_ONE);
    fp_add(&four, &four, &four);
    fp_add(&y2, &y2, &four);
    if (!fp_sqrt(&p.y, &y2)) { status[i] = 3; continue; }
    if (lex_largest(&p.y) != ((flags & 0x20) != 0)) fp_neg(&p.y, &p.y);
    if (!in_subgroup(&p)) { status[i] = 4; continue; }
    store_aff(o, &p);
    status[i] = 0;
  }
}

/* ---------------------------------------------------------------- Keccak */
static const uint64_t KRC[24] = {
    0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull,
    0x000000000000808bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull,
    0x000000000000008aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000aull,
    0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull,
    0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
    0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
static const int KROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
static uint64_t rol(uint64_t v, int n) { return n ? (v << n) | (v >> (64 - n)) : v; }

void co_keccak_f1600(uint8_t* state) {
  uint64_t a[25];
  memcpy(a, state, 200); /* little-endian host */
  for (int rnd = 0; rnd < 24; rnd++) {
    uint64_t c[5], d[5], b[25];
    for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
    for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rol(c[(x + 1) % 5], 1);
    for (int i = 0; i < 25; i++) a[i] ^= d[i % 5];
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rol(a[x + 5 * y], KROT[x + 5 * y]);
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
    a[0] ^= KRC[rnd];
  }
  memcpy(state, a, 200);
}
