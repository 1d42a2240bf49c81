/* CPU restatement ("port") of the reference's hot path in plain C — TEST INFRASTRUCTURE ONLY.
 *
 * Nothing in gnark_whir_b200/ links or loads this file.  It exists (a) to check the CUDA path at
 * sizes the python big-int oracle cannot reach, and (b) as the CPU baseline bench.py times on
 * the GPU box's host cores (cpu_baseline.kind = "port"; gnark itself cannot be built here: no Go
 * toolchain, modules not vendored).  PARITY UNPINNED by the reference (it ships no tests); this
 * file is pinned against the python modules in oracle/, which is pinned by first-principles known answers.
 *
 * What it follows (gnark-crypto v0.14.1-0.20241217131346-b998989abdbe, gnark v0.11.0; go.mod:6-7;
 * reached from /root/reference/mt.go:448,496 and keccakSponge/keccakSponge.go:48,69):
 *   fe_mul_*            ecc/bn254/fp, fr  element.Mul   : 4x64 CIOS Montgomery, R = 2^256
 *   g1_* (XYZZ)         ecc/bn254/g1.go g1JacExtended   : add / mixed add / double
 *   oracle_msm_g1       ecc/bn254/multiexp.go MultiExp  : bestC window rule, signed digits
 *                       (partitionScalars), per-window bucket accumulation with extended
 *                       Jacobian buckets, running-sum reduction, Horner over windows; windows
 *                       (and point ranges, when there are more threads than windows) in parallel
 *   oracle_ntt          ecc/bn254/fr/fft fft.go         : difFFT / ditFFT, coset, 1/N
 *   oracle_compute_h    backend/groth16/bn254/prove.go  : computeH
 *   oracle_keccak_f / oracle_sponge / oracle_merkle_paths
 *                       std/permutation/keccakf + keccakSponge.go:17-75 + mtUtilities.go:109-141
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef uint64_t u64;
typedef struct { u64 l[4]; } fe;

static const u64 PM[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
static const u64 RM[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
#define P_INV 0x87d20782e4866389ull
#define R_INV 0xc2e1f593efffffffull
static const fe FP_ONE = {{0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full}};
static const fe FR_ONE = {{0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full}};
static const fe FR_R2 = {{0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull}};
static const fe FR_ROOT28 = {{0x636e735580d13d9cull, 0xa22bf3742445ffd6ull, 0x56452ac01eb203d8ull, 0x1860ef942963f9e7ull}};
static const fe FR_GEN = {{0x1b0d0ef99fffffe6ull, 0xeaba68a3a32a913full, 0x47d8eb76d8dd0689ull, 0x15d0085520f5bbc3ull}};

static inline void fe_mul_mod(fe* r, const fe* a, const fe* b, const u64* M, u64 inv) {
  u64 t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u64 carry = 0;
    for (int j = 0; j < 4; j++) {
      u128 p = (u128)a->l[j] * b->l[i] + t[j] + carry;
      t[j] = (u64)p; carry = (u64)(p >> 64);
    }
    u128 q = (u128)t[4] + carry;
    t[4] = (u64)q; t[5] = (u64)(q >> 64);
    u64 m = t[0] * inv;
    u128 p = (u128)m * M[0] + t[0];
    carry = (u64)(p >> 64);
    for (int j = 1; j < 4; j++) {
      p = (u128)m * M[j] + t[j] + carry;
      t[j - 1] = (u64)p; carry = (u64)(p >> 64);
    }
    q = (u128)t[4] + carry;
    t[3] = (u64)q; t[4] = t[5] + (u64)(q >> 64);
  }
  u64 d[4], borrow = 0;
  for (int i = 0; i < 4; i++) {
    u128 x = (u128)t[i] - M[i] - borrow;
    d[i] = (u64)x; borrow = (u64)(x >> 64) & 1;
  }
  int ge = t[4] != 0 || borrow == 0;
  for (int i = 0; i < 4; i++) r->l[i] = ge ? d[i] : t[i];
}
static inline void fe_add_mod(fe* r, const fe* a, const fe* b, const u64* M) {
  u64 t[4], carry = 0;
  for (int i = 0; i < 4; i++) { u128 x = (u128)a->l[i] + b->l[i] + carry; t[i] = (u64)x; carry = (u64)(x >> 64); }
  u64 d[4], borrow = 0;
  for (int i = 0; i < 4; i++) { u128 x = (u128)t[i] - M[i] - borrow; d[i] = (u64)x; borrow = (u64)(x >> 64) & 1; }
  int ge = carry || !borrow;
  for (int i = 0; i < 4; i++) r->l[i] = ge ? d[i] : t[i];
}
static inline void fe_sub_mod(fe* r, const fe* a, const fe* b, const u64* M) {
  u64 t[4], borrow = 0;
  for (int i = 0; i < 4; i++) { u128 x = (u128)a->l[i] - b->l[i] - borrow; t[i] = (u64)x; borrow = (u64)(x >> 64) & 1; }
  if (borrow) {
    u64 carry = 0;
    for (int i = 0; i < 4; i++) { u128 x = (u128)t[i] + M[i] + carry; t[i] = (u64)x; carry = (u64)(x >> 64); }
  }
  for (int i = 0; i < 4; i++) r->l[i] = t[i];
}
static inline int fe_is_zero(const fe* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int fe_eq(const fe* a, const fe* b) {
  return ((a->l[0] ^ b->l[0]) | (a->l[1] ^ b->l[1]) | (a->l[2] ^ b->l[2]) | (a->l[3] ^ b->l[3])) == 0;
}
#define fp_mul(r, a, b) fe_mul_mod(r, a, b, PM, P_INV)
#define fp_add(r, a, b) fe_add_mod(r, a, b, PM)
#define fp_sub(r, a, b) fe_sub_mod(r, a, b, PM)
#define fr_mul(r, a, b) fe_mul_mod(r, a, b, RM, R_INV)
#define fr_add(r, a, b) fe_add_mod(r, a, b, RM)
#define fr_sub(r, a, b) fe_sub_mod(r, a, b, RM)

static void fe_pow_mod(fe* r, const fe* a, const u64* e, const fe* one, const u64* M, u64 inv) {
  fe acc = *one;
  for (int i = 255; i >= 0; i--) {
    fe_mul_mod(&acc, &acc, &acc, M, inv);
    if ((e[i >> 6] >> (i & 63)) & 1) fe_mul_mod(&acc, &acc, a, M, inv);
  }
  *r = acc;
}
static void fp_inv(fe* r, const fe* a) { u64 e[4] = {PM[0] - 2, PM[1], PM[2], PM[3]}; fe_pow_mod(r, a, e, &FP_ONE, PM, P_INV); }
static void fr_inv(fe* r, const fe* a) { u64 e[4] = {RM[0] - 2, RM[1], RM[2], RM[3]}; fe_pow_mod(r, a, e, &FR_ONE, RM, R_INV); }
static void fr_from_mont(fe* r, const fe* a) { fe one = {{1, 0, 0, 0}}; fr_mul(r, a, &one); }
static void fr_from_u64(fe* r, u64 v) { fe t = {{v, 0, 0, 0}}; fr_mul(r, &t, &FR_R2); }

/* ---------------------------------------------------------------- G1, extended Jacobian */
typedef struct { fe x, y; } g1a;              /* affine; infinity = (0,0) */
typedef struct { fe x, y, zz, zzz; } g1x;     /* x = X/ZZ, y = Y/ZZZ; infinity: zz = 0 */

static inline int g1a_is_inf(const g1a* p) { return fe_is_zero(&p->x) && fe_is_zero(&p->y); }
static inline void g1x_set_inf(g1x* p) { memset(p, 0, sizeof(*p)); }

static void g1x_dbl_affine(g1x* r, const g1a* p) {
  fe U, V, W, S, X2, M, t;
  fp_add(&U, &p->y, &p->y);
  fp_mul(&V, &U, &U);
  fp_mul(&W, &U, &V);
  fp_mul(&S, &p->x, &V);
  fp_mul(&X2, &p->x, &p->x);
  fp_add(&M, &X2, &X2); fp_add(&M, &M, &X2);
  fp_mul(&r->x, &M, &M); fp_sub(&r->x, &r->x, &S); fp_sub(&r->x, &r->x, &S);
  fp_sub(&t, &S, &r->x); fp_mul(&t, &M, &t);
  fp_mul(&r->y, &W, &p->y); fp_sub(&r->y, &t, &r->y);
  r->zz = V; r->zzz = W;
}
static void g1x_dbl(g1x* p) {
  if (fe_is_zero(&p->zz)) return;
  fe U, V, W, S, X2, M, t, X3, Y3;
  fp_add(&U, &p->y, &p->y);
  fp_mul(&V, &U, &U);
  fp_mul(&W, &U, &V);
  fp_mul(&S, &p->x, &V);
  fp_mul(&X2, &p->x, &p->x);
  fp_add(&M, &X2, &X2); fp_add(&M, &M, &X2);
  fp_mul(&X3, &M, &M); fp_sub(&X3, &X3, &S); fp_sub(&X3, &X3, &S);
  fp_sub(&t, &S, &X3); fp_mul(&t, &M, &t);
  fp_mul(&Y3, &W, &p->y); fp_sub(&Y3, &t, &Y3);
  p->x = X3; p->y = Y3;
  fp_mul(&p->zz, &V, &p->zz);
  fp_mul(&p->zzz, &W, &p->zzz);
}
/* acc += (neg ? -p : p) */
static void g1x_madd(g1x* acc, const g1a* p, int neg) {
  if (g1a_is_inf(p)) return;
  fe py = p->y;
  if (neg && !fe_is_zero(&py)) { fe z = {{0, 0, 0, 0}}; fp_sub(&py, &z, &py); }
  if (fe_is_zero(&acc->zz)) { acc->x = p->x; acc->y = py; acc->zz = FP_ONE; acc->zzz = FP_ONE; return; }
  fe Pq, Rq, PP, PPP, Q, X3, t;
  fp_mul(&Pq, &p->x, &acc->zz); fp_sub(&Pq, &Pq, &acc->x);
  fp_mul(&Rq, &py, &acc->zzz); fp_sub(&Rq, &Rq, &acc->y);
  if (fe_is_zero(&Pq)) {
    if (fe_is_zero(&Rq)) { g1a q = {p->x, py}; g1x_dbl_affine(acc, &q); }
    else g1x_set_inf(acc);
    return;
  }
  fp_mul(&PP, &Pq, &Pq);
  fp_mul(&PPP, &Pq, &PP);
  fp_mul(&Q, &acc->x, &PP);
  fp_mul(&X3, &Rq, &Rq); fp_sub(&X3, &X3, &PPP); fp_sub(&X3, &X3, &Q); fp_sub(&X3, &X3, &Q);
  fp_sub(&t, &Q, &X3); fp_mul(&t, &Rq, &t);
  fp_mul(&acc->y, &acc->y, &PPP); fp_sub(&acc->y, &t, &acc->y);
  acc->x = X3;
  fp_mul(&acc->zz, &acc->zz, &PP);
  fp_mul(&acc->zzz, &acc->zzz, &PPP);
}
static void g1x_add(g1x* acc, const g1x* q) {
  if (fe_is_zero(&q->zz)) return;
  if (fe_is_zero(&acc->zz)) { *acc = *q; return; }
  fe U1, U2, S1, S2, Pq, Rq, PP, PPP, Q, X3, t;
  fp_mul(&U1, &acc->x, &q->zz);
  fp_mul(&U2, &q->x, &acc->zz);
  fp_mul(&S1, &acc->y, &q->zzz);
  fp_mul(&S2, &q->y, &acc->zzz);
  fp_sub(&Pq, &U2, &U1);
  fp_sub(&Rq, &S2, &S1);
  if (fe_is_zero(&Pq)) {
    if (fe_is_zero(&Rq)) g1x_dbl(acc); else g1x_set_inf(acc);
    return;
  }
  fp_mul(&PP, &Pq, &Pq);
  fp_mul(&PPP, &Pq, &PP);
  fp_mul(&Q, &U1, &PP);
  fp_mul(&X3, &Rq, &Rq); fp_sub(&X3, &X3, &PPP); fp_sub(&X3, &X3, &Q); fp_sub(&X3, &X3, &Q);
  fp_sub(&t, &Q, &X3); fp_mul(&t, &Rq, &t);
  fp_mul(&S1, &S1, &PPP); fp_sub(&acc->y, &t, &S1);
  acc->x = X3;
  fp_mul(&acc->zz, &acc->zz, &q->zz); fp_mul(&acc->zz, &acc->zz, &PP);
  fp_mul(&acc->zzz, &acc->zzz, &q->zzz); fp_mul(&acc->zzz, &acc->zzz, &PPP);
}
static void g1x_to_affine(g1a* r, const g1x* p) {
  if (fe_is_zero(&p->zz)) { memset(r, 0, sizeof(*r)); return; }
  fe zi, a, izz;
  fp_inv(&zi, &p->zzz);
  fp_mul(&a, &p->zz, &zi);
  fp_mul(&izz, &a, &a);
  fp_mul(&r->x, &p->x, &izz);
  fp_mul(&r->y, &p->y, &zi);
}

/* ---------------------------------------------------------------- MSM (gnark MultiExp) */
static int best_c(size_t n) {           /* multiexp.go bestC: argmin over implemented c of (bits+1)(n+2^c)/c */
  int best = 4; double bc = 1e300;
  for (int c = 4; c <= 16; c++) {
    double cost = 255.0 * ((double)n + (double)(1u << c)) / c;
    if (cost < bc) { bc = cost; best = c; }
  }
  return best;
}

int oracle_msm_window(size_t n) { return best_c(n); }

/* out (affine, Montgomery) = sum scalars[i] * points[i];  scalars are fr Montgomery. */
int oracle_msm_g1(const u64* points, const u64* scalars, size_t n, int threads, u64* out) {
  const g1a* pts = (const g1a*)points;
  const fe* sc = (const fe*)scalars;
  if (n == 0) { memset(out, 0, 64); return 0; }
  int c = best_c(n);
  int W = (254 + c - 1) / c;
  if (c * W < 256) { /* the signed recoding may carry out of the top window */
    u64 topbits = 254 - (u64)(W - 1) * c;
    if (topbits >= (u64)(c - 1)) W += 1;
  }
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  threads = 1;
#endif
  /* partitionScalars: canonical form, signed digits in (-2^(c-1), 2^(c-1)] */
  int32_t* digits = (int32_t*)malloc((size_t)W * n * sizeof(int32_t));
  if (!digits) return -1;
#pragma omp parallel for num_threads(threads) schedule(static)
  for (size_t i = 0; i < n; i++) {
    fe s; fr_from_mont(&s, &sc[i]);
    u64 l[5] = {s.l[0], s.l[1], s.l[2], s.l[3], 0};
    int carry = 0;
    for (int w = 0; w < W; w++) {
      int bit = w * c, q = bit >> 6, r = bit & 63;
      u64 raw = 0;
      if (q < 4) { raw = l[q] >> r; if (r + c > 64) raw |= l[q + 1] << (64 - r); raw &= ((1ull << c) - 1); }
      int64_t d = (int64_t)raw + carry;
      if (d > (1ll << (c - 1))) { d -= (1ll << c); carry = 1; } else carry = 0;
      digits[(size_t)w * n + i] = (int32_t)d;
    }
  }
  /* point-range splits per window: gnark's MultiExp splits the points when it has more workers than
   * windows ("nbSplits"); W windows rarely divide the thread count evenly, so aim for >= 2 tasks per
   * thread and let the dynamic schedule balance them */
  int S = (2 * threads + W - 1) / W;
  if (S < 1) S = 1;
  if ((size_t)S > n / 1024 + 1) S = (int)(n / 1024 + 1);
  size_t nb = (size_t)1 << (c - 1);
  g1x* sums = (g1x*)calloc((size_t)W * S, sizeof(g1x));
  int fail = 0;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
  for (int task = 0; task < W * S; task++) {
    int w = task / S, sp = task % S;
    size_t lo = n * (size_t)sp / S, hi = n * (size_t)(sp + 1) / S;
    g1x* buckets = (g1x*)calloc(nb, sizeof(g1x));
    if (!buckets) { fail = 1; continue; }
    const int32_t* dg = digits + (size_t)w * n;
    for (size_t i = lo; i < hi; i++) {
      int32_t d = dg[i];
      if (d > 0) g1x_madd(&buckets[d - 1], &pts[i], 0);
      else if (d < 0) g1x_madd(&buckets[-d - 1], &pts[i], 1);
    }
    g1x run, acc; g1x_set_inf(&run); g1x_set_inf(&acc);
    for (size_t b = nb; b-- > 0;) { g1x_add(&run, &buckets[b]); g1x_add(&acc, &run); }
    sums[task] = acc;
    free(buckets);
  }
  g1x total; g1x_set_inf(&total);
  for (int w = W - 1; w >= 0; w--) {
    for (int k = 0; k < c; k++) g1x_dbl(&total);
    for (int sp = 0; sp < S; sp++) g1x_add(&total, &sums[w * S + sp]);
  }
  g1a res; g1x_to_affine(&res, &total);
  memcpy(out, &res, 64);
  free(sums); free(digits);
  return fail ? -1 : 0;
}

/* points[i] = (k0 + i*d) * G for the generator (1,2): an arithmetic progression of known
 * discrete logs (cheap to make: one mixed add each + a batched normalisation). */
static void g1_progression_range(const g1a* G, const fe* k0, const fe* dd, const g1a* Da, size_t lo, size_t hi,
                                 g1a* o, g1x* jac, fe* pref) {
  /* start = (k0 + lo*d) * G by double-and-add, then one mixed add per point and a batched normalisation */
  fe start = *k0, lo_fe, t; fr_from_u64(&lo_fe, (u64)lo); fr_mul(&t, &lo_fe, dd); fr_add(&start, &start, &t);
  fe sc; fr_from_mont(&sc, &start);
  g1x cur; g1x_set_inf(&cur);
  for (int i = 255; i >= 0; i--) { g1x_dbl(&cur); if ((sc.l[i >> 6] >> (i & 63)) & 1) g1x_madd(&cur, G, 0); }
  for (size_t i = lo; i < hi; i++) { jac[i] = cur; g1x_madd(&cur, Da, 0); }
  fe run = FP_ONE;
  for (size_t i = lo; i < hi; i++) { pref[i] = run; if (!fe_is_zero(&jac[i].zz)) fp_mul(&run, &run, &jac[i].zzz); }
  fe inv; fp_inv(&inv, &run);
  for (size_t i = hi; i-- > lo;) {
    if (fe_is_zero(&jac[i].zz)) { memset(&o[i], 0, sizeof(g1a)); continue; }
    fe zi, a, izz; fp_mul(&zi, &inv, &pref[i]); fp_mul(&inv, &inv, &jac[i].zzz);
    fp_mul(&a, &jac[i].zz, &zi); fp_mul(&izz, &a, &a);
    fp_mul(&o[i].x, &jac[i].x, &izz); fp_mul(&o[i].y, &jac[i].y, &zi);
  }
}

int oracle_g1_progression(const u64* k0_mont, const u64* d_mont, size_t n, u64* out_points) {
  g1a G; G.x = FP_ONE; fp_add(&G.y, &FP_ONE, &FP_ONE);
  const fe* k0 = (const fe*)k0_mont; const fe* dm = (const fe*)d_mont;
  fe dd; fr_from_mont(&dd, dm);
  g1x D; g1x_set_inf(&D);
  for (int i = 255; i >= 0; i--) { g1x_dbl(&D); if ((dd.l[i >> 6] >> (i & 63)) & 1) g1x_madd(&D, &G, 0); }
  g1a Da; g1x_to_affine(&Da, &D);
  g1x* jac = (g1x*)malloc((n ? n : 1) * sizeof(g1x));
  fe* pref = (fe*)malloc((n ? n : 1) * sizeof(fe));
  if (!jac || !pref) return -1;
  size_t chunk = 1 << 14, nchunks = (n + chunk - 1) / chunk;
#pragma omp parallel for schedule(dynamic, 1)
  for (size_t q = 0; q < nchunks; q++) {
    size_t lo = q * chunk, hi = lo + chunk < n ? lo + chunk : n;
    g1_progression_range(&G, k0, dm, &Da, lo, hi, (g1a*)out_points, jac, pref);
  }
  free(jac); free(pref);
  return 0;
}

/* out = k * G (affine Montgomery), k fr Montgomery: closed-form check value */
void oracle_g1_gen_mul(const u64* k_mont, u64* out) {
  g1a G; G.x = FP_ONE; fp_add(&G.y, &FP_ONE, &FP_ONE);
  fe k; fr_from_mont(&k, (const fe*)k_mont);
  g1x acc; g1x_set_inf(&acc);
  for (int i = 255; i >= 0; i--) { g1x_dbl(&acc); if ((k.l[i >> 6] >> (i & 63)) & 1) g1x_madd(&acc, &G, 0); }
  g1a r; g1x_to_affine(&r, &acc); memcpy(out, &r, 64);
}

/* out = sum a[i]*b[i] in Fr (Montgomery in, Montgomery out) */
void oracle_fr_dot(const u64* a, const u64* b, size_t n, int threads, u64* out) {
  const fe* A = (const fe*)a; const fe* B = (const fe*)b;
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  threads = 1;
#endif
  fe* part = (fe*)calloc((size_t)threads, sizeof(fe));
#pragma omp parallel num_threads(threads)
  {
#ifdef _OPENMP
    int t = omp_get_thread_num(), nt = omp_get_num_threads();
#else
    int t = 0, nt = 1;
#endif
    fe acc = {{0, 0, 0, 0}};
    for (size_t i = (size_t)t; i < n; i += (size_t)nt) { fe p; fr_mul(&p, &A[i], &B[i]); fr_add(&acc, &acc, &p); }
    part[t] = acc;
  }
  fe tot = {{0, 0, 0, 0}};
  for (int t = 0; t < threads; t++) fr_add(&tot, &tot, &part[t]);
  memcpy(out, &tot, 32); free(part);
}

/* sum_i a[i]*(k0 + i*d) in Fr: dot product against the progression's discrete logs */
void oracle_fr_dot_progression(const u64* a, const u64* k0_mont, const u64* d_mont, size_t n, u64* out) {
  const fe* A = (const fe*)a; fe k = *(const fe*)k0_mont, d = *(const fe*)d_mont, tot = {{0, 0, 0, 0}};
  for (size_t i = 0; i < n; i++) { fe p; fr_mul(&p, &A[i], &k); fr_add(&tot, &tot, &p); fr_add(&k, &k, &d); }
  memcpy(out, &tot, 32);
}

/* ---------------------------------------------------------------- NTT (gnark fft.Domain) */
static u64 bitrev64(u64 i, int L) { u64 r = 0; for (int k = 0; k < L; k++) { r = (r << 1) | (i & 1); i >>= 1; } return r; }

static void domain_gen(int L, int inverse, fe* w) {
  fe g = FR_ROOT28;
  for (int i = L; i < 28; i++) fr_mul(&g, &g, &g);
  if (inverse) fr_inv(&g, &g);
  *w = g;
}

/* in-place; decimation 0 = DIF (natural in, bit-reversed out), 1 = DIT (bit-reversed in, natural out) */
int oracle_ntt(u64* data, int L, int inverse, int coset, int decimation, int threads) {
  fe* a = (fe*)data; size_t n = (size_t)1 << L;
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  threads = 1;
#endif
  fe w; domain_gen(L, inverse, &w);
  fe* tw = (fe*)malloc((n / 2 + 1) * sizeof(fe));        /* w^e, e < n/2 */
  if (!tw) return -1;
  tw[0] = FR_ONE;
  for (size_t e = 1; e < n / 2; e++) fr_mul(&tw[e], &tw[e - 1], &w);
  fe* cos = NULL;
  if (coset) {
    cos = (fe*)malloc(n * sizeof(fe));
    if (!cos) { free(tw); return -1; }
    fe g = FR_GEN, first = FR_ONE;
    if (inverse) { fr_inv(&g, &g); fe nn; fr_from_u64(&nn, (u64)n); fr_inv(&first, &nn); }
    cos[0] = first;
    for (size_t i = 1; i < n; i++) fr_mul(&cos[i], &cos[i - 1], &g);
  }
  if (coset && !inverse) {
#pragma omp parallel for num_threads(threads) schedule(static)
    for (size_t i = 0; i < n; i++) { size_t k = decimation ? bitrev64(i, L) : i; fr_mul(&a[i], &a[i], &cos[k]); }
  }
  if (decimation == 0) {
    for (size_t m = n / 2, stride = 1; m >= 1; m >>= 1, stride <<= 1) {
#pragma omp parallel for num_threads(threads) schedule(static)
      for (size_t q = 0; q < n / 2; q++) {
        size_t j = q % m, s = (q / m) * 2 * m;
        fe x = a[s + j], y = a[s + j + m], d;
        fr_add(&a[s + j], &x, &y);
        fr_sub(&d, &x, &y);
        fr_mul(&a[s + j + m], &d, &tw[j * stride]);
      }
    }
  } else {
    for (size_t m = 1, stride = n / 2; m < n; m <<= 1, stride >>= 1) {
#pragma omp parallel for num_threads(threads) schedule(static)
      for (size_t q = 0; q < n / 2; q++) {
        size_t j = q % m, s = (q / m) * 2 * m;
        fe x = a[s + j], y;
        fr_mul(&y, &a[s + j + m], &tw[j * stride]);
        fr_add(&a[s + j], &x, &y);
        fr_sub(&a[s + j + m], &x, &y);
      }
    }
  }
  if (inverse) {
    if (coset) {
#pragma omp parallel for num_threads(threads) schedule(static)
      for (size_t i = 0; i < n; i++) { size_t k = decimation ? i : bitrev64(i, L); fr_mul(&a[i], &a[i], &cos[k]); }
    } else {
      fe nn, ni; fr_from_u64(&nn, (u64)n); fr_inv(&ni, &nn);
#pragma omp parallel for num_threads(threads) schedule(static)
      for (size_t i = 0; i < n; i++) fr_mul(&a[i], &a[i], &ni);
    }
  }
  free(tw); if (cos) free(cos);
  return 0;
}

/* a, b, c: N = 2^L elements each (already zero padded); h overwrites a (bit-reversed order) */
int oracle_compute_h(u64* a, u64* b, u64* c, int L, int threads) {
  size_t n = (size_t)1 << L;
  u64* v[3] = {a, b, c};
  for (int i = 0; i < 3; i++) if (oracle_ntt(v[i], L, 1, 0, 0, threads)) return -1;
  for (int i = 0; i < 3; i++) if (oracle_ntt(v[i], L, 0, 1, 1, threads)) return -1;
  fe gn = FR_GEN;
  for (int i = 0; i < L; i++) fr_mul(&gn, &gn, &gn);
  fe den; fr_sub(&den, &gn, &FR_ONE); fr_inv(&den, &den);
  fe *A = (fe*)a, *B = (fe*)b, *C = (fe*)c;
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#endif
#pragma omp parallel for num_threads(threads) schedule(static)
  for (size_t i = 0; i < n; i++) { fe t; fr_mul(&t, &A[i], &B[i]); fr_sub(&t, &t, &C[i]); fr_mul(&A[i], &t, &den); }
  return oracle_ntt(a, L, 1, 1, 0, threads);
}

/* ---------------------------------------------------------------- Keccak */
static const u64 KRC[24] = {
    0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull,
    0x000000000000808bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull,
    0x000000000000008aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000aull,
    0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull,
    0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
    0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
static const int KROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};

void oracle_keccak_f(u64* a) {
  for (int rnd = 0; rnd < 24; rnd++) {
    u64 c[5], d[5], b[25];
    for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
    for (int x = 0; x < 5; x++) { u64 t = c[(x + 1) % 5]; d[x] = c[(x + 4) % 5] ^ ((t << 1) | (t >> 63)); }
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) {
        u64 v = a[x + 5 * y] ^ d[x]; int r = KROT[x + 5 * y];
        b[y + 5 * ((2 * x + 3 * y) % 5)] = r ? ((v << r) | (v >> (64 - r))) : v;
      }
    for (int y = 0; y < 5; y++)
      for (int x = 0; x < 5; x++) a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
    a[0] ^= KRC[rnd];
  }
}

void oracle_keccak_f_batch(u64* states, size_t n, int threads) {
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#endif
#pragma omp parallel for num_threads(threads) schedule(static)
  for (size_t i = 0; i < n; i++) oracle_keccak_f(states + 25 * i);
}

/* keccakSponge.Digest: NewKeccak(); Absorb(in[:in_len]); Squeeze(out_len) */
void oracle_sponge(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) {
  union { u64 w[25]; uint8_t b[200]; } st;
  memset(&st, 0, sizeof(st));
  size_t absorb_pos = 0, squeeze_pos = 136;
  for (size_t i = 0; i < in_len; i++) {
    if (absorb_pos == 136) { oracle_keccak_f(st.w); absorb_pos = 0; }
    st.b[absorb_pos++] = in[i];
  }
  squeeze_pos = 136;
  for (size_t i = 0; i < out_len; i++) {
    if (squeeze_pos == 136) { squeeze_pos = 0; absorb_pos = 0; oracle_keccak_f(st.w); }
    out[i] = st.b[squeeze_pos++];
  }
}

/* VerifyMerkleTreeProofs with the Keccak duplex as the 2-to-1 hash; layout as b200g16_keccak_merkle_paths */
void oracle_merkle_paths(const uint8_t* leaves, size_t leaf_len, const uint8_t* siblings, const uint8_t* auth,
                         const u64* indexes, unsigned height, size_t n, uint8_t* roots_out, int threads) {
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#endif
#pragma omp parallel for num_threads(threads) schedule(static)
  for (size_t i = 0; i < n; i++) {
    uint8_t cur[32], buf[64];
    oracle_sponge(leaves + i * leaf_len, leaf_len, cur, 32);
    for (unsigned level = 0; level < height; level++) {
      const uint8_t* sib = level == 0 ? siblings + i * 32 : auth + (i * (size_t)(height - 1) + (level - 1)) * 32;
      if ((indexes[i] >> level) & 1) { memcpy(buf, sib, 32); memcpy(buf + 32, cur, 32); }
      else { memcpy(buf, cur, 32); memcpy(buf + 32, sib, 32); }
      oracle_sponge(buf, 64, cur, 32);
    }
    memcpy(roots_out + i * 32, cur, 32);
  }
}

int oracle_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

#include "oracle_groth16.c"
