/* Second part of the C restatement (included at the end of oracle.c; TEST INFRASTRUCTURE ONLY, see its
 * header): Fp2, G2 in extended Jacobian coordinates, G2 MultiExp, single scalar multiplications, and the
 * Groth16 prover after the constraint solver.
 *
 * What it follows (gnark-crypto v0.14.1-0.20241217131346-b998989abdbe, gnark v0.11.0; go.mod:6-7; reached
 * from /root/reference/mt.go:496):
 *   fp2_*                 ecc/bn254/internal/fptower e2.go      : Fp[u]/(u^2+1), Karatsuba product
 *   g2x_*                 ecc/bn254/g2.go g2JacExtended          : add / mixed add / double
 *   oracle_msm_g2         ecc/bn254/multiexp.go (G2) MultiExp    : same bucket method as G1
 *   oracle_groth16_prove  backend/groth16/bn254/prove.go Prove   : wire filtering by InfinityA/B and by
 *                         public + committed wires, computeH, the five MultiExps, Ar / Bs / Krs assembly
 *                         (r, s are parameters so results are reproducible)
 * PARITY UNPINNED by the reference (no tests, arithmetic in absent modules); pinned against oracle/groth16.py
 * (python big ints + independent pairing) in tests/test_oracle_cpu.py.
 */

typedef struct { fe c0, c1; } fe2;
static const fe G2X0 = {{0x8e83b5d102bc2026ull, 0xdceb1935497b0172ull, 0xfbb8264797811adfull, 0x19573841af96503bull}};
static const fe G2X1 = {{0xafb4737da84c6140ull, 0x6043dd5a5802d8c4ull, 0x09e950fc52a02f86ull, 0x14fef0833aea7b6bull}};
static const fe G2Y0 = {{0x619dfa9d886be9f6ull, 0xfe7fd297f59e9b78ull, 0xff9e1a62231b7dfeull, 0x28fd7eebae9e4206ull}};
static const fe G2Y1 = {{0x64095b56c71856eeull, 0xdc57f922327d3cbbull, 0x55f935be33351076ull, 0x0da4a0e693fd6482ull}};

static inline void fp_neg(fe* r, const fe* a) {
  if (fe_is_zero(a)) { *r = *a; return; }
  fe z = {{0, 0, 0, 0}};
  fp_sub(r, &z, a);
}
static inline int fe2_is_zero(const fe2* a) { return fe_is_zero(&a->c0) && fe_is_zero(&a->c1); }
static inline void fp2_add(fe2* r, const fe2* a, const fe2* b) { fp_add(&r->c0, &a->c0, &b->c0); fp_add(&r->c1, &a->c1, &b->c1); }
static inline void fp2_sub(fe2* r, const fe2* a, const fe2* b) { fp_sub(&r->c0, &a->c0, &b->c0); fp_sub(&r->c1, &a->c1, &b->c1); }
static inline void fp2_neg(fe2* r, const fe2* a) { fp_neg(&r->c0, &a->c0); fp_neg(&r->c1, &a->c1); }
static inline void fp2_mul(fe2* r, const fe2* a, const fe2* b) {
  fe t0, t1, t2, sa, sb;
  fp_mul(&t0, &a->c0, &b->c0);
  fp_mul(&t1, &a->c1, &b->c1);
  fp_add(&sa, &a->c0, &a->c1);
  fp_add(&sb, &b->c0, &b->c1);
  fp_mul(&t2, &sa, &sb);
  fp_sub(&r->c0, &t0, &t1);
  fp_sub(&t2, &t2, &t0);
  fp_sub(&r->c1, &t2, &t1);
}
static inline void fp2_sqr(fe2* r, const fe2* a) {
  fe s, d, t;
  fp_add(&s, &a->c0, &a->c1);
  fp_sub(&d, &a->c0, &a->c1);
  fp_mul(&t, &a->c0, &a->c1);
  fp_mul(&r->c0, &s, &d);
  fp_add(&r->c1, &t, &t);
}
static void fp2_inv(fe2* r, const fe2* a) {
  fe n, t;
  fp_mul(&n, &a->c0, &a->c0);
  fp_mul(&t, &a->c1, &a->c1);
  fp_add(&n, &n, &t);
  fp_inv(&n, &n);
  fp_mul(&r->c0, &a->c0, &n);
  fp_mul(&t, &a->c1, &n);
  fp_neg(&r->c1, &t);
}

typedef struct { fe2 x, y; } g2a;
typedef struct { fe2 x, y, zz, zzz; } g2x;
static inline int g2a_is_inf(const g2a* p) { return fe2_is_zero(&p->x) && fe2_is_zero(&p->y); }
static inline void g2x_set_inf(g2x* p) { memset(p, 0, sizeof(*p)); }
static const fe2 FP2_ONE = {{{0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full}}, {{0, 0, 0, 0}}};

static void g2x_dbl_affine(g2x* r, const g2a* p) {
  fe2 U, V, W, S, X2, M, t;
  fp2_add(&U, &p->y, &p->y);
  fp2_sqr(&V, &U);
  fp2_mul(&W, &U, &V);
  fp2_mul(&S, &p->x, &V);
  fp2_sqr(&X2, &p->x);
  fp2_add(&M, &X2, &X2); fp2_add(&M, &M, &X2);
  fp2_sqr(&r->x, &M); fp2_sub(&r->x, &r->x, &S); fp2_sub(&r->x, &r->x, &S);
  fp2_sub(&t, &S, &r->x); fp2_mul(&t, &M, &t);
  fp2_mul(&r->y, &W, &p->y); fp2_sub(&r->y, &t, &r->y);
  r->zz = V; r->zzz = W;
}
static void g2x_dbl(g2x* p) {
  if (fe2_is_zero(&p->zz)) return;
  fe2 U, V, W, S, X2, M, t, X3, Y3;
  fp2_add(&U, &p->y, &p->y);
  fp2_sqr(&V, &U);
  fp2_mul(&W, &U, &V);
  fp2_mul(&S, &p->x, &V);
  fp2_sqr(&X2, &p->x);
  fp2_add(&M, &X2, &X2); fp2_add(&M, &M, &X2);
  fp2_sqr(&X3, &M); fp2_sub(&X3, &X3, &S); fp2_sub(&X3, &X3, &S);
  fp2_sub(&t, &S, &X3); fp2_mul(&t, &M, &t);
  fp2_mul(&Y3, &W, &p->y); fp2_sub(&Y3, &t, &Y3);
  p->x = X3; p->y = Y3;
  fp2_mul(&p->zz, &V, &p->zz);
  fp2_mul(&p->zzz, &W, &p->zzz);
}
static void g2x_madd(g2x* acc, const g2a* p, int neg) {
  if (g2a_is_inf(p)) return;
  fe2 py = p->y;
  if (neg) fp2_neg(&py, &py);
  if (fe2_is_zero(&acc->zz)) { acc->x = p->x; acc->y = py; acc->zz = FP2_ONE; acc->zzz = FP2_ONE; return; }
  fe2 Pq, Rq, PP, PPP, Q, X3, t;
  fp2_mul(&Pq, &p->x, &acc->zz); fp2_sub(&Pq, &Pq, &acc->x);
  fp2_mul(&Rq, &py, &acc->zzz); fp2_sub(&Rq, &Rq, &acc->y);
  if (fe2_is_zero(&Pq)) {
    if (fe2_is_zero(&Rq)) { g2a q = {p->x, py}; g2x_dbl_affine(acc, &q); }
    else g2x_set_inf(acc);
    return;
  }
  fp2_sqr(&PP, &Pq);
  fp2_mul(&PPP, &Pq, &PP);
  fp2_mul(&Q, &acc->x, &PP);
  fp2_sqr(&X3, &Rq); fp2_sub(&X3, &X3, &PPP); fp2_sub(&X3, &X3, &Q); fp2_sub(&X3, &X3, &Q);
  fp2_sub(&t, &Q, &X3); fp2_mul(&t, &Rq, &t);
  fp2_mul(&acc->y, &acc->y, &PPP); fp2_sub(&acc->y, &t, &acc->y);
  acc->x = X3;
  fp2_mul(&acc->zz, &acc->zz, &PP);
  fp2_mul(&acc->zzz, &acc->zzz, &PPP);
}
static void g2x_add(g2x* acc, const g2x* q) {
  if (fe2_is_zero(&q->zz)) return;
  if (fe2_is_zero(&acc->zz)) { *acc = *q; return; }
  fe2 U1, U2, S1, S2, Pq, Rq, PP, PPP, Q, X3, t;
  fp2_mul(&U1, &acc->x, &q->zz);
  fp2_mul(&U2, &q->x, &acc->zz);
  fp2_mul(&S1, &acc->y, &q->zzz);
  fp2_mul(&S2, &q->y, &acc->zzz);
  fp2_sub(&Pq, &U2, &U1);
  fp2_sub(&Rq, &S2, &S1);
  if (fe2_is_zero(&Pq)) {
    if (fe2_is_zero(&Rq)) g2x_dbl(acc); else g2x_set_inf(acc);
    return;
  }
  fp2_sqr(&PP, &Pq);
  fp2_mul(&PPP, &Pq, &PP);
  fp2_mul(&Q, &U1, &PP);
  fp2_sqr(&X3, &Rq); fp2_sub(&X3, &X3, &PPP); fp2_sub(&X3, &X3, &Q); fp2_sub(&X3, &X3, &Q);
  fp2_sub(&t, &Q, &X3); fp2_mul(&t, &Rq, &t);
  fp2_mul(&S1, &S1, &PPP); fp2_sub(&acc->y, &t, &S1);
  acc->x = X3;
  fp2_mul(&acc->zz, &acc->zz, &q->zz); fp2_mul(&acc->zz, &acc->zz, &PP);
  fp2_mul(&acc->zzz, &acc->zzz, &q->zzz); fp2_mul(&acc->zzz, &acc->zzz, &PPP);
}
static void g2x_to_affine(g2a* r, const g2x* p) {
  if (fe2_is_zero(&p->zz)) { memset(r, 0, sizeof(*r)); return; }
  fe2 zi, a, izz;
  fp2_inv(&zi, &p->zzz);
  fp2_mul(&a, &p->zz, &zi);
  fp2_sqr(&izz, &a);
  fp2_mul(&r->x, &p->x, &izz);
  fp2_mul(&r->y, &p->y, &zi);
}

/* signed-digit recoding shared by the G2 MultiExp (same rule as oracle_msm_g1) */
static int msm_shape(size_t n, int* c_out) {
  int c = best_c(n);
  int W = (254 + c - 1) / c;
  if (c * W < 256) {
    u64 topbits = 254 - (u64)(W - 1) * c;
    if (topbits >= (u64)(c - 1)) W += 1;
  }
  *c_out = c;
  return W;
}
static int32_t* msm_digits(const fe* sc, size_t n, int c, int W, int threads) {
  int32_t* digits = (int32_t*)malloc((size_t)W * n * sizeof(int32_t));
  if (!digits) return NULL;
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
  return digits;
}

int oracle_msm_g2(const u64* points, const u64* scalars, size_t n, int threads, u64* out) {
  const g2a* pts = (const g2a*)points;
  if (n == 0) { memset(out, 0, 128); return 0; }
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  threads = 1;
#endif
  int c, W = msm_shape(n, &c);
  int32_t* digits = msm_digits((const fe*)scalars, n, c, W, threads);
  if (!digits) return -1;
  int S = (2 * threads + W - 1) / W;
  if (S < 1) S = 1;
  if ((size_t)S > n / 1024 + 1) S = (int)(n / 1024 + 1);
  size_t nb = (size_t)1 << (c - 1);
  g2x* sums = (g2x*)calloc((size_t)W * S, sizeof(g2x));
  int fail = 0;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
  for (int task = 0; task < W * S; task++) {
    int w = task / S, sp = task % S;
    size_t lo = n * (size_t)sp / S, hi = n * (size_t)(sp + 1) / S;
    g2x* buckets = (g2x*)calloc(nb, sizeof(g2x));
    if (!buckets) { fail = 1; continue; }
    const int32_t* dg = digits + (size_t)w * n;
    for (size_t i = lo; i < hi; i++) {
      int32_t d = dg[i];
      if (d > 0) g2x_madd(&buckets[d - 1], &pts[i], 0);
      else if (d < 0) g2x_madd(&buckets[-d - 1], &pts[i], 1);
    }
    g2x run, acc; g2x_set_inf(&run); g2x_set_inf(&acc);
    for (size_t b = nb; b-- > 0;) { g2x_add(&run, &buckets[b]); g2x_add(&acc, &run); }
    sums[task] = acc;
    free(buckets);
  }
  g2x total; g2x_set_inf(&total);
  for (int w = W - 1; w >= 0; w--) {
    for (int k = 0; k < c; k++) g2x_dbl(&total);
    for (int sp = 0; sp < S; sp++) g2x_add(&total, &sums[w * S + sp]);
  }
  g2a res; g2x_to_affine(&res, &total);
  memcpy(out, &res, 128);
  free(sums); free(digits);
  return fail ? -1 : 0;
}

/* out = k * P (affine Montgomery), k fr Montgomery: double-and-add, independent of the bucket method */
static void g1_mul_fe(g1a* r, const g1a* p, const fe* k_mont) {
  fe k; fr_from_mont(&k, k_mont);
  g1x acc; g1x_set_inf(&acc);
  for (int i = 255; i >= 0; i--) { g1x_dbl(&acc); if ((k.l[i >> 6] >> (i & 63)) & 1) g1x_madd(&acc, p, 0); }
  g1x_to_affine(r, &acc);
}
static void g2_mul_fe(g2a* r, const g2a* p, const fe* k_mont) {
  fe k; fr_from_mont(&k, k_mont);
  g2x acc; g2x_set_inf(&acc);
  for (int i = 255; i >= 0; i--) { g2x_dbl(&acc); if ((k.l[i >> 6] >> (i & 63)) & 1) g2x_madd(&acc, p, 0); }
  g2x_to_affine(r, &acc);
}
void oracle_g1_mul(const u64* p, const u64* k_mont, u64* out) { g1a r; g1_mul_fe(&r, (const g1a*)p, (const fe*)k_mont); memcpy(out, &r, 64); }
void oracle_g2_mul(const u64* p, const u64* k_mont, u64* out) { g2a r; g2_mul_fe(&r, (const g2a*)p, (const fe*)k_mont); memcpy(out, &r, 128); }
void oracle_g2_gen_mul(const u64* k_mont, u64* out) {
  g2a G; G.x.c0 = G2X0; G.x.c1 = G2X1; G.y.c0 = G2Y0; G.y.c1 = G2Y1;
  g2a r; g2_mul_fe(&r, &G, (const fe*)k_mont); memcpy(out, &r, 128);
}
static void g1_sum(g1a* r, const g1a** pts, int n) {
  g1x acc; g1x_set_inf(&acc);
  for (int i = 0; i < n; i++) g1x_madd(&acc, pts[i], 0);
  g1x_to_affine(r, &acc);
}

/* Groth16 prover after Solve (gnark backend/groth16/bn254/prove.go).  All arrays in gnark-crypto memory layout.
 *   pk_a[n_a], pk_b1[n_b], pk_k[n_k], pk_z[N-1] G1Affine; pk_b2[n_b] G2Affine   (points at infinity already
 *   filtered out of A / B, as gnark's Setup stores them); inf_a / inf_b / k_skip: n_wires flags
 *   wires[n_wires], a / b / c [n_constraints] Fr Montgomery; r, s Fr Montgomery.
 * out: ar[8] bs[16] krs[8] then the five MultiExp results msm_a[8] msm_b1[8] msm_k[8] msm_z[8] msm_b2[16]
 * (80 u64); h_out (N Fr, bit-reversed; may be NULL). */
int oracle_groth16_prove(int log2n, size_t n_wires, const u64* pk_a, size_t n_a, const u64* pk_b1, size_t n_b,
                         const u64* pk_k, size_t n_k, const u64* pk_z, const u64* pk_b2, const u64* alpha,
                         const u64* beta, const u64* delta, const u64* beta2, const u64* delta2, const uint8_t* inf_a,
                         const uint8_t* inf_b, const uint8_t* k_skip, const u64* wires, const u64* a, const u64* b,
                         const u64* c, size_t n_constraints, const u64* r_mont, const u64* s_mont, int threads,
                         u64* out, u64* h_out) {
  size_t N = (size_t)1 << log2n;
  if (n_constraints > N) return -2;
  const fe* w = (const fe*)wires;
  fe* wa = (fe*)malloc((n_a ? n_a : 1) * sizeof(fe));
  fe* wb = (fe*)malloc((n_b ? n_b : 1) * sizeof(fe));
  fe* wk = (fe*)malloc((n_k ? n_k : 1) * sizeof(fe));
  fe* abc = (fe*)calloc(3 * N, sizeof(fe));
  if (!wa || !wb || !wk || !abc) return -1;
  size_t ia = 0, ib = 0, ik = 0;
  for (size_t i = 0; i < n_wires; i++) {
    if (!inf_a[i]) { if (ia >= n_a) return -3; wa[ia++] = w[i]; }
    if (!inf_b[i]) { if (ib >= n_b) return -3; wb[ib++] = w[i]; }
    if (!k_skip[i]) { if (ik >= n_k) return -3; wk[ik++] = w[i]; }
  }
  if (ia != n_a || ib != n_b || ik != n_k) return -3;
  memcpy(abc, a, n_constraints * sizeof(fe));
  memcpy(abc + N, b, n_constraints * sizeof(fe));
  memcpy(abc + 2 * N, c, n_constraints * sizeof(fe));
  if (oracle_compute_h((u64*)abc, (u64*)(abc + N), (u64*)(abc + 2 * N), log2n, threads)) return -1;
  if (h_out) memcpy(h_out, abc, N * sizeof(fe));
  g1a A, B1, K, Z;
  g2a B2;
  if (oracle_msm_g1(pk_a, (const u64*)wa, n_a, threads, (u64*)&A)) return -1;
  if (oracle_msm_g1(pk_b1, (const u64*)wb, n_b, threads, (u64*)&B1)) return -1;
  if (oracle_msm_g1(pk_k, (const u64*)wk, n_k, threads, (u64*)&K)) return -1;
  if (oracle_msm_g1(pk_z, (const u64*)abc, N - 1, threads, (u64*)&Z)) return -1;
  if (oracle_msm_g2(pk_b2, (const u64*)wb, n_b, threads, (u64*)&B2)) return -1;
  const fe* r = (const fe*)r_mont; const fe* s = (const fe*)s_mont;
  fe rs, nrs, zero = {{0, 0, 0, 0}};
  fr_mul(&rs, r, s); fr_sub(&nrs, &zero, &rs);
  g1a rd, sd, krd, ar, bs1, sar, rbs1, krs;
  g1_mul_fe(&rd, (const g1a*)delta, r);
  g1_mul_fe(&sd, (const g1a*)delta, s);
  g1_mul_fe(&krd, (const g1a*)delta, &nrs);
  { const g1a* t[3] = {&A, (const g1a*)alpha, &rd}; g1_sum(&ar, t, 3); }
  { const g1a* t[3] = {&B1, (const g1a*)beta, &sd}; g1_sum(&bs1, t, 3); }
  g1_mul_fe(&sar, &ar, s);
  g1_mul_fe(&rbs1, &bs1, r);
  { const g1a* t[5] = {&K, &Z, &krd, &sar, &rbs1}; g1_sum(&krs, t, 5); }
  g2a sd2, bs;
  g2_mul_fe(&sd2, (const g2a*)delta2, s);
  { g2x acc; g2x_set_inf(&acc); g2x_madd(&acc, &B2, 0); g2x_madd(&acc, (const g2a*)beta2, 0); g2x_madd(&acc, &sd2, 0);
    g2x_to_affine(&bs, &acc); }
  memcpy(out, &ar, 64); memcpy(out + 8, &bs, 128); memcpy(out + 24, &krs, 64);
  memcpy(out + 32, &A, 64); memcpy(out + 40, &B1, 64); memcpy(out + 48, &K, 64); memcpy(out + 56, &Z, 64);
  memcpy(out + 64, &B2, 128);
  free(wa); free(wb); free(wk); free(abc);
  return 0;
}
