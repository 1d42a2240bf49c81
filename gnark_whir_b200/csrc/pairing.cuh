// Optimal-ate pairing on BN254 for groth16.Verify (reference call site /root/reference/mt.go:497;
// gnark v0.11.0 backend/groth16/bn254/verify.go -> gnark-crypto ecc/bn254 PairingCheck / Pair,
// MillerLoop + FinalExponentiation in ecc/bn254/pairing.go).
//
// Constant work per proof (4 + 2 Miller loops, 2 final exponentiations), so this is written for
// clarity and for being checkable, not for throughput: one thread per Miller loop, the same
// source compiled for the host (B200_HD) so the CPU test-suite pins it against the oracle.
//
//   GT = Fp12 = Fp2[w]/(w^6 - xi), xi = 9 + u, element = 6 Fp2 coefficients of w^0..w^5.
//   gnark's tower E12 = E6[w]/(w^2 - v), E6 = E2[v]/(v^3 - xi) is the same field with
//   C0 = (c0, c2, c4), C1 = (c1, c3, c5); to_gnark_layout() emits that order.
//   Untwist (x, y) -> (x w^2, y w^3); line through T, Q evaluated at P = (xP, yP):
//       l = yP - lambda xP w + (lambda xT - yT) w^3,
//   scaled by the slope's denominator (an Fp2 factor, killed by the final exponentiation) so that
//   T stays in homogeneous projective coordinates and the loop has no inversion:
//       tangent at T = (X, Y, Z):  2YZ yP - 3X^2 xP w + (Y^2 - 3b'Z^2) w^3
//       chord T, Q = (xQ, yQ):     L yP - Th xP w + (Th xQ - L yQ) w^3,   Th = Y - yQ Z, L = X - xQ Z
//   Miller loop over 6u+2 (MSB first) plus the two Frobenius lines; final exponentiation: easy part
//   (p^6-1)(p^2+1) by conjugation / inversion / Frobenius, hard part (p^4-p^2+1)/r =
//   p^3 + (6u^2+1) p^2 + (-36u^3-18u^2-12u+1) p + (-36u^3-30u^2-18u-2) by three exponentiations by u
//   and the vectorial addition chain of Scott et al. (inverses are conjugates in the cyclotomic subgroup).
#pragma once
#include "ec.cuh"
#include "pairing_constants.h"

// free functions defined in this header and included from several translation units: inline linkage,
// but kept out of line on the device (code size)
#if defined(__CUDACC__)
#define B200_FN inline __host__ __device__ __noinline__
#else
#define B200_FN inline
#endif

namespace b200 {

struct Fp12 {
  Fp2 c[6];
};

B200_HD Fp fp_from_words(const uint32_t* v) {
  Fp r;
  for (int i = 0; i < 8; i++) r.l[i] = v[i];
  return r;
}

B200_HD Fp2 twist_frob_x() { constexpr uint32_t v[2][8] = B200_TWIST_FROB_X; return {fp_from_words(v[0]), fp_from_words(v[1])}; }
B200_HD Fp2 twist_frob_y() { constexpr uint32_t v[2][8] = B200_TWIST_FROB_Y; return {fp_from_words(v[0]), fp_from_words(v[1])}; }
B200_HD Fp2 twist_b() { constexpr uint32_t v[2][8] = B200_TWIST_B; return {fp_from_words(v[0]), fp_from_words(v[1])}; }
B200_HD Fp2 twist_3b() { constexpr uint32_t v[2][8] = B200_TWIST_3B; return {fp_from_words(v[0]), fp_from_words(v[1])}; }
B200_HD Fp frob2_gamma(int i) { constexpr uint32_t v[6][8] = B200_FROB2_GAMMA; return fp_from_words(v[i]); }
B200_HD Fp2 frob1_gamma(int i) { constexpr uint32_t v[6][2][8] = B200_FROB1_GAMMA; return {fp_from_words(v[i][0]), fp_from_words(v[i][1])}; }
B200_HD Fp2 frob3_gamma(int i) { constexpr uint32_t v[6][2][8] = B200_FROB3_GAMMA; return {fp_from_words(v[i][0]), fp_from_words(v[i][1])}; }
B200_HD Fp fp_inv2() { constexpr uint32_t v[8] = B200_FP_INV2; return fp_from_words(v); }

B200_HD Fp2 fp2_conj(const Fp2& a) { return {a.c0, Fp::neg(a.c1)}; }
B200_HD Fp2 fp2_scale(const Fp2& a, const Fp& s) { return {Fp::mul_call(a.c0, s), Fp::mul_call(a.c1, s)}; }

// a * (9 + u) = (9 a0 - a1) + (9 a1 + a0) u
B200_FN Fp2 mul_xi(const Fp2& a) {
  Fp2 t = Fp2::dbl(Fp2::dbl(Fp2::dbl(a)));
  t = Fp2::add(t, a);
  return {Fp::sub(t.c0, a.c1), Fp::add(t.c1, a.c0)};
}

B200_HD Fp12 f12_one() {
  Fp12 r;
  r.c[0] = Fp2::one();
  for (int i = 1; i < 6; i++) r.c[i] = Fp2::zero();
  return r;
}

B200_HD bool f12_is_one(const Fp12& a) {
  bool ok = a.c[0] == Fp2::one();
  for (int i = 1; i < 6; i++) ok = ok && a.c[i].is_zero();
  return ok;
}

B200_FN Fp12 f12_mul(const Fp12& a, const Fp12& b) {
  Fp2 t[11];
  for (int k = 0; k < 11; k++) t[k] = Fp2::zero();
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) t[i + j] = Fp2::add(t[i + j], Fp2::mul(a.c[i], b.c[j]));
  Fp12 r;
  for (int k = 0; k < 5; k++) r.c[k] = Fp2::add(t[k], mul_xi(t[k + 6]));
  r.c[5] = t[5];
  return r;
}

// a^2: 15 cross products (doubled) + 6 squares instead of 36 products
B200_FN Fp12 f12_sqr(const Fp12& a) {
  Fp2 t[11];
  for (int k = 0; k < 11; k++) t[k] = Fp2::zero();
  for (int i = 0; i < 6; i++)
    for (int j = i + 1; j < 6; j++) t[i + j] = Fp2::add(t[i + j], Fp2::mul(a.c[i], a.c[j]));
  for (int k = 0; k < 11; k++) t[k] = Fp2::dbl(t[k]);
  for (int i = 0; i < 6; i++) t[2 * i] = Fp2::add(t[2 * i], Fp2::sqr(a.c[i]));
  Fp12 r;
  for (int k = 0; k < 5; k++) r.c[k] = Fp2::add(t[k], mul_xi(t[k + 6]));
  r.c[5] = t[5];
  return r;
}

// f * (l0 + l1 w + l3 w^3)
B200_FN Fp12 f12_mul_line(const Fp12& f, const Fp2& l0, const Fp2& l1, const Fp2& l3) {
  Fp2 t[9];
  for (int k = 0; k < 9; k++) t[k] = Fp2::zero();
  for (int i = 0; i < 6; i++) {
    t[i] = Fp2::add(t[i], Fp2::mul(f.c[i], l0));
    t[i + 1] = Fp2::add(t[i + 1], Fp2::mul(f.c[i], l1));
    t[i + 3] = Fp2::add(t[i + 3], Fp2::mul(f.c[i], l3));
  }
  Fp12 r;
  for (int k = 0; k < 3; k++) r.c[k] = Fp2::add(t[k], mul_xi(t[k + 6]));
  for (int k = 3; k < 6; k++) r.c[k] = t[k];
  return r;
}

// f^(p^6): w -> -w
B200_HD Fp12 f12_conj(const Fp12& a) {
  Fp12 r = a;
  r.c[1] = Fp2::neg(a.c[1]);
  r.c[3] = Fp2::neg(a.c[3]);
  r.c[5] = Fp2::neg(a.c[5]);
  return r;
}

// f^(p^2): Fp2 is fixed, w^i -> xi^(i (p^2-1)/6) w^i with the factor in Fp
B200_FN Fp12 f12_frob2(const Fp12& a) {
  Fp12 r;
  r.c[0] = a.c[0];
  for (int i = 1; i < 6; i++) r.c[i] = fp2_scale(a.c[i], frob2_gamma(i));
  return r;
}

// f^p and f^(p^3): coefficients conjugated (odd power of the Fp2 Frobenius) and scaled by xi^(i (p^k-1)/6)
B200_FN Fp12 f12_frob1(const Fp12& a) {
  Fp12 r;
  r.c[0] = fp2_conj(a.c[0]);
  for (int i = 1; i < 6; i++) r.c[i] = Fp2::mul(fp2_conj(a.c[i]), frob1_gamma(i));
  return r;
}
B200_FN Fp12 f12_frob3(const Fp12& a) {
  Fp12 r;
  r.c[0] = fp2_conj(a.c[0]);
  for (int i = 1; i < 6; i++) r.c[i] = Fp2::mul(fp2_conj(a.c[i]), frob3_gamma(i));
  return r;
}

// 1/f = conj(f) / (f conj(f)); the norm f conj(f) lies in Fp6 = Fp2[v]/(v^3 - xi), v = w^2
B200_FN Fp12 f12_inv(const Fp12& f) {
  Fp12 cf = f12_conj(f);
  Fp12 n = f12_mul(f, cf);
  const Fp2 a0 = n.c[0], a1 = n.c[2], a2 = n.c[4];
  Fp2 A = Fp2::sub(Fp2::sqr(a0), mul_xi(Fp2::mul(a1, a2)));
  Fp2 B = Fp2::sub(mul_xi(Fp2::sqr(a2)), Fp2::mul(a0, a1));
  Fp2 C = Fp2::sub(Fp2::sqr(a1), Fp2::mul(a0, a2));
  Fp2 F = Fp2::add(Fp2::mul(a0, A), mul_xi(Fp2::add(Fp2::mul(a2, B), Fp2::mul(a1, C))));
  Fp2 Fi = Fp2::inv(F);
  Fp12 inv6;
  for (int i = 0; i < 6; i++) inv6.c[i] = Fp2::zero();
  inv6.c[0] = Fp2::mul(A, Fi);
  inv6.c[2] = Fp2::mul(B, Fi);
  inv6.c[4] = Fp2::mul(C, Fi);
  return f12_mul(cf, inv6);
}

// a^e, e given as `bits` bits in little-endian 64-bit words
B200_FN Fp12 f12_pow(const Fp12& a, const uint64_t* e, int bits) {
  Fp12 r = f12_one();
  for (int i = bits - 1; i >= 0; i--) {
    r = f12_sqr(r);
    if ((e[i >> 6] >> (i & 63)) & 1) r = f12_mul(r, a);
  }
  return r;
}

// a^u, u the curve parameter (63 bits)
B200_FN Fp12 f12_pow_u(const Fp12& a) {
  const uint64_t u = B200_BN_U;
  return f12_pow(a, &u, 63);
}

// Twist point in homogeneous projective coordinates (x = X/Z, y = Y/Z)
struct G2Proj {
  Fp2 x, y, z;
};

// T <- 2T, f <- f * tangent(P)
B200_FN void miller_double(G2Proj& T, const Fp& xp, const Fp& yp, Fp12& f) {
  const Fp inv2 = fp_inv2();
  Fp2 A = fp2_scale(Fp2::mul(T.x, T.y), inv2);
  Fp2 B = Fp2::sqr(T.y);
  Fp2 C = Fp2::sqr(T.z);
  Fp2 E = Fp2::mul(twist_3b(), C);                    // 3 b' Z^2
  Fp2 F = Fp2::add(Fp2::dbl(E), E);
  Fp2 G = fp2_scale(Fp2::add(B, F), inv2);
  Fp2 H = Fp2::sub(Fp2::sqr(Fp2::add(T.y, T.z)), Fp2::add(B, C));  // 2YZ
  Fp2 J = Fp2::sqr(T.x);
  Fp2 E2 = Fp2::sqr(E);
  Fp2 l0 = fp2_scale(H, yp);
  Fp2 l1 = Fp2::neg(fp2_scale(Fp2::add(Fp2::dbl(J), J), xp));
  Fp2 l3 = Fp2::sub(B, E);
  T.x = Fp2::mul(A, Fp2::sub(B, F));
  T.y = Fp2::sub(Fp2::sqr(G), Fp2::add(Fp2::dbl(E2), E2));
  T.z = Fp2::mul(B, H);
  f = f12_mul_line(f, l0, l1, l3);
}

// T <- T + Q (Q affine, Q != +-T), f <- f * chord(P)
B200_FN void miller_add(G2Proj& T, const Affine<Fp2>& Q, const Fp& xp, const Fp& yp, Fp12& f) {
  Fp2 th = Fp2::sub(T.y, Fp2::mul(Q.y, T.z));
  Fp2 la = Fp2::sub(T.x, Fp2::mul(Q.x, T.z));
  Fp2 C = Fp2::sqr(th);
  Fp2 D = Fp2::sqr(la);
  Fp2 E = Fp2::mul(la, D);
  Fp2 F = Fp2::mul(T.z, C);
  Fp2 G = Fp2::mul(T.x, D);
  Fp2 H = Fp2::sub(Fp2::add(E, F), Fp2::dbl(G));
  Fp2 l0 = fp2_scale(la, yp);
  Fp2 l1 = Fp2::neg(fp2_scale(th, xp));
  Fp2 l3 = Fp2::sub(Fp2::mul(th, Q.x), Fp2::mul(la, Q.y));
  T.x = Fp2::mul(la, H);
  T.y = Fp2::sub(Fp2::mul(th, Fp2::sub(G, H)), Fp2::mul(E, T.y));
  T.z = Fp2::mul(T.z, E);
  f = f12_mul_line(f, l0, l1, l3);
}

B200_HD Affine<Fp2> twist_frobenius(const Affine<Fp2>& q) {
  return {Fp2::mul(fp2_conj(q.x), twist_frob_x()), Fp2::mul(fp2_conj(q.y), twist_frob_y())};
}

B200_FN Fp12 miller_loop(const Affine<Fp>& P, const Affine<Fp2>& Q) {
  Fp12 f = f12_one();
  if (P.is_inf() || Q.is_inf()) return f;
  constexpr uint64_t loop[2] = B200_ATE_LOOP;
  G2Proj T = {Q.x, Q.y, Fp2::one()};
  for (int i = B200_ATE_LOOP_BITS - 2; i >= 0; i--) {
    f = f12_sqr(f);
    miller_double(T, P.x, P.y, f);
    if ((loop[i >> 6] >> (i & 63)) & 1) miller_add(T, Q, P.x, P.y, f);
  }
  Affine<Fp2> Q1 = twist_frobenius(Q);
  Affine<Fp2> Q2 = twist_frobenius(Q1);
  Q2.y = Fp2::neg(Q2.y);
  miller_add(T, Q1, P.x, P.y, f);
  miller_add(T, Q2, P.x, P.y, f);
  return f;
}

// f^((p^12-1)/r * s), s = 2u(6u^2+3u+1): the cofactor gnark-crypto's FinalExponentiation carries
// (ecc/bn254/pairing.go, "we use instead d = s (p^6-1)(p^2+1)(p^4-p^2+1)/r"), so that the GT
// element equals bn254.Pair's.  with_cofactor = false gives the plain reduced pairing.
B200_FN Fp12 final_exponentiation(const Fp12& f, bool with_cofactor) {
  Fp12 t = f12_mul(f12_conj(f), f12_inv(f));  // f^(p^6-1)
  t = f12_mul(f12_frob2(t), t);               // ^(p^2+1): t is now in the cyclotomic subgroup, 1/t = conj(t)
  // hard part
  Fp12 fu = f12_pow_u(t), fu2 = f12_pow_u(fu), fu3 = f12_pow_u(fu2);
  Fp12 y0 = f12_mul(f12_mul(f12_frob1(t), f12_frob2(t)), f12_frob3(t));
  Fp12 y1 = f12_conj(t);
  Fp12 y2 = f12_frob2(fu2);
  Fp12 y3 = f12_conj(f12_frob1(fu));
  Fp12 y4 = f12_conj(f12_mul(fu, f12_frob1(fu2)));
  Fp12 y5 = f12_conj(fu2);
  Fp12 y6 = f12_conj(f12_mul(fu3, f12_frob1(fu3)));
  // y0 y1^2 y2^6 y3^12 y4^18 y5^30 y6^36
  Fp12 t0 = f12_sqr(y6);
  t0 = f12_mul(t0, y4);
  t0 = f12_mul(t0, y5);
  Fp12 t1 = f12_mul(y3, y5);
  t1 = f12_mul(t1, t0);
  t0 = f12_mul(t0, y2);
  t1 = f12_sqr(t1);
  t1 = f12_mul(t1, t0);
  t1 = f12_sqr(t1);
  t0 = f12_mul(t1, y1);
  t1 = f12_mul(t1, y0);
  t0 = f12_sqr(t0);
  t = f12_mul(t0, t1);
  if (with_cofactor) {
    constexpr uint64_t s[3] = B200_FINAL_EXP_COFACTOR;
    t = f12_pow(t, s, B200_FINAL_EXP_COFACTOR_BITS);
  }
  return t;
}

B200_HD bool g1_on_curve(const Affine<Fp>& p) {
  if (p.is_inf()) return true;
  constexpr uint32_t three[8] = B200_FP_THREE;
  Fp rhs = Fp::add(Fp::mul(Fp::sqr(p.x), p.x), fp_from_words(three));
  return Fp::sqr(p.y) == rhs;
}

B200_HD bool g2_on_curve(const Affine<Fp2>& q) {
  if (q.is_inf()) return true;
  Fp2 rhs = Fp2::add(Fp2::mul(Fp2::sqr(q.x), q.x), twist_b());
  return Fp2::sqr(q.y) == rhs;
}

// [r]Q == infinity (the twist has a large cofactor, G1 has none)
B200_FN bool g2_in_subgroup(const Affine<Fp2>& q) {
  if (q.is_inf()) return true;
  if (!g2_on_curve(q)) return false;
  constexpr uint32_t r[8] = B200_FR_MOD;
  XYZZ<Fp2> acc = XYZZ<Fp2>::inf();
  for (int i = 253; i >= 0; i--) {
    acc.dbl();
    if ((r[i >> 5] >> (i & 31)) & 1) acc.madd(q);
  }
  return acc.is_inf();
}

// gnark-crypto E12 memory order: C0.B0, C0.B1, C0.B2, C1.B0, C1.B1, C1.B2 (each E2 = A0, A1)
B200_HD void to_gnark_layout(const Fp12& f, uint64_t out[48]) {
  const int order[6] = {0, 2, 4, 1, 3, 5};
  for (int k = 0; k < 6; k++) {
    const Fp2& c = f.c[order[k]];
    for (int i = 0; i < 4; i++) {
      out[8 * k + i] = (uint64_t)c.c0.l[2 * i] | ((uint64_t)c.c0.l[2 * i + 1] << 32);
      out[8 * k + 4 + i] = (uint64_t)c.c1.l[2 * i] | ((uint64_t)c.c1.l[2 * i + 1] << 32);
    }
  }
}

}  // namespace b200
