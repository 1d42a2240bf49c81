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
//       l = yP - lambda xP w + (lambda xT - yT) w^3                  (affine, one Fp2 inversion)
//   Miller loop over 6u+2 (MSB first) plus the two Frobenius lines; final exponentiation
//   (p^6-1)(p^2+1) by conjugation / inversion / Frobenius, hard part by plain square-and-multiply.
#pragma once
#include "ec.cuh"
#include "pairing_constants.h"

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
B200_HD Fp frob2_gamma(int i) { constexpr uint32_t v[6][8] = B200_FROB2_GAMMA; return fp_from_words(v[i]); }

B200_HD Fp2 fp2_conj(const Fp2& a) { return {a.c0, Fp::neg(a.c1)}; }
B200_HD Fp2 fp2_scale(const Fp2& a, const Fp& s) { return {Fp::mul_call(a.c0, s), Fp::mul_call(a.c1, s)}; }

// a * (9 + u) = (9 a0 - a1) + (9 a1 + a0) u
B200_HD_NOINLINE Fp2 mul_xi(const Fp2& a) {
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

B200_HD_NOINLINE Fp12 f12_mul(const Fp12& a, const Fp12& b) {
  Fp2 t[11];
  for (int k = 0; k < 11; k++) t[k] = Fp2::zero();
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) t[i + j] = Fp2::add(t[i + j], Fp2::mul(a.c[i], b.c[j]));
  Fp12 r;
  for (int k = 0; k < 5; k++) r.c[k] = Fp2::add(t[k], mul_xi(t[k + 6]));
  r.c[5] = t[5];
  return r;
}

// f * (l0 + l1 w + l3 w^3), l0 in Fp
B200_HD_NOINLINE Fp12 f12_mul_line(const Fp12& f, const Fp& l0, const Fp2& l1, const Fp2& l3) {
  Fp2 t[9];
  for (int k = 0; k < 9; k++) t[k] = Fp2::zero();
  for (int i = 0; i < 6; i++) {
    t[i] = Fp2::add(t[i], fp2_scale(f.c[i], l0));
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
B200_HD_NOINLINE Fp12 f12_frob2(const Fp12& a) {
  Fp12 r;
  r.c[0] = a.c[0];
  for (int i = 1; i < 6; i++) r.c[i] = fp2_scale(a.c[i], frob2_gamma(i));
  return r;
}

// 1/f = conj(f) / (f conj(f)); the norm f conj(f) lies in Fp6 = Fp2[v]/(v^3 - xi), v = w^2
B200_HD_NOINLINE Fp12 f12_inv(const Fp12& f) {
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
B200_HD_NOINLINE Fp12 f12_pow(const Fp12& a, const uint64_t* e, int bits) {
  Fp12 r = f12_one();
  for (int i = bits - 1; i >= 0; i--) {
    r = f12_mul(r, r);
    if ((e[i >> 6] >> (i & 63)) & 1) r = f12_mul(r, a);
  }
  return r;
}

// One Miller step on the twist: T <- T + Q (or 2T when tangent), f <- f * line(P).
B200_HD_NOINLINE void miller_step(Affine<Fp2>& T, const Affine<Fp2>& Q, bool tangent, const Fp& xp, const Fp& yp,
                                  Fp12& f) {
  Fp2 lam;
  if (tangent) {
    Fp2 x2 = Fp2::sqr(T.x);
    lam = Fp2::mul(Fp2::add(Fp2::dbl(x2), x2), Fp2::inv(Fp2::dbl(T.y)));
  } else {
    lam = Fp2::mul(Fp2::sub(Q.y, T.y), Fp2::inv(Fp2::sub(Q.x, T.x)));
  }
  Fp2 x3 = Fp2::sub(Fp2::sub(Fp2::sqr(lam), T.x), Q.x);
  Fp2 y3 = Fp2::sub(Fp2::mul(lam, Fp2::sub(T.x, x3)), T.y);
  Fp2 l1 = Fp2::neg(fp2_scale(lam, xp));
  Fp2 l3 = Fp2::sub(Fp2::mul(lam, T.x), T.y);
  f = f12_mul_line(f, yp, l1, l3);
  T.x = x3;
  T.y = y3;
}

B200_HD Affine<Fp2> twist_frobenius(const Affine<Fp2>& q) {
  return {Fp2::mul(fp2_conj(q.x), twist_frob_x()), Fp2::mul(fp2_conj(q.y), twist_frob_y())};
}

B200_HD_NOINLINE Fp12 miller_loop(const Affine<Fp>& P, const Affine<Fp2>& Q) {
  Fp12 f = f12_one();
  if (P.is_inf() || Q.is_inf()) return f;
  constexpr uint64_t loop[2] = B200_ATE_LOOP;
  Affine<Fp2> T = Q;
  for (int i = B200_ATE_LOOP_BITS - 2; i >= 0; i--) {
    f = f12_mul(f, f);
    miller_step(T, T, true, P.x, P.y, f);
    if ((loop[i >> 6] >> (i & 63)) & 1) miller_step(T, Q, false, P.x, P.y, f);
  }
  Affine<Fp2> Q1 = twist_frobenius(Q);
  Affine<Fp2> Q2 = twist_frobenius(Q1);
  Q2.y = Fp2::neg(Q2.y);
  miller_step(T, Q1, false, P.x, P.y, f);
  miller_step(T, Q2, false, P.x, P.y, f);
  return f;
}

// f^((p^12-1)/r * s), s = 2u(6u^2+3u+1): the cofactor gnark-crypto's FinalExponentiation carries
// (ecc/bn254/pairing.go, "we use instead d = s (p^6-1)(p^2+1)(p^4-p^2+1)/r"), so that the GT
// element equals bn254.Pair's.  with_cofactor = false gives the plain reduced pairing.
B200_HD_NOINLINE Fp12 final_exponentiation(const Fp12& f, bool with_cofactor) {
  Fp12 t = f12_mul(f12_conj(f), f12_inv(f));  // f^(p^6-1)
  t = f12_mul(f12_frob2(t), t);               // ^(p^2+1)
  constexpr uint64_t hard[12] = B200_HARD_EXP;
  t = f12_pow(t, hard, B200_HARD_EXP_BITS);
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
B200_HD_NOINLINE bool g2_in_subgroup(const Affine<Fp2>& q) {
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
