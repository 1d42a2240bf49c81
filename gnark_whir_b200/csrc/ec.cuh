// BN254 G1 / G2 group arithmetic, templated on the coordinate field (Fp or Fp2).
//
// Affine points use gnark-crypto's layout (G1Affine{X,Y}, G2Affine{X{A0,A1},Y{A0,A1}},
// infinity = all-zero).  Accumulators are extended Jacobian "XYZZ"
// (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2): mixed add 8M+2S, full add 12M+2S, double 6M+3S+...,
// the same coordinate system gnark-crypto's MultiExp uses for its extended-Jacobian
// buckets (ecc/bn254/g1.go g1JacExtended; reached from mt.go:496).  Curve a = 0.
#pragma once
#include "field.cuh"

namespace b200 {

// ---- Fp2 = Fp[u]/(u^2+1) -----------------------------------------------------------------
struct alignas(16) Fp2 {
  Fp c0, c1;
  static B200_HD Fp2 zero() { return {Fp::zero(), Fp::zero()}; }
  static B200_HD Fp2 one() { return {Fp::one(), Fp::zero()}; }
  B200_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  B200_HD bool operator==(const Fp2& b) const { return c0 == b.c0 && c1 == b.c1; }
  static B200_HD Fp2 add(const Fp2& a, const Fp2& b) { return {Fp::add(a.c0, b.c0), Fp::add(a.c1, b.c1)}; }
  static B200_HD Fp2 sub(const Fp2& a, const Fp2& b) { return {Fp::sub(a.c0, b.c0), Fp::sub(a.c1, b.c1)}; }
  static B200_HD Fp2 dbl(const Fp2& a) { return {Fp::dbl(a.c0), Fp::dbl(a.c1)}; }
  static B200_HD Fp2 neg(const Fp2& a) { return {Fp::neg(a.c0), Fp::neg(a.c1)}; }
  // device: (a0 b0 - a1 b1) + (a0 b1 + a1 b0) u as two fused two-term products (2 x 200 multiply-adds, no
  // field additions: the G2 kernels are limited by the ALU work around their products), both inside ONE out-of-line
  // routine so that the scheduler interleaves the two independent carry chains — the G2 accumulate kernel runs 2 warps
  // per scheduler (252 registers) and was waiting on fixed-latency dependencies (ncu: stall_wait 3.8 per issue,
  // pipe_fmaheavy 85%) with one chain at a time;
  // host: Karatsuba, 3 base-field products
#if defined(__CUDACC__)
  static __host__ __device__ __noinline__ Fp2 mul_pair_call(Fp2 a, Fp2 b) {
    return {Fp::mul_sub(a.c0, b.c0, a.c1, b.c1), Fp::mul_add(a.c0, b.c1, a.c1, b.c0)};
  }
  static __host__ __device__ __noinline__ Fp2 sqr_pair_call(Fp2 a) {
    return {Fp::mul(Fp::add(a.c0, a.c1), Fp::sub(a.c0, a.c1)), Fp::dbl(Fp::mul(a.c0, a.c1))};
  }
#endif
  static B200_HD Fp2 mul(const Fp2& a, const Fp2& b) {
#if defined(__CUDA_ARCH__)
    return mul_pair_call(a, b);
#else
    Fp t0 = Fp::mul_call(a.c0, b.c0);
    Fp t1 = Fp::mul_call(a.c1, b.c1);
    Fp t2 = Fp::mul_call(Fp::add(a.c0, a.c1), Fp::add(b.c0, b.c1));
    return {Fp::sub(t0, t1), Fp::sub(Fp::sub(t2, t0), t1)};
#endif
  }
  static B200_HD Fp2 mul_sub(const Fp2& a, const Fp2& b, const Fp2& c, const Fp2& d) { return sub(mul(a, b), mul(c, d)); }
  // (a0+a1)(a0-a1) + 2 a0 a1 u : 2 products
  static B200_HD Fp2 sqr(const Fp2& a) {
#if defined(__CUDA_ARCH__)
    return sqr_pair_call(a);
#else
    Fp t0 = Fp::mul_call(Fp::add(a.c0, a.c1), Fp::sub(a.c0, a.c1));
    Fp t1 = Fp::mul_call(a.c0, a.c1);
    return {t0, Fp::dbl(t1)};
#endif
  }
  static B200_HD Fp2 inv(const Fp2& a) {
    Fp n = Fp::inv(Fp::add(Fp::sqr(a.c0), Fp::sqr(a.c1)));
    return {Fp::mul(a.c0, n), Fp::neg(Fp::mul(a.c1, n))};
  }
};

template <class F>
struct alignas(16) Affine {
  F x, y;
  B200_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
  static B200_HD Affine inf() { return {F::zero(), F::zero()}; }
};

template <class F>
struct alignas(16) XYZZ {
  F x, y, zz, zzz;
  B200_HD bool is_inf() const { return zz.is_zero(); }
  static B200_HD XYZZ inf() { return {F::zero(), F::zero(), F::zero(), F::zero()}; }
  static B200_HD XYZZ from_affine(const Affine<F>& p) {
    if (p.is_inf()) return inf();
    return {p.x, p.y, F::one(), F::one()};
  }

  // 2*(affine p)   (dbl-2008-s-1 with ZZ=ZZZ=1)
  static B200_HD_NOINLINE XYZZ dbl_affine(const Affine<F>& p) {
    if (p.is_inf() || p.y.is_zero()) return inf();
    F U = F::dbl(p.y);
    F V = F::sqr(U);
    F W = F::mul(U, V);
    F S = F::mul(p.x, V);
    F X2 = F::sqr(p.x);
    F M = F::add(F::dbl(X2), X2);
    XYZZ r;
    r.x = F::sub(F::sqr(M), F::dbl(S));
    r.y = F::mul_sub(M, F::sub(S, r.x), W, p.y);
    r.zz = V;
    r.zzz = W;
    return r;
  }

  B200_HD_NOINLINE void dbl() {
    if (is_inf()) return;
    F U = F::dbl(y);
    F V = F::sqr(U);
    F W = F::mul(U, V);
    F S = F::mul(x, V);
    F X2 = F::sqr(x);
    F M = F::add(F::dbl(X2), X2);
    F X3 = F::sub(F::sqr(M), F::dbl(S));
    F Y3 = F::mul_sub(M, F::sub(S, X3), W, y);
    x = X3;
    y = Y3;
    zz = F::mul(V, zz);
    zzz = F::mul(W, zzz);
  }

  // this += p (affine), all special cases handled  (madd-2008-s)
  B200_HD void madd(const Affine<F>& p) {
    if (p.is_inf()) return;
    if (is_inf()) {
      x = p.x; y = p.y; zz = F::one(); zzz = F::one();
      return;
    }
    F Pq = F::sub(F::mul(p.x, zz), x);
    F Rq = F::sub(F::mul(p.y, zzz), y);
    if (Pq.is_zero()) {
      if (Rq.is_zero()) *this = dbl_affine(p);
      else *this = inf();
      return;
    }
    F PP = F::sqr(Pq);
    F PPP = F::mul(Pq, PP);
    F Q = F::mul(x, PP);
    F X3 = F::sub(F::sub(F::sqr(Rq), PPP), F::dbl(Q));
    y = F::mul_sub(Rq, F::sub(Q, X3), y, PPP);
    x = X3;
    zz = F::mul(zz, PP);
    zzz = F::mul(zzz, PPP);
  }

  // this += q  (add-2008-s).  `add` is one out-of-line routine (operands reached through pointers: fine for the
  // serial tails); `add_inline` is the same code expanded in place for kernels whose whole body is one addition per
  // thread with both operands in registers (the bucket-reduction tree, msm_impl.cuh).
  B200_HD_NOINLINE void add(const XYZZ& q) { add_inline(q); }

  B200_HD void add_inline(const XYZZ& q) {
    if (q.is_inf()) return;
    if (is_inf()) { *this = q; return; }
    F U1 = F::mul(x, q.zz);
    F U2 = F::mul(q.x, zz);
    F S1 = F::mul(y, q.zzz);
    F S2 = F::mul(q.y, zzz);
    F Pq = F::sub(U2, U1);
    F Rq = F::sub(S2, S1);
    if (Pq.is_zero()) {
      if (Rq.is_zero()) dbl();
      else *this = inf();
      return;
    }
    F PP = F::sqr(Pq);
    F PPP = F::mul(Pq, PP);
    F Q = F::mul(U1, PP);
    F X3 = F::sub(F::sub(F::sqr(Rq), PPP), F::dbl(Q));
    y = F::mul_sub(Rq, F::sub(Q, X3), S1, PPP);
    x = X3;
    zz = F::mul(F::mul(zz, q.zz), PP);
    zzz = F::mul(F::mul(zzz, q.zzz), PPP);
  }

  B200_HD void negate() { y = F::neg(y); }

  // this = k * this for a small non-negative k < 2^nbits (double-and-add, MSB first)
  B200_HD void mul_small(uint32_t k, int nbits) {
    XYZZ base = *this;
    *this = inf();
    for (int i = nbits - 1; i >= 0; i--) {
      dbl();
      if ((k >> i) & 1) add(base);
    }
  }

  // normalisation (one inversion): 1/ZZ = (ZZ/ZZZ)^2 because ZZ^3 = ZZZ^2
  B200_HD Affine<F> to_affine() const {
    if (is_inf()) return Affine<F>::inf();
    F zi = F::inv(zzz);
    F izz = F::sqr(F::mul(zz, zi));
    return {F::mul(x, izz), F::mul(y, zi)};
  }
};

using G1Affine = Affine<Fp>;
using G2Affine = Affine<Fp2>;
using G1XYZZ = XYZZ<Fp>;
using G2XYZZ = XYZZ<Fp2>;

}  // namespace b200
