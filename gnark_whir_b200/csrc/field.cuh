// 254-bit prime-field arithmetic in Montgomery form (R = 2^256), 8 x 32-bit limbs.
//
// Memory layout is gnark-crypto's fp.Element / fr.Element ([4]uint64 little-endian limbs,
// Montgomery form) reinterpreted as 8 little-endian uint32 — byte-identical on x86-64 and
// on the GPU, so Go slices can be handed to the C-ABI untouched (the reference
// shows the 4xu64 LE limb convention at main.go:19-21 and typeConverters.go:26-44).
//
// mul(): word-serial Montgomery product with interleaved reduction, organised as two
// carry chains over 64-bit-aligned column pairs ("even"/"odd" accumulators) so that every
// 32x32 partial product is one mad.lo.cc/madc.hi.cc pair (one IMAD.WIDE on sm_100a) and no
// partial product ever needs a separate carry fix-up: 2*8*8 + 8 = 136 multiply-adds.
//
// Lineage: the even/odd word-serial scheme and its helper structure (mul_n, cmad_n, madc_n_rshift,
// mad_n_redc, the final "even[i] + odd[i+1]" hand-over) are the published technique of Supranational's
// sppark (mont_t.cuh), which ICICLE's field code also follows; it is the natural mapping of a Montgomery
// product onto mad.lo.cc / madc.hi.cc and is restated here, not copied from the reference (which contains
// no GPU code).  Original to this file: the dedicated square (sqr_ptx), the fused two-term product
// (dot2_ptx), and the bit-exact host emulation the CPU test-suite pins them with.  fp52.cuh holds the
// measured alternative on the FP64 pipe (no faster on B200: DESIGN.md §2).
#pragma once
#include "bn254_constants.h"
#include "ptx.cuh"

namespace b200 {

struct FpTag {
  static constexpr uint32_t INV = FpParams::INV;
  static constexpr uint64_t INV64 = 0x87d20782e4866389ull;  // -p^-1 mod 2^64
  static B200_HD constexpr uint32_t mod(int i) { constexpr uint32_t m[8] = B200_FP_MOD; return m[i]; }
  static B200_HD constexpr uint32_t one(int i) { constexpr uint32_t m[8] = B200_FP_ONE; return m[i]; }
  static B200_HD constexpr uint32_t r2(int i) { constexpr uint32_t m[8] = B200_FP_R2; return m[i]; }
};
struct FrTag {
  static constexpr uint32_t INV = FrParams::INV;
  static constexpr uint64_t INV64 = 0xc2e1f593efffffffull;  // -r^-1 mod 2^64
  static B200_HD constexpr uint32_t mod(int i) { constexpr uint32_t m[8] = B200_FR_MOD; return m[i]; }
  static B200_HD constexpr uint32_t one(int i) { constexpr uint32_t m[8] = B200_FR_ONE; return m[i]; }
  static B200_HD constexpr uint32_t r2(int i) { constexpr uint32_t m[8] = B200_FR_R2; return m[i]; }
};

template <class T>
struct alignas(16) Field {
  static constexpr int N = 8;
  uint32_t l[N];

  static B200_HD Field zero() { Field r; for (int i = 0; i < N; i++) r.l[i] = 0; return r; }
  static B200_HD Field one() { Field r; for (int i = 0; i < N; i++) r.l[i] = T::one(i); return r; }
  static B200_HD Field modulus() { Field r; for (int i = 0; i < N; i++) r.l[i] = T::mod(i); return r; }
  static B200_HD Field rsquared() { Field r; for (int i = 0; i < N; i++) r.l[i] = T::r2(i); return r; }

  B200_HD bool is_zero() const {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < N; i++) o |= l[i];
    return o == 0;
  }
  B200_HD bool operator==(const Field& b) const {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < N; i++) o |= l[i] ^ b.l[i];
    return o == 0;
  }
  B200_HD bool operator!=(const Field& b) const { return !(*this == b); }

  // r = a - p if a >= p  (a < 2p)
  static B200_HD void reduce_once(Field& a) {
    uint32_t t[N];
    t[0] = ptx::sub_cc(a.l[0], T::mod(0));
#pragma unroll
    for (int i = 1; i < N; i++) t[i] = ptx::subc_cc(a.l[i], T::mod(i));
    uint32_t borrow = ptx::subc(0, 0);  // 0 or 0xffffffff
#pragma unroll
    for (int i = 0; i < N; i++) a.l[i] = borrow ? a.l[i] : t[i];
  }

  static B200_HD Field add(const Field& a, const Field& b) {
    Field r;
    r.l[0] = ptx::add_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(a.l[i], b.l[i]);
    r.l[N - 1] = ptx::addc(a.l[N - 1], b.l[N - 1]);  // p < 2^254: no carry out
    reduce_once(r);
    return r;
  }
  static B200_HD Field dbl(const Field& a) { return add(a, a); }

  static B200_HD Field sub(const Field& a, const Field& b) {
    Field r;
    r.l[0] = ptx::sub_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < N; i++) r.l[i] = ptx::subc_cc(a.l[i], b.l[i]);
    uint32_t borrow = ptx::subc(0, 0);
    r.l[0] = ptx::add_cc(r.l[0], T::mod(0) & borrow);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(r.l[i], T::mod(i) & borrow);
    r.l[N - 1] = ptx::addc(r.l[N - 1], T::mod(N - 1) & borrow);
    return r;
  }
  static B200_HD Field neg(const Field& a) {
    if (a.is_zero()) return a;
    Field r;
    r.l[0] = ptx::sub_cc(T::mod(0), a.l[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = ptx::subc_cc(T::mod(i), a.l[i]);
    r.l[N - 1] = ptx::subc(T::mod(N - 1), a.l[N - 1]);
    return r;
  }

  // ---- Montgomery product ------------------------------------------------------------
  // acc[j],acc[j+1] = a[j]*bi for even j (independent wide products)
  static B200_HD void mul_n(uint32_t* acc, const uint32_t* a, uint32_t bi) {
#pragma unroll
    for (int j = 0; j < N; j += 2) {
      acc[j] = ptx::mul_lo(a[j], bi);
      acc[j + 1] = ptx::mul_hi(a[j], bi);
    }
  }
  // acc[j],acc[j+1] += a[j]*bi for even j, one carry chain; carry-out left in CC
  static B200_HD void cmad_n(uint32_t* acc, const uint32_t* a, uint32_t bi) {
    acc[0] = ptx::mad_lo_cc(a[0], bi, acc[0]);
    acc[1] = ptx::madc_hi_cc(a[0], bi, acc[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      acc[j] = ptx::madc_lo_cc(a[j], bi, acc[j]);
      acc[j + 1] = ptx::madc_hi_cc(a[j], bi, acc[j + 1]);
    }
  }
  // same chain against the compile-time modulus limbs (offset 0 or 1)
  template <int OFF>
  static B200_HD void cmad_mod(uint32_t* acc, uint32_t mi) {
    acc[0] = ptx::mad_lo_cc(T::mod(OFF), mi, acc[0]);
    acc[1] = ptx::madc_hi_cc(T::mod(OFF), mi, acc[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      acc[j] = ptx::madc_lo_cc(T::mod(OFF + j), mi, acc[j]);
      acc[j + 1] = ptx::madc_hi_cc(T::mod(OFF + j), mi, acc[j + 1]);
    }
  }
  // acc[j],acc[j+1] = a[j]*bi + acc[j+2],acc[j+3] (+ carry-in from CC); top pair gets 0
  static B200_HD void madc_n_rshift(uint32_t* acc, const uint32_t* a, uint32_t bi) {
#pragma unroll
    for (int j = 0; j < N - 2; j += 2) {
      acc[j] = ptx::madc_lo_cc(a[j], bi, acc[j + 2]);
      acc[j + 1] = ptx::madc_hi_cc(a[j], bi, acc[j + 3]);
    }
    acc[N - 2] = ptx::madc_lo_cc(a[N - 2], bi, 0);
    acc[N - 1] = ptx::madc_hi(a[N - 2], bi, 0);
  }
  // One word of b: lo holds columns 0,1,..,7 (weights 2^(32k)), hi holds columns 1..8
  // (weights 2^(32(k+1))) from the PREVIOUS word, i.e. one limb ahead; after the call the
  // roles swap (the implicit divide-by-2^32).
  template <bool FIRST>
  static B200_HD void mad_n_redc(uint32_t* lo, uint32_t* hi, const uint32_t* a, uint32_t bi) {
    if (FIRST) {
      mul_n(hi, a + 1, bi);
      mul_n(lo, a, bi);
    } else {
      lo[0] = ptx::add_cc(lo[0], hi[1]);
      madc_n_rshift(hi, a + 1, bi);
      cmad_n(lo, a, bi);
      hi[N - 1] = ptx::addc(hi[N - 1], 0);
    }
    uint32_t mi = ptx::mul_lo(lo[0], T::INV);
    cmad_mod<1>(hi, mi);
    cmad_mod<0>(lo, mi);
    hi[N - 1] = ptx::addc(hi[N - 1], 0);
  }

  // The device multiplier.  Compiled for the host too (through ptx.cuh's emulation) so the
  // CPU test-suite can pin the exact instruction sequence the GPU executes.
  static B200_HD Field mul_ptx(const Field& a, const Field& b) {
    uint32_t even[N], odd[N];
    mad_n_redc<true>(even, odd, a.l, b.l[0]);
    mad_n_redc<false>(odd, even, a.l, b.l[1]);
#pragma unroll
    for (int i = 2; i < N; i += 2) {
      mad_n_redc<false>(even, odd, a.l, b.l[i]);
      mad_n_redc<false>(odd, even, a.l, b.l[i + 1]);
    }
    Field r;
    r.l[0] = ptx::add_cc(even[0], odd[1]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(even[i], odd[i + 1]);
    r.l[N - 1] = ptx::addc(even[N - 1], 0);
    reduce_once(r);
    return r;
  }
  // ---- fused a*b + c*d ------------------------------------------------------------------
  // One word-serial pass accumulates both products row by row and reduces once per word:
  // 64 + 64 + 72 = 200 multiply-adds instead of 2 x 136.  Bounds (a, b, c, d < p < 2^254): a row adds less
  // than 3 * 2^286 to an accumulator below 3p, so the even/odd accumulator pair (288 bits) cannot overflow,
  // and the result is below p (1 + 2p / 2^256) < 2p: one conditional subtraction, as for the product.
  template <bool FIRST>
  static B200_HD void mad2_n_redc(uint32_t* lo, uint32_t* hi, const uint32_t* a, uint32_t bi, const uint32_t* c,
                                  uint32_t di) {
    if (FIRST) {
      mul_n(hi, a + 1, bi);
      mul_n(lo, a, bi);
    } else {
      lo[0] = ptx::add_cc(lo[0], hi[1]);
      madc_n_rshift(hi, a + 1, bi);
      cmad_n(lo, a, bi);
      hi[N - 1] = ptx::addc(hi[N - 1], 0);
    }
    cmad_n(hi, c + 1, di);  // no carry out: the odd accumulator stays below 2^256 (see bounds)
    cmad_n(lo, c, di);
    hi[N - 1] = ptx::addc(hi[N - 1], 0);
    uint32_t mi = ptx::mul_lo(lo[0], T::INV);
    cmad_mod<1>(hi, mi);
    cmad_mod<0>(lo, mi);
    hi[N - 1] = ptx::addc(hi[N - 1], 0);
  }
  static B200_HD Field dot2_ptx(const Field& a, const Field& b, const Field& c, const Field& d) {
    uint32_t even[N], odd[N];
    mad2_n_redc<true>(even, odd, a.l, b.l[0], c.l, d.l[0]);
    mad2_n_redc<false>(odd, even, a.l, b.l[1], c.l, d.l[1]);
#pragma unroll
    for (int i = 2; i < N; i += 2) {
      mad2_n_redc<false>(even, odd, a.l, b.l[i], c.l, d.l[i]);
      mad2_n_redc<false>(odd, even, a.l, b.l[i + 1], c.l, d.l[i + 1]);
    }
    Field r;
    r.l[0] = ptx::add_cc(even[0], odd[1]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(even[i], odd[i + 1]);
    r.l[N - 1] = ptx::addc(even[N - 1], 0);
    reduce_once(r);
    return r;
  }

  // ---- Montgomery square ---------------------------------------------------------------
  // a^2 = sum_i a_i 2^(32i) * V_i with V_i = a_i 2^(32i) + 2 * (a with limbs 0..i cleared): row i of the
  // word-serial product above multiplies the word a_i by the limbs of V_i, whose limbs below i are zero, so
  // those partial products become plain carry propagation (IADD3.X on the otherwise idle ALU pipe):
  // 36 + 72 = 108 wide multiply-adds instead of 136.  Limbs of V_i: j = i: a_i; j = i+1: a_j << 1;
  // j > i+1: (a_j << 1) | (a_(j-1) >> 31)  (a < 2^254, so nothing is shifted out at the top).
  // Bounds: V_i < 2^255, hence every intermediate accumulator stays below 2^255 + 2^254 < 2^256 and the
  // result below 2p, as for the product.
  template <bool Z> static B200_HD uint32_t zmad_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
    return Z ? ptx::add_cc(c, 0) : ptx::mad_lo_cc(a, b, c);
  }
  template <bool Z> static B200_HD uint32_t zmadc_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
    return Z ? ptx::addc_cc(c, 0) : ptx::madc_lo_cc(a, b, c);
  }
  template <bool Z> static B200_HD uint32_t zmadc_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
    return Z ? ptx::addc_cc(c, 0) : ptx::madc_hi_cc(a, b, c);
  }
  template <bool Z> static B200_HD uint32_t zmadc_hi(uint32_t a, uint32_t b, uint32_t c) {
    return Z ? ptx::addc(c, 0) : ptx::madc_hi(a, b, c);
  }
  // cmad_n / madc_n_rshift over an operand whose limbs below S are zero (OFF = 0: even limbs, 1: odd limbs)
  template <int S, int OFF>
  static B200_HD void cmad_n_from(uint32_t* acc, const uint32_t* v, uint32_t bi) {
    acc[0] = zmad_lo_cc<(OFF < S)>(v[OFF], bi, acc[0]);
    acc[1] = zmadc_hi_cc<(OFF < S)>(v[OFF], bi, acc[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      acc[j] = (j + OFF < S) ? ptx::addc_cc(acc[j], 0) : ptx::madc_lo_cc(v[j + OFF], bi, acc[j]);
      acc[j + 1] = (j + OFF < S) ? ptx::addc_cc(acc[j + 1], 0) : ptx::madc_hi_cc(v[j + OFF], bi, acc[j + 1]);
    }
  }
  template <int S, int OFF>
  static B200_HD void madc_n_rshift_from(uint32_t* acc, const uint32_t* v, uint32_t bi) {
#pragma unroll
    for (int j = 0; j < N - 2; j += 2) {
      acc[j] = (j + OFF < S) ? ptx::addc_cc(acc[j + 2], 0) : ptx::madc_lo_cc(v[j + OFF], bi, acc[j + 2]);
      acc[j + 1] = (j + OFF < S) ? ptx::addc_cc(acc[j + 3], 0) : ptx::madc_hi_cc(v[j + OFF], bi, acc[j + 3]);
    }
    acc[N - 2] = zmadc_lo_cc<(N - 2 + OFF < S)>(v[N - 2 + OFF], bi, 0);
    acc[N - 1] = zmadc_hi<(N - 2 + OFF < S)>(v[N - 2 + OFF], bi, 0);
  }
  // row I of the square: operand limbs v[j], j >= I (lower limbs are zero), word bi = a_I
  template <int I>
  static B200_HD void sqr_row_redc(uint32_t* lo, uint32_t* hi, const uint32_t* v, uint32_t bi) {
    if (I == 0) {
      mul_n(hi, v + 1, bi);
      mul_n(lo, v, bi);
    } else {
      lo[0] = ptx::add_cc(lo[0], hi[1]);
      madc_n_rshift_from<I, 1>(hi, v, bi);
      cmad_n_from<I, 0>(lo, v, bi);
      hi[N - 1] = ptx::addc(hi[N - 1], 0);
    }
    uint32_t mi = ptx::mul_lo(lo[0], T::INV);
    cmad_mod<1>(hi, mi);
    cmad_mod<0>(lo, mi);
    hi[N - 1] = ptx::addc(hi[N - 1], 0);
  }
  template <int I>
  static B200_HD void sqr_operand(uint32_t* v, const Field& a) {  // limbs of V_I at positions >= I
#pragma unroll
    for (int j = 0; j < N; j++) {
      if (j < I) v[j] = 0;
      else if (j == I) v[j] = a.l[j];
      else if (j == I + 1) v[j] = a.l[j] << 1;
      else v[j] = (a.l[j] << 1) | (a.l[j - 1] >> 31);
    }
  }
  static B200_HD Field sqr_ptx(const Field& a) {
    uint32_t even[N], odd[N], v[N];
    sqr_operand<0>(v, a); sqr_row_redc<0>(even, odd, v, a.l[0]);
    sqr_operand<1>(v, a); sqr_row_redc<1>(odd, even, v, a.l[1]);
    sqr_operand<2>(v, a); sqr_row_redc<2>(even, odd, v, a.l[2]);
    sqr_operand<3>(v, a); sqr_row_redc<3>(odd, even, v, a.l[3]);
    sqr_operand<4>(v, a); sqr_row_redc<4>(even, odd, v, a.l[4]);
    sqr_operand<5>(v, a); sqr_row_redc<5>(odd, even, v, a.l[5]);
    sqr_operand<6>(v, a); sqr_row_redc<6>(even, odd, v, a.l[6]);
    sqr_operand<7>(v, a); sqr_row_redc<7>(odd, even, v, a.l[7]);
    Field r;
    r.l[0] = ptx::add_cc(even[0], odd[1]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(even[i], odd[i + 1]);
    r.l[N - 1] = ptx::addc(even[N - 1], 0);
    reduce_once(r);
    return r;
  }
#if !defined(__CUDA_ARCH__)
  // Host multiplier: 4 x 64-bit CIOS on unsigned __int128 (the host finishes every MSM with a
  // Horner pass and the proof with a handful of scalar multiplications).
  static inline Field mul_host64(const Field& a, const Field& b) {
    typedef unsigned __int128 u128;
    uint64_t A[4], B[4], M[4], t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
      A[i] = (uint64_t)a.l[2 * i] | ((uint64_t)a.l[2 * i + 1] << 32);
      B[i] = (uint64_t)b.l[2 * i] | ((uint64_t)b.l[2 * i + 1] << 32);
      M[i] = (uint64_t)T::mod(2 * i) | ((uint64_t)T::mod(2 * i + 1) << 32);
    }
    for (int i = 0; i < 4; i++) {
      uint64_t carry = 0;
      for (int j = 0; j < 4; j++) {
        u128 p = (u128)A[j] * B[i] + t[j] + carry;
        t[j] = (uint64_t)p;
        carry = (uint64_t)(p >> 64);
      }
      u128 q = (u128)t[4] + carry;
      t[4] = (uint64_t)q;
      t[5] = (uint64_t)(q >> 64);
      uint64_t m = t[0] * T::INV64;
      u128 p = (u128)m * M[0] + t[0];
      carry = (uint64_t)(p >> 64);
      for (int j = 1; j < 4; j++) {
        p = (u128)m * M[j] + t[j] + carry;
        t[j - 1] = (uint64_t)p;
        carry = (uint64_t)(p >> 64);
      }
      q = (u128)t[4] + carry;
      t[3] = (uint64_t)q;
      t[4] = t[5] + (uint64_t)(q >> 64);
    }
    // t < 2m: subtract m once if needed
    uint64_t d[4], borrow = 0;
    for (int i = 0; i < 4; i++) {
      u128 x = (u128)t[i] - M[i] - borrow;
      d[i] = (uint64_t)x;
      borrow = (uint64_t)(x >> 64) & 1;
    }
    bool ge = t[4] != 0 || borrow == 0;
    Field r;
    for (int i = 0; i < 4; i++) {
      uint64_t v = ge ? d[i] : t[i];
      r.l[2 * i] = (uint32_t)v;
      r.l[2 * i + 1] = (uint32_t)(v >> 32);
    }
    return r;
  }
#endif
  static B200_HD Field mul(const Field& a, const Field& b) {
#if defined(__CUDA_ARCH__)
    return mul_ptx(a, b);
#else
    return mul_host64(a, b);
#endif
  }
  static B200_HD Field sqr(const Field& a) {
#if defined(__CUDA_ARCH__)
    return sqr_ptx(a);
#else
    return mul(a, a);
#endif
  }

  // a*b - c*d and a*b + c*d with one reduction on the device (dot2_ptx); two products on the host
  static B200_HD Field mul_add(const Field& a, const Field& b, const Field& c, const Field& d) {
#if defined(__CUDA_ARCH__)
    return dot2_ptx(a, b, c, d);
#else
    return add(mul(a, b), mul(c, d));
#endif
  }
  static B200_HD Field mul_sub(const Field& a, const Field& b, const Field& c, const Field& d) {
#if defined(__CUDA_ARCH__)
    return dot2_ptx(a, b, neg(c), d);
#else
    return sub(mul(a, b), mul(c, d));
#endif
  }
#if defined(__CUDACC__)
  static __host__ __device__ __noinline__ Field mul_add_call(Field a, Field b, Field c, Field d) { return mul_add(a, b, c, d); }
  static __host__ __device__ __noinline__ Field mul_sub_call(Field a, Field b, Field c, Field d) { return mul_sub(a, b, c, d); }
#else
  static inline Field mul_add_call(Field a, Field b, Field c, Field d) { return mul_add(a, b, c, d); }
  static inline Field mul_sub_call(Field a, Field b, Field c, Field d) { return mul_sub(a, b, c, d); }
#endif

  // Out-of-line product, operands and result by value (registers): one copy of the multiplier
  // shared by every call site.  Used by Fp2 so that the G2 kernels stay small and keep their
  // register count (and therefore occupancy) close to the G1 kernels'.
#if defined(__CUDACC__)
  static __host__ __device__ __noinline__ Field mul_call(Field a, Field b) { return mul(a, b); }
#else
  static inline Field mul_call(Field a, Field b) { return mul(a, b); }
#endif

  // x*R^-1 (Montgomery -> canonical) and x*R (canonical -> Montgomery)
  static B200_HD Field from_mont(const Field& a) {
    Field o = zero();
    o.l[0] = 1;
    return mul(a, o);
  }
  static B200_HD Field to_mont(const Field& a) { return mul(a, rsquared()); }

  // a^e for a canonical (non-Montgomery) exponent given as 8 limbs
  static B200_HD_NOINLINE Field pow(const Field& a, const uint32_t* e) {
    Field r = one();
    for (int i = N * 32 - 1; i >= 0; i--) {
      r = sqr(r);
      if ((e[i >> 5] >> (i & 31)) & 1) r = mul(r, a);
    }
    return r;
  }
  // Fermat inverse (0 -> 0)
  static B200_HD Field inv(const Field& a) {
    uint32_t e[N];
    for (int i = 0; i < N; i++) e[i] = T::mod(i);
    e[0] -= 2;  // both moduli end in ...47 / ...01: no borrow for Fp; Fr: 0xf0000001-2 ok
    return pow(a, e);
  }
};

using Fp = Field<FpTag>;
using Fr = Field<FrTag>;

}  // namespace b200
