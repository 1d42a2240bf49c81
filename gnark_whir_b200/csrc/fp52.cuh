// Montgomery product on the FP64 pipe: 5 x 52-bit limbs held in doubles, R' = 2^260.
//
// B200 issues DFMA at twice the rate of the 32x32->64 integer multiply-add (IMAD.WIDE) the 8 x 32-bit
// multiplier in field.cuh is built from (csrc/pipe_probe.cu modes 2 and 6 measure both).  A 52 x 52-bit
// product is split exactly into its high and low halves by two round-toward-zero FMAs against the
// constants 2^104 and 2^104 + 2^52:
//     ph = fma_rz(a, b, 2^104)            = 2^104 + floor(a b / 2^52) 2^52     (ulp of [2^104, 2^105) is 2^52)
//     pl = fma_rz(a, b, (2^104 + 2^52) - ph) = 2^52 + (a b mod 2^52)           (exact)
// so the IEEE bit patterns of ph / pl carry the two 52-bit halves in their mantissa fields and are
// accumulated as 64-bit integers; the exponent fields add up to compile-time constants that the column
// accumulators are pre-loaded with (negated).  Word-serial Montgomery reduction on top: after row i the
// low column is cleared by q_i * M with q_i = C_i * (-M^-1) mod 2^52, and its carry moves up.
//
// Per product: 50 limb products x (2 DFMA + 1 DADD) + 10 conversions = 160 FP64-pipe instructions and
// ~100 64-bit integer additions (IADD3 pairs, three-input), against 136 IMAD.WIDE-class instructions at half
// the rate.  Operands: limbs are integers in [0, 2^52) (normalised), values may be any integer below 2^257
// (lazy: no final subtraction); the result is below a b / 2^260 + M < 2^255, limbs normalised.
//
// The same source runs on the host: fma_rz is emulated exactly with 128-bit integers, which is how
// tests/test_lib_cpu.py pins the algorithm without a GPU.
#pragma once
#include <cmath>

#include "field.cuh"

namespace b200 {

struct Fp52Tag {
  static B200_HD constexpr uint64_t mod(int i) {
    constexpr uint64_t m[5] = {0x08c16d87cfd47ull, 0x916871ca8d3c2ull, 0x181585d97816aull, 0xa029b85045b68ull,
                               0x030644e72e131ull};
    return m[i];
  }
  static constexpr uint64_t NP = 0x20782e4866389ull;  // -p^-1 mod 2^52
};
struct Fr52Tag {
  static B200_HD constexpr uint64_t mod(int i) {
    constexpr uint64_t m[5] = {0x1f593f0000001ull, 0x4879b9709143eull, 0x181585d2833e8ull, 0xa029b85045b68ull,
                               0x030644e72e131ull};
    return m[i];
  }
  static constexpr uint64_t NP = 0x1f593efffffffull;  // -r^-1 mod 2^52
};

namespace f52 {
constexpr uint64_t MASK = (1ull << 52) - 1;
constexpr uint64_t BIAS_LO = 0x4330000000000000ull;  // bit pattern of 2^52
constexpr uint64_t BIAS_HI = 0x4670000000000000ull;  // bit pattern of 2^104

#if defined(__CUDA_ARCH__)
B200_D double fma_rz(double a, double b, double c) { return __fma_rz(a, b, c); }
B200_D double sub_exact(double a, double b) { return __dsub_rn(a, b); }
B200_D uint64_t bits(double x) { return (uint64_t)__double_as_longlong(x); }
B200_D double from_bits(uint64_t x) { return __longlong_as_double((long long)x); }
#else
// exact round-toward-zero a*b + c for non-negative integer-valued a, b < 2^53 and integer-valued c
inline double fma_rz(double a, double b, double c) {
  typedef __int128 i128;
  typedef unsigned __int128 u128;
  i128 s = (i128)((u128)(uint64_t)a * (u128)(uint64_t)b) + (i128)c;
  const bool neg = s < 0;
  u128 m = neg ? (u128)(-s) : (u128)s;
  int msb = 127;
  while (msb > 0 && !((m >> msb) & 1)) msb--;
  const int sh = msb > 52 ? msb - 52 : 0;
  double r = std::ldexp((double)(uint64_t)(m >> sh), sh);
  return neg ? -r : r;
}
inline double sub_exact(double a, double b) { return a - b; }
inline uint64_t bits(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
inline double from_bits(uint64_t x) { double d; memcpy(&d, &x, 8); return d; }
#endif

// integer in [0, 2^52) -> double with that value
B200_HD double to_double(uint64_t v) { return sub_exact(from_bits(v | BIAS_LO), 4503599627370496.0); }

// pairs (i, j), 0 <= i, j <= 4, with i + j == s
B200_HD constexpr int ndiag(int s) { return (s < 0 || s > 8) ? 0 : (s <= 4 ? s + 1 : 9 - s); }
}  // namespace f52

template <class T>
struct Field52 {
  double l[5];

  // a * b * 2^-260 mod M (lazy: below a b / 2^260 + M)
  static B200_HD Field52 mul(const Field52& a, const Field52& b) {
    using namespace f52;
    const double C1 = 20282409603651670423947251286016.0;          // 2^104
    const double C2 = 20282409603651674927546878656512.0;          // 2^104 + 2^52
    uint64_t C[10];
#pragma unroll
    for (int g = 0; g < 10; g++)  // minus the exponent fields column g is going to receive (a*b and q*M terms)
      C[g] = 0ull - (2ull * (uint64_t)ndiag(g) * BIAS_LO + 2ull * (uint64_t)ndiag(g - 1) * BIAS_HI);
#pragma unroll
    for (int i = 0; i < 5; i++) {
#pragma unroll
      for (int j = 0; j < 5; j++) {
        const double ph = fma_rz(a.l[i], b.l[j], C1);
        const double pl = fma_rz(a.l[i], b.l[j], sub_exact(C2, ph));
        C[i + j] += bits(pl);
        C[i + j + 1] += bits(ph);
      }
      const uint64_t q = (C[i] * T::NP) & MASK;
      const double qd = to_double(q);
#pragma unroll
      for (int j = 0; j < 5; j++) {
        const double mj = (double)T::mod(j);
        const double ph = fma_rz(qd, mj, C1);
        const double pl = fma_rz(qd, mj, sub_exact(C2, ph));
        C[i + j] += bits(pl);
        C[i + j + 1] += bits(ph);
      }
      C[i + 1] += C[i] >> 52;  // C[i] is now an exact multiple of 2^52
    }
    Field52 r;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      r.l[k] = to_double(C[5 + k] & MASK);
      C[6 + k] += C[5 + k] >> 52;
    }
    r.l[4] = to_double(C[9]);
    return r;
  }

  // 256-bit integer (8 x 32-bit limbs) <-> 5 x 52-bit limbs
  static B200_HD Field52 from_words(const uint32_t* w) {
    using namespace f52;
    uint64_t v[4];
#pragma unroll
    for (int i = 0; i < 4; i++) v[i] = (uint64_t)w[2 * i] | ((uint64_t)w[2 * i + 1] << 32);
    Field52 r;
    r.l[0] = to_double(v[0] & MASK);
    r.l[1] = to_double(((v[0] >> 52) | (v[1] << 12)) & MASK);
    r.l[2] = to_double(((v[1] >> 40) | (v[2] << 24)) & MASK);
    r.l[3] = to_double(((v[2] >> 28) | (v[3] << 36)) & MASK);
    r.l[4] = to_double(v[3] >> 16);
    return r;
  }
  // value must be below 2^256
  B200_HD void to_words(uint32_t* w) const {
    uint64_t x[5];
#pragma unroll
    for (int i = 0; i < 5; i++) x[i] = (uint64_t)l[i];
    uint64_t v[4];
    v[0] = x[0] | (x[1] << 52);
    v[1] = (x[1] >> 12) | (x[2] << 40);
    v[2] = (x[2] >> 24) | (x[3] << 28);
    v[3] = (x[3] >> 36) | (x[4] << 16);
#pragma unroll
    for (int i = 0; i < 4; i++) { w[2 * i] = (uint32_t)v[i]; w[2 * i + 1] = (uint32_t)(v[i] >> 32); }
  }
};

using Fp52 = Field52<Fp52Tag>;
using Fr52 = Field52<Fr52Tag>;

}  // namespace b200
