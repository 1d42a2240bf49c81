// Test harness (NOT part of libb200g16): compiles gnark_whir_b200/csrc/pairing.cuh for the host with
// g++ so the CPU test-suite can pin the exact tower / Miller-loop / final-exponentiation code the GPU
// kernel k_pairing_products executes against the oracle's independent pairing (oracle/bn254.py).
#include <cstring>

#include "pairing.cuh"

using namespace b200;

extern "C" {

// prod e(P_i, Q_i), reduced; out in gnark's E12 layout.  Returns 1 when the product is one.
int host_pair(const uint64_t* g1, const uint64_t* g2, int n, int with_cofactor, uint64_t out[48]) {
  Fp12 f = f12_one();
  for (int i = 0; i < n; i++) {
    Affine<Fp> p;
    Affine<Fp2> q;
    memcpy(&p, g1 + 8 * i, sizeof(p));
    memcpy(&q, g2 + 16 * i, sizeof(q));
    f = f12_mul(f, miller_loop(p, q));
  }
  Fp12 g = final_exponentiation(f, with_cofactor != 0);
  to_gnark_layout(g, out);
  return f12_is_one(g) ? 1 : 0;
}

int host_g1_on_curve(const uint64_t* g1) {
  Affine<Fp> p;
  memcpy(&p, g1, sizeof(p));
  return g1_on_curve(p) ? 1 : 0;
}

int host_g2_in_subgroup(const uint64_t* g2) {
  Affine<Fp2> q;
  memcpy(&q, g2, sizeof(q));
  return g2_in_subgroup(q) ? 1 : 0;
}

// a * b and 1/a in Fp12 (flat coefficient order c0..c5, each Fp2 = 8 u64 Montgomery)
void host_f12_mul(const uint64_t* a, const uint64_t* b, uint64_t* out) {
  Fp12 x, y;
  memcpy(&x, a, sizeof(x));
  memcpy(&y, b, sizeof(y));
  Fp12 r = f12_mul(x, y);
  memcpy(out, &r, sizeof(r));
}
void host_f12_inv(const uint64_t* a, uint64_t* out) {
  Fp12 x;
  memcpy(&x, a, sizeof(x));
  Fp12 r = f12_inv(x);
  memcpy(out, &r, sizeof(r));
}
void host_f12_frob(const uint64_t* a, int k, uint64_t* out) {
  Fp12 x;
  memcpy(&x, a, sizeof(x));
  Fp12 r = k == 1 ? f12_frob1(x) : (k == 2 ? f12_frob2(x) : f12_frob3(x));
  memcpy(out, &r, sizeof(r));
}
void host_f12_sqr(const uint64_t* a, uint64_t* out) {
  Fp12 x;
  memcpy(&x, a, sizeof(x));
  Fp12 r = f12_sqr(x);
  memcpy(out, &r, sizeof(r));
}
void host_f12_frob2(const uint64_t* a, uint64_t* out) {
  Fp12 x;
  memcpy(&x, a, sizeof(x));
  Fp12 r = f12_frob2(x);
  memcpy(out, &r, sizeof(r));
}
}
