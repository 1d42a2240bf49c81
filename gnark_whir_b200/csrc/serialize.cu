// gnark-crypto point encodings on the GPU (batch), for the gnark wire formats.
//
// Replaces gnark-crypto ecc/bn254/marshal.go G1Affine.Bytes / RawBytes / SetBytes and the G2Affine
// counterparts, as the curve Encoder / Decoder apply them element by element to the point slices of
// groth16_bn254.ProvingKey / VerifyingKey / Proof (gnark backend/groth16/bn254/marshal.go; the proof and
// keys of the reference's flow, mt.go:448,496).  Reading a compressed key is one square root per point —
// 2^24 of them for this circuit class — which is the data-parallel part moved here.
//
// Encodings (big-endian field elements; two flag bits on top of the first byte):
//   0b00 uncompressed X||Y, 0b01 infinity, 0b10 compressed / Y lexicographically smallest, 0b11 / largest.
//   G2: X = X.A1 || X.A0; "largest" compares A1 first, then A0.
// Records are fixed-size: all compressed (32 / 64 B) or all raw (64 / 128 B).
#include "common.cuh"
#include "pairing.cuh"

namespace b200 {

constexpr uint8_t M_MASK = 0xC0, M_UNCOMPRESSED = 0x00, M_INFINITY = 0x40, M_SMALLEST = 0x80, M_LARGEST = 0xC0;

// 32 big-endian bytes (top two bits of byte 0 masked off) -> 8 little-endian limbs, plain integer
__device__ __forceinline__ Fp load_be(const uint8_t* p, bool mask_flags) {
  Fp r;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint8_t* q = p + 4 * (7 - i);
    uint32_t b0 = q[0];
    if (mask_flags && i == 7) b0 &= 0x3Fu;
    r.l[i] = (b0 << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
  }
  return r;
}
__device__ __forceinline__ void store_be(uint8_t* p, const Fp& v) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint8_t* q = p + 4 * (7 - i);
    q[0] = (uint8_t)(v.l[i] >> 24); q[1] = (uint8_t)(v.l[i] >> 16); q[2] = (uint8_t)(v.l[i] >> 8); q[3] = (uint8_t)v.l[i];
  }
}
// a < b on plain integers
__device__ __forceinline__ bool lt(const Fp& a, const Fp& b) {
  for (int i = 7; i >= 0; i--) {
    if (a.l[i] != b.l[i]) return a.l[i] < b.l[i];
  }
  return false;
}
__device__ __forceinline__ bool reduced(const Fp& a) { return lt(a, Fp::modulus()); }
__device__ __forceinline__ bool lex_largest(const Fp& y_mont) {  // y > (p-1)/2
  constexpr uint32_t h[8] = B200_FP_HALF;
  return lt(fp_from_words(h), Fp::from_mont(y_mont));
}
__device__ __forceinline__ bool lex_largest(const Fp2& y) {
  return y.c1.is_zero() ? lex_largest(y.c0) : lex_largest(y.c1);
}
// square root in Fp (p = 3 mod 4); ok = false when a is not a square
__device__ __noinline__ Fp fp_sqrt(const Fp& a, bool* ok) {
  constexpr uint32_t e[8] = B200_FP_SQRT_EXP;
  Fp r = Fp::pow(a, e);
  *ok = Fp::sqr(r) == a;
  return r;
}
// square root in Fp2 = Fp[u]/(u^2+1) by the complex method
__device__ __noinline__ Fp2 fp2_sqrt(const Fp2& a, bool* ok) {
  bool g;
  if (a.c1.is_zero()) {
    Fp r = fp_sqrt(a.c0, &g);
    if (g) { *ok = true; return {r, Fp::zero()}; }
    r = fp_sqrt(Fp::neg(a.c0), &g);
    *ok = g;
    return {Fp::zero(), r};
  }
  Fp s = fp_sqrt(Fp::add(Fp::sqr(a.c0), Fp::sqr(a.c1)), &g);
  if (!g) { *ok = false; return Fp2::zero(); }
  constexpr uint32_t i2[8] = B200_FP_INV2;
  const Fp inv2 = fp_from_words(i2);
  Fp x0 = fp_sqrt(Fp::mul(Fp::add(a.c0, s), inv2), &g);
  if (!g) x0 = fp_sqrt(Fp::mul(Fp::sub(a.c0, s), inv2), &g);
  if (!g) { *ok = false; return Fp2::zero(); }
  Fp x1 = Fp::mul(a.c1, Fp::inv(Fp::dbl(x0)));
  Fp2 r = {x0, x1};
  *ok = Fp2::sqr(r) == a;
  return r;
}

__device__ __forceinline__ bool all_zero(const uint8_t* p, int n, bool skip_flag_bits) {
  uint32_t o = skip_flag_bits ? (p[0] & 0x3Fu) : p[0];
  for (int i = 1; i < n; i++) o |= p[i];
  return o == 0;
}

// ---- G1
__global__ void __launch_bounds__(128) k_g1_decode(const uint8_t* __restrict__ in, uint32_t n, int raw,
                                                    G1Affine* __restrict__ out, uint8_t* __restrict__ ok_out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int rec = raw ? 64 : 32;
  const uint8_t* p = in + (size_t)i * rec;
  const uint8_t flag = p[0] & M_MASK;
  G1Affine r = G1Affine::inf();
  bool ok = false;
  if (raw) {
    // bn254 has two flag bits only: the uncompressed point at infinity is flag 0b00 followed by zeros (X = Y = 0);
    // 0b01 marks the COMPRESSED infinity (a compressed-size record) and has no place in a raw batch
    if (flag == M_UNCOMPRESSED) {
      Fp x = load_be(p, false), y = load_be(p + 32, false);
      if (x.is_zero() && y.is_zero()) {
        ok = true;
      } else if (reduced(x) && reduced(y)) {
        r = {Fp::to_mont(x), Fp::to_mont(y)};
        ok = g1_on_curve(r);
      }
    }
  } else if (flag == M_INFINITY) {
    ok = all_zero(p, rec, true);
  } else if (flag != M_UNCOMPRESSED) {
    Fp x = load_be(p, true);
    if (reduced(x)) {
      Fp xm = Fp::to_mont(x);
      constexpr uint32_t three[8] = B200_FP_THREE;
      Fp y = fp_sqrt(Fp::add(Fp::mul(Fp::sqr(xm), xm), fp_from_words(three)), &ok);
      if (ok) {
        if (lex_largest(y) != (flag == M_LARGEST)) y = Fp::neg(y);
        r = {xm, y};
      }
    }
  }
  out[i] = ok ? r : G1Affine::inf();
  ok_out[i] = ok ? 1 : 0;
}

__global__ void __launch_bounds__(128) k_g1_encode(const G1Affine* __restrict__ pts, uint32_t n, int raw,
                                                    uint8_t* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int rec = raw ? 64 : 32;
  uint8_t* p = out + (size_t)i * rec;
  const G1Affine a = pts[i];
  if (a.is_inf()) {
    for (int k = 0; k < rec; k++) p[k] = 0;
    if (!raw) p[0] = M_INFINITY;   // RawBytes() of infinity is all zero (flag 0b00)
    return;
  }
  store_be(p, Fp::from_mont(a.x));
  if (raw) store_be(p + 32, Fp::from_mont(a.y));
  else p[0] |= lex_largest(a.y) ? M_LARGEST : M_SMALLEST;
}

// ---- G2
__global__ void __launch_bounds__(128) k_g2_decode(const uint8_t* __restrict__ in, uint32_t n, int raw, int subgroup,
                                                    G2Affine* __restrict__ out, uint8_t* __restrict__ ok_out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int rec = raw ? 128 : 64;
  const uint8_t* p = in + (size_t)i * rec;
  const uint8_t flag = p[0] & M_MASK;
  G2Affine r = G2Affine::inf();
  bool ok = false;
  if (raw) {
    if (flag == M_UNCOMPRESSED) {
      Fp x1 = load_be(p, false), x0 = load_be(p + 32, false), y1 = load_be(p + 64, false), y0 = load_be(p + 96, false);
      if (x0.is_zero() && x1.is_zero() && y0.is_zero() && y1.is_zero()) {
        ok = true;                 // uncompressed infinity: flag 0b00 + zeros
      } else if (reduced(x0) && reduced(x1) && reduced(y0) && reduced(y1)) {
        r = {{Fp::to_mont(x0), Fp::to_mont(x1)}, {Fp::to_mont(y0), Fp::to_mont(y1)}};
        ok = g2_on_curve(r);
      }
    }
  } else if (flag == M_INFINITY) {
    ok = all_zero(p, rec, true);
  } else if (flag != M_UNCOMPRESSED) {
    Fp x1 = load_be(p, true), x0 = load_be(p + 32, false);
    if (reduced(x0) && reduced(x1)) {
      Fp2 x = {Fp::to_mont(x0), Fp::to_mont(x1)};
      Fp2 y = fp2_sqrt(Fp2::add(Fp2::mul(Fp2::sqr(x), x), twist_b()), &ok);
      if (ok) {
        if (lex_largest(y) != (flag == M_LARGEST)) y = Fp2::neg(y);
        r = {x, y};
      }
    }
  }
  if (ok && subgroup && !r.is_inf()) ok = g2_in_subgroup(r);
  out[i] = ok ? r : G2Affine::inf();
  ok_out[i] = ok ? 1 : 0;
}

__global__ void __launch_bounds__(128) k_g2_encode(const G2Affine* __restrict__ pts, uint32_t n, int raw,
                                                    uint8_t* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int rec = raw ? 128 : 64;
  uint8_t* p = out + (size_t)i * rec;
  const G2Affine a = pts[i];
  if (a.is_inf()) {
    for (int k = 0; k < rec; k++) p[k] = 0;
    if (!raw) p[0] = M_INFINITY;
    return;
  }
  store_be(p, Fp::from_mont(a.x.c1));
  store_be(p + 32, Fp::from_mont(a.x.c0));
  if (raw) {
    store_be(p + 64, Fp::from_mont(a.y.c1));
    store_be(p + 96, Fp::from_mont(a.y.c0));
  } else {
    p[0] |= lex_largest(a.y) ? M_LARGEST : M_SMALLEST;
  }
}

// host drivers: staging through io_a (bytes) / io_b (points) / io_c (flags)
static int decode(b200g16_ctx* ctx, int group, const uint8_t* in, size_t n, int raw, int subgroup, uint64_t* out_points,
                  uint8_t* ok_out) {
  if (!ctx || (n && (!in || !out_points || !ok_out))) return fail(B200G16_ERR_ARG, "decode: null");
  if (n >= (1ull << 32)) return fail(B200G16_ERR_ARG, "decode: n too large");
  if (n == 0) return 0;
  std::lock_guard<std::mutex> lock(ctx->mu);
  B200_CUDA(cudaSetDevice(ctx->device));
  const size_t rec = (group == 1 ? 32 : 64) * (raw ? 2 : 1), psz = group == 1 ? sizeof(G1Affine) : sizeof(G2Affine);
  B200_TRY(ctx->io_a.ensure(n * rec));
  B200_TRY(ctx->io_b.ensure(n * psz));
  B200_TRY(ctx->io_c.ensure(n));
  cudaStream_t st = ctx->stream;
  B200_CUDA(cudaMemcpyAsync(ctx->io_a.p, in, n * rec, cudaMemcpyHostToDevice, st));
  const unsigned grid = (unsigned)((n + 127) / 128);
  if (group == 1)
    k_g1_decode<<<grid, 128, 0, st>>>(ctx->io_a.as<uint8_t>(), (uint32_t)n, raw, ctx->io_b.as<G1Affine>(), ctx->io_c.as<uint8_t>());
  else
    k_g2_decode<<<grid, 128, 0, st>>>(ctx->io_a.as<uint8_t>(), (uint32_t)n, raw, subgroup, ctx->io_b.as<G2Affine>(),
                                      ctx->io_c.as<uint8_t>());
  ctx->launches++;
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaMemcpyAsync(out_points, ctx->io_b.p, n * psz, cudaMemcpyDeviceToHost, st));
  B200_CUDA(cudaMemcpyAsync(ok_out, ctx->io_c.p, n, cudaMemcpyDeviceToHost, st));
  B200_CUDA(cudaStreamSynchronize(st));
  return 0;
}

static int encode(b200g16_ctx* ctx, int group, const uint64_t* points, size_t n, int raw, uint8_t* out) {
  if (!ctx || (n && (!points || !out))) return fail(B200G16_ERR_ARG, "encode: null");
  if (n >= (1ull << 32)) return fail(B200G16_ERR_ARG, "encode: n too large");
  if (n == 0) return 0;
  std::lock_guard<std::mutex> lock(ctx->mu);
  B200_CUDA(cudaSetDevice(ctx->device));
  const size_t rec = (group == 1 ? 32 : 64) * (raw ? 2 : 1), psz = group == 1 ? sizeof(G1Affine) : sizeof(G2Affine);
  B200_TRY(ctx->io_a.ensure(n * rec));
  B200_TRY(ctx->io_b.ensure(n * psz));
  cudaStream_t st = ctx->stream;
  B200_CUDA(cudaMemcpyAsync(ctx->io_b.p, points, n * psz, cudaMemcpyHostToDevice, st));
  const unsigned grid = (unsigned)((n + 127) / 128);
  if (group == 1) k_g1_encode<<<grid, 128, 0, st>>>(ctx->io_b.as<G1Affine>(), (uint32_t)n, raw, ctx->io_a.as<uint8_t>());
  else k_g2_encode<<<grid, 128, 0, st>>>(ctx->io_b.as<G2Affine>(), (uint32_t)n, raw, ctx->io_a.as<uint8_t>());
  ctx->launches++;
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaMemcpyAsync(out, ctx->io_a.p, n * rec, cudaMemcpyDeviceToHost, st));
  B200_CUDA(cudaStreamSynchronize(st));
  return 0;
}

}  // namespace b200

using namespace b200;

extern "C" {
int b200g16_g1_decode(b200g16_ctx* ctx, const uint8_t* in, size_t n, int raw, uint64_t* out_points, uint8_t* ok_out) {
  return decode(ctx, 1, in, n, raw, 0, out_points, ok_out);
}
int b200g16_g2_decode(b200g16_ctx* ctx, const uint8_t* in, size_t n, int raw, int subgroup_check, uint64_t* out_points,
                      uint8_t* ok_out) {
  return decode(ctx, 2, in, n, raw, subgroup_check, out_points, ok_out);
}
int b200g16_g1_encode(b200g16_ctx* ctx, const uint64_t* points, size_t n, int raw, uint8_t* out) {
  return encode(ctx, 1, points, n, raw, out);
}
int b200g16_g2_encode(b200g16_ctx* ctx, const uint64_t* points, size_t n, int raw, uint8_t* out) {
  return encode(ctx, 2, points, n, raw, out);
}
}
