// Instruction-level integer-pipe probes: the measured roofline denominators for the MSM / NTT
// kernels ("Integer-pipe (IMAD / IMAD.WIDE) peak: to be measured").
//
// Each mode runs ACC independent dependency chains per thread, 128-thread CTAs, `bps` CTAs/SM:
//   0  mad.lo.u32                       (IMAD,          32x32+32 -> 32)
//   1  mad.wide.u32                     (IMAD.WIDE.U32, 32x32+64 -> 64, no carry in/out)
//   2  mad.lo.cc / madc.hi.cc chains    (IMAD.WIDE.U32.X: what the Montgomery multiplier issues)
//   3  mad.hi.u32                       (IMAD.HI.U32)
//   4  add.cc / addc chains             (IADD3.X on the ALU pipe)
//   5  mode 2 and mode 4 interleaved    (can the two pipes overlap?)
//   6  fma.rz.f64 chains                (DFMA: the FP64 pipe, csrc/fp52.cuh's multiplier)
//   7  mode 6 and three-input 64-bit integer adds interleaved (DFMA + IADD3 / IADD3.X: fp52's instruction mix)
//   8  mode 6 and mode 2 interleaved    (FP64 pipe + the integer multiplier's pipe: do they co-issue?)
//   9  add.rn.f64 chains                (DADD)
// b200g16_fp52_probe times whole Montgomery products: fp52.cuh's DFMA product alone and next to field.cuh's
// integer product in the same thread.
#include "common.cuh"
#include "fp52.cuh"

namespace b200 {

constexpr int PP_ACC = 8;
constexpr int PP_UNROLL = 16;

template <int MODE>
__global__ void __launch_bounds__(128) k_pipe_probe(uint32_t* __restrict__ out, int iters, uint32_t seed) {
  uint32_t a[PP_ACC], b[PP_ACC];
  uint64_t w[PP_ACC];
  double f[PP_ACC];
  uint32_t x = seed ^ (blockIdx.x * 128 + threadIdx.x) * 2654435761u, y = x * 40503u + 12345u;
#pragma unroll
  for (int k = 0; k < PP_ACC; k++) { a[k] = x + k; b[k] = y ^ k; w[k] = ((uint64_t)x << 32) | (y + k); f[k] = 1.0 + (double)((x + k) & 1023u) * 0x1p-20; }
  const double fm = 1.0 - (double)(y & 255u) * 0x1p-40, fa = (double)(x & 255u) * 0x1p-30;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < PP_UNROLL; u++) {
      if (MODE == 0) {
#pragma unroll
        for (int k = 0; k < PP_ACC; k++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(y), "r"(x));
      } else if (MODE == 1) {
#pragma unroll
        for (int k = 0; k < PP_ACC; k++)
          asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0; }" : "+l"(w[k]) : "r"(y));
      } else if (MODE == 2) {
        // one carry chain across the 8 (lo,hi) pairs, like cmad_n in field.cuh
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(a[0]), "+r"(b[0]) : "r"(x), "r"(y));
#pragma unroll
        for (int k = 1; k < PP_ACC; k++)
          asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(a[k]), "+r"(b[k]) : "r"(x), "r"(y));
      } else if (MODE == 3) {
#pragma unroll
        for (int k = 0; k < PP_ACC; k++) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(y), "r"(x));
      } else if (MODE == 4) {
        asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(a[0]) : "r"(x));
#pragma unroll
        for (int k = 1; k < PP_ACC; k++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(a[k]) : "r"(y));
      } else if (MODE == 6) {
#pragma unroll
        for (int k = 0; k < PP_ACC; k++) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(f[k]) : "d"(fm), "d"(fa));
      } else if (MODE == 7) {
#pragma unroll
        for (int k = 0; k < PP_ACC; k++) {
          asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(f[k]) : "d"(fm), "d"(fa));
          // w[k] += x:y + y:x as one three-input 64-bit addition (IADD3 with two carry-outs + IADD3.X)
          asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo, hi}, %0; add.cc.u32 lo, lo, %1; addc.u32 hi, hi, %2; "
                       "add.cc.u32 lo, lo, %2; addc.u32 hi, hi, %1; mov.b64 %0, {lo, hi}; }" : "+l"(w[k]) : "r"(x), "r"(y));
        }
      } else if (MODE == 8) {
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(a[0]), "+r"(b[0]) : "r"(x), "r"(y));
#pragma unroll
        for (int k = 1; k < PP_ACC; k++)
          asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(a[k]), "+r"(b[k]) : "r"(x), "r"(y));
#pragma unroll
        for (int k = 0; k < PP_ACC; k++) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(f[k]) : "d"(fm), "d"(fa));
      } else if (MODE == 9) {
#pragma unroll
        for (int k = 0; k < PP_ACC; k++) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(f[k]) : "d"(fa));
      } else {
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(a[0]), "+r"(b[0]) : "r"(x), "r"(y));
#pragma unroll
        for (int k = 1; k < PP_ACC; k++)
          asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(a[k]), "+r"(b[k]) : "r"(x), "r"(y));
        uint32_t lo = (uint32_t)w[0], hi = (uint32_t)(w[0] >> 32);
        asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(lo), "+r"(hi) : "r"(x), "r"(y));
        w[0] = ((uint64_t)hi << 32) | lo;
#pragma unroll
        for (int k = 1; k < 4; k++) {
          lo = (uint32_t)w[k]; hi = (uint32_t)(w[k] >> 32);
          asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(lo), "+r"(hi) : "r"(y), "r"(x));
          w[k] = ((uint64_t)hi << 32) | lo;
        }
      }
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int k = 0; k < PP_ACC; k++)
    r ^= a[k] ^ b[k] ^ (uint32_t)w[k] ^ (uint32_t)(w[k] >> 32) ^ (uint32_t)__double2loint(f[k]) ^ (uint32_t)__double2hiint(f[k]);
  out[blockIdx.x * 128 + threadIdx.x] = r;
}

// ops counted: mode 0,1,3: 1 per asm; mode 2: 1 per (lo,hi) pair (= one IMAD.WIDE.X); mode 4: 1 per add;
// mode 5: 8 wide MADs + 8 adds per unroll step (reports the wide-MAD rate); 6, 9: 1 per instruction;
// mode 7: 8 DFMA + 8 three-input 64-bit adds per step (reports the DFMA rate); mode 8: 8 wide MADs + 8 DFMA per
// step (reports the rate of EACH: the two pipes' rates are equal by construction)
int pipe_probe(b200g16_ctx* ctx, int mode, int blocks_per_sm, int iters, double* ops_per_s, float* ms_out) {
  size_t threads = (size_t)ctx->sm_count * blocks_per_sm * 128;
  B200_TRY(ctx->io_a.ensure(threads * sizeof(uint32_t)));
  uint32_t* d = ctx->io_a.as<uint32_t>();
  unsigned grid = (unsigned)(threads / 128);
  auto launch = [&]() {
    switch (mode) {
      case 0: k_pipe_probe<0><<<grid, 128, 0, ctx->stream>>>(d, iters, 1u); break;
      case 1: k_pipe_probe<1><<<grid, 128, 0, ctx->stream>>>(d, iters, 1u); break;
      case 2: k_pipe_probe<2><<<grid, 128, 0, ctx->stream>>>(d, iters, 1u); break;
      case 3: k_pipe_probe<3><<<grid, 128, 0, ctx->stream>>>(d, iters, 1u); break;
      case 4: k_pipe_probe<4><<<grid, 128, 0, ctx->stream>>>(d, iters, 1u); break;
      case 5: k_pipe_probe<5><<<grid, 128, 0, ctx->stream>>>(d, iters, 1u); break;
      case 6: k_pipe_probe<6><<<grid, 128, 0, ctx->stream>>>(d, iters, 1u); break;
      case 7: k_pipe_probe<7><<<grid, 128, 0, ctx->stream>>>(d, iters, 1u); break;
      case 8: k_pipe_probe<8><<<grid, 128, 0, ctx->stream>>>(d, iters, 1u); break;
      default: k_pipe_probe<9><<<grid, 128, 0, ctx->stream>>>(d, iters, 1u); break;
    }
  };
  launch();
  cudaEventRecord(ctx->ev[0], ctx->stream);
  launch();
  cudaEventRecord(ctx->ev[1], ctx->stream);
  ctx->launches += 2;
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
  *ms_out = ms;
  *ops_per_s = (double)threads * PP_ACC * PP_UNROLL * iters / (ms * 1e-3);
  return 0;
}

// ---- whole Montgomery products ---------------------------------------------------------------------
// NI chains of field.cuh's integer product and NF chains of fp52.cuh's DFMA product per thread.
template <int NI, int NF>
__global__ void __launch_bounds__(128) k_fp52_probe(uint32_t* __restrict__ data, int iters) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t w[8];
#pragma unroll
  for (int k = 0; k < 8; k++) w[k] = data[(size_t)i * 8 + k] & (k == 7 ? 0x0fffffffu : 0xffffffffu);
  Fp xi[NI > 0 ? NI : 1], yi;
  Fp52 xf[NF > 0 ? NF : 1], yf = Fp52::from_words(w);
#pragma unroll
  for (int k = 0; k < 8; k++) yi.l[k] = w[k];
#pragma unroll
  for (int c = 0; c < NI; c++) { xi[c] = yi; xi[c].l[0] ^= (uint32_t)c; }
#pragma unroll
  for (int c = 0; c < NF; c++) { w[0] ^= (uint32_t)(c + 1); xf[c] = Fp52::from_words(w); }
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < NI; c++) xi[c] = Fp::mul(xi[c], yi);
#pragma unroll
    for (int c = 0; c < NF; c++) xf[c] = Fp52::mul(xf[c], yf);
  }
  uint32_t o[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t[8];
#pragma unroll
  for (int c = 0; c < NI; c++)
#pragma unroll
    for (int k = 0; k < 8; k++) o[k] ^= xi[c].l[k];
#pragma unroll
  for (int c = 0; c < NF; c++) {
    xf[c].to_words(t);
#pragma unroll
    for (int k = 0; k < 8; k++) o[k] ^= t[k];
  }
#pragma unroll
  for (int k = 0; k < 8; k++) data[(size_t)i * 8 + k] = o[k];
}

// variant: 0 = 1 DFMA chain, 1 = 2 DFMA chains, 2 = 4 DFMA chains, 3 = 1 integer + 1 DFMA chain,
// 4 = 1 integer + 2 DFMA chains, 5 = 2 integer chains (reference point in the same harness)
// The device's DFMA product against the host's exact emulation of the same source (fp52.cuh) and against the
// integer multiplier's value of the same chain: thread 0 of a one-CTA launch, 64 dependent products.
static int fp52_selfcheck(b200g16_ctx* ctx) {
  const int iters = 64;
  B200_TRY(ctx->io_a.ensure(128 * 32));
  B200_CUDA(cudaMemsetAsync(ctx->io_a.p, 0x1a, 128 * 32, ctx->stream));
  k_fp52_probe<0, 1><<<1, 128, 0, ctx->stream>>>(ctx->io_a.as<uint32_t>(), iters);
  uint32_t got[8], w[8], want[8];
  B200_CUDA(cudaMemcpyAsync(got, ctx->io_a.p, 32, cudaMemcpyDeviceToHost, ctx->stream));
  B200_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < 8; k++) w[k] = 0x1a1a1a1au & (k == 7 ? 0x0fffffffu : 0xffffffffu);
  Fp52 y = Fp52::from_words(w);
  w[0] ^= 1u;
  Fp52 x = Fp52::from_words(w);
  for (int it = 0; it < iters; it++) x = Fp52::mul(x, y);
  x.to_words(want);
  if (memcmp(got, want, 32) != 0)
    return fail(B200G16_ERR_CUDA, "fp52_probe: the device's DFMA Montgomery product differs from the host emulation");
  return 0;
}

int fp52_probe(b200g16_ctx* ctx, int variant, int blocks_per_sm, int iters, double* modmul_per_s, float* ms_out) {
  B200_TRY(fp52_selfcheck(ctx));
  size_t threads = (size_t)ctx->sm_count * blocks_per_sm * 128;
  B200_TRY(ctx->io_a.ensure(threads * 32));
  B200_CUDA(cudaMemsetAsync(ctx->io_a.p, 0x1a, threads * 32, ctx->stream));
  uint32_t* d = ctx->io_a.as<uint32_t>();
  unsigned grid = (unsigned)(threads / 128);
  int per_iter = 1;
  auto launch = [&]() {
    switch (variant) {
      case 0: k_fp52_probe<0, 1><<<grid, 128, 0, ctx->stream>>>(d, iters); per_iter = 1; break;
      case 1: k_fp52_probe<0, 2><<<grid, 128, 0, ctx->stream>>>(d, iters); per_iter = 2; break;
      case 2: k_fp52_probe<0, 4><<<grid, 128, 0, ctx->stream>>>(d, iters); per_iter = 4; break;
      case 3: k_fp52_probe<1, 1><<<grid, 128, 0, ctx->stream>>>(d, iters); per_iter = 2; break;
      case 4: k_fp52_probe<1, 2><<<grid, 128, 0, ctx->stream>>>(d, iters); per_iter = 3; break;
      default: k_fp52_probe<2, 0><<<grid, 128, 0, ctx->stream>>>(d, iters); per_iter = 2; break;
    }
  };
  launch();
  cudaEventRecord(ctx->ev[0], ctx->stream);
  launch();
  cudaEventRecord(ctx->ev[1], ctx->stream);
  ctx->launches += 2;
  B200_CUDA(cudaGetLastError());
  B200_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
  *ms_out = ms;
  *modmul_per_s = (double)threads * per_iter * iters / (ms * 1e-3);
  return 0;
}

}  // namespace b200
