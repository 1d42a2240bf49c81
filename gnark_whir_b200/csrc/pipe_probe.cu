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
#include "common.cuh"

namespace b200 {

constexpr int PP_ACC = 8;
constexpr int PP_UNROLL = 16;

template <int MODE>
__global__ void __launch_bounds__(128) k_pipe_probe(uint32_t* __restrict__ out, int iters, uint32_t seed) {
  uint32_t a[PP_ACC], b[PP_ACC];
  uint64_t w[PP_ACC];
  uint32_t x = seed ^ (blockIdx.x * 128 + threadIdx.x) * 2654435761u, y = x * 40503u + 12345u;
#pragma unroll
  for (int k = 0; k < PP_ACC; k++) { a[k] = x + k; b[k] = y ^ k; w[k] = ((uint64_t)x << 32) | (y + k); }
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
  for (int k = 0; k < PP_ACC; k++) r ^= a[k] ^ b[k] ^ (uint32_t)w[k] ^ (uint32_t)(w[k] >> 32);
  out[blockIdx.x * 128 + threadIdx.x] = r;
}

// ops counted: mode 0,1,3: 1 per asm; mode 2: 1 per (lo,hi) pair (= one IMAD.WIDE.X); mode 4: 1 per add;
// mode 5: 8 wide MADs + 8 adds per unroll step (reports the wide-MAD rate)
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
      default: k_pipe_probe<5><<<grid, 128, 0, ctx->stream>>>(d, iters, 1u); break;
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

}  // namespace b200
