// Group-independent front half of the MSM: scalar recoding, histogram, scan, scatter.
// See msm_impl.cuh for the pipeline overview.
#include "msm_impl.cuh"

namespace b200 {


// r - 1 as 4 x u64 (for the top-window overflow check)
static const uint64_t kRm1[4] = {0x43e1f593f0000000ull, 0x2833e84879b97091ull, 0xb85045b68181585dull,
                                 0x30644e72e131a029ull};

static uint64_t shr256(const uint64_t v[4], int s) {  // low 64 bits of v >> s
  if (s >= 256) return 0;
  int q = s >> 6, r = s & 63;
  uint64_t lo = v[q] >> r;
  if (r && q + 1 < 4) lo |= v[q + 1] << (64 - r);
  return lo;
}

int msm_num_windows(int c) {
  int W = (254 + c - 1) / c;
  // the top window must absorb the carry of the signed recoding without overflowing
  uint64_t top = shr256(kRm1, (W - 1) * c) + 1;
  if (top >= (1ull << (c - 1))) W += 1;
  return W;
}

int msm_pick_window(size_t n) {
  // minimise W(c) * (n + k * 2^(c-1)): k models the serial running-sum cost of the bucket
  // reduction relative to one mixed add in a full-occupancy accumulate.
  int best = 4;
  double best_cost = 1e300;
  for (int c = 4; c <= 16; c++) {
    double cost = (double)msm_num_windows(c) * ((double)n + 6.0 * (double)(1u << (c - 1)));
    if (cost < best_cost) { best_cost = cost; best = c; }
  }
  return best;
}

// ------------------------------------------------------------------------------ kernels
__global__ void __launch_bounds__(256) k_digits(const Fr* __restrict__ scalars, uint32_t n, int c, int W,
                                                 uint32_t nbw, int32_t* __restrict__ digits,
                                                 uint32_t* __restrict__ counts) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr s = Fr::from_mont(scalars[i]);
  uint32_t limbs[9];
#pragma unroll
  for (int k = 0; k < 8; k++) limbs[k] = s.l[k];
  limbs[8] = 0;
  const uint32_t mask = (c == 32) ? 0xffffffffu : ((1u << c) - 1u);
  const uint32_t half = 1u << (c - 1);
  uint32_t carry = 0;
  for (int w = 0; w < W; w++) {
    int bit = w * c;
    int q = bit >> 5, r = bit & 31;
    uint32_t raw = 0;
    if (q < 8) {
      uint64_t two = (uint64_t)limbs[q] | ((uint64_t)limbs[q + 1] << 32);
      raw = (uint32_t)(two >> r) & mask;
    }
    raw += carry;
    int32_t d;
    if (raw >= half) { d = (int32_t)raw - (int32_t)(1u << c); carry = 1; }
    else { d = (int32_t)raw; carry = 0; }
    digits[(size_t)w * n + i] = d;
    if (d != 0) {
      uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
      atomicAdd(&counts[(uint32_t)w * nbw + mag - 1], 1u);
    }
  }
}

__global__ void __launch_bounds__(1024) k_scan(const uint32_t* __restrict__ counts, uint32_t nb, uint32_t seg,
                                                uint32_t* __restrict__ offsets, uint32_t* __restrict__ cursor,
                                                uint32_t* __restrict__ task_off, uint32_t* __restrict__ totals) {
  __shared__ uint32_t s1[1024], s2[1024];
  uint32_t t = threadIdx.x;
  uint32_t per = (nb + 1023) / 1024;
  uint32_t lo = t * per, hi = lo + per < nb ? lo + per : nb;
  if (lo > nb) lo = nb;
  uint32_t a = 0, b = 0;
  for (uint32_t i = lo; i < hi; i++) {
    uint32_t cnt = counts[i];
    a += cnt;
    b += (cnt + seg - 1) / seg;
  }
  s1[t] = a;
  s2[t] = b;
  __syncthreads();
  for (uint32_t off = 1; off < 1024; off <<= 1) {
    uint32_t x = 0, y = 0;
    if (t >= off) { x = s1[t - off]; y = s2[t - off]; }
    __syncthreads();
    s1[t] += x;
    s2[t] += y;
    __syncthreads();
  }
  uint32_t ea = s1[t] - a, eb = s2[t] - b;
  for (uint32_t i = lo; i < hi; i++) {
    uint32_t cnt = counts[i];
    offsets[i] = ea;
    cursor[i] = ea;
    task_off[i] = eb;
    ea += cnt;
    eb += (cnt + seg - 1) / seg;
  }
  if (t == 1023) {
    totals[0] = s1[1023];
    totals[1] = s2[1023];
    totals[2] = 0;  // heavy-bucket counter
  }
}

// grid: (ceil(n/256), W) — window-major so that one window's scatter targets (4n bytes) live in L2
__global__ void __launch_bounds__(256) k_scatter(const int32_t* __restrict__ digits, uint32_t n, uint32_t nbw,
                                                  uint32_t* __restrict__ cursor, uint32_t* __restrict__ entries) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t w = blockIdx.y;
  if (i >= n) return;
  int32_t d = digits[(size_t)w * n + i];
  if (d == 0) return;
  uint32_t neg = d < 0;
  uint32_t mag = neg ? (uint32_t)(-d) : (uint32_t)d;
  uint32_t pos = atomicAdd(&cursor[w * nbw + mag - 1], 1u);
  entries[pos] = (i << 1) | neg;
}

__global__ void __launch_bounds__(256) k_tasks(const uint32_t* __restrict__ counts, const uint32_t* __restrict__ task_off,
                                                uint32_t nb, uint32_t seg, uint32_t* __restrict__ task_bucket) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  uint32_t nt = (counts[b] + seg - 1) / seg;
  uint32_t o = task_off[b];
  for (uint32_t t = 0; t < nt; t++) task_bucket[o + t] = b;
}


int msm_sort_phase(b200g16_ctx* ctx, const MsmCfg& cfg, const Fr* d_scalars, uint32_t n, int32_t* digits,
                   uint32_t* counts, uint32_t* offsets, uint32_t* cursor, uint32_t* task_off, uint32_t* totals,
                   uint32_t* entries, uint32_t* task_bucket, int* ev) {
  cudaStream_t st = ctx->stream;
  auto mark = [&]() { if (ev && *ev < 18) cudaEventRecord(ctx->ev[(*ev)++], st); };
  mark();
  B200_CUDA(cudaMemsetAsync(counts, 0, (size_t)cfg.nb * sizeof(uint32_t), st));
  k_digits<<<cdiv(n, 256), 256, 0, st>>>(d_scalars, n, cfg.c, cfg.W, cfg.nbw, digits, counts);
  mark();
  k_scan<<<1, 1024, 0, st>>>(counts, cfg.nb, cfg.seg, offsets, cursor, task_off, totals);
  k_scatter<<<dim3(cdiv(n, 256), cfg.W), 256, 0, st>>>(digits, n, cfg.nbw, cursor, entries);
  k_tasks<<<cdiv(cfg.nb, 256), 256, 0, st>>>(counts, task_off, cfg.nb, cfg.seg, task_bucket);
  mark();
  ctx->launches += 4;
  B200_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b200
