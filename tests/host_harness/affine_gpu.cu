// Kernel-level harness (NOT part of libb200g16; built and run by tests/test_gpu_msm.py on the GPU box):
// the same synthetic bucket lists as affine_host.cc, through the real k_accumulate_affine / k_aff_fixup launches,
// compared bucket by bucket with a host-side mixed-addition chain.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 --expt-relaxed-constexpr -I gnark_whir_b200/csrc \
//        tests/host_harness/affine_gpu.cu -o affine_gpu && ./affine_gpu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "msm_affine.cuh"

using namespace b200;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e_)); exit(2); } } while (0)

struct Lcg {
  uint64_t s;
  uint32_t next() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 33); }
  uint32_t below(uint32_t n) { return n ? next() % n : 0; }
};

template <class T> T* to_dev(const std::vector<T>& v) {
  T* d;
  CK(cudaMalloc(&d, (v.size() + 1) * sizeof(T)));
  CK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

static int run_case(uint32_t npts, uint32_t nb, uint32_t mean, int shape, uint32_t seg, unsigned grid, int levels,
                    uint32_t min_pairs, uint64_t seed, bool few_points) {
  Lcg rng{seed * 2 + 1};
  // points: multiples of the generator (1, 2) in Montgomery form, by repeated mixed addition on the host
  std::vector<Affine<Fp>> points(npts);
  Affine<Fp> gen;
  gen.x = Fp::one();
  gen.y = Fp::add(Fp::one(), Fp::one());
  XYZZ<Fp> acc = XYZZ<Fp>::inf();
  for (uint32_t i = 0; i < npts; i++) {
    acc.madd(gen);
    points[i] = acc.to_affine();
  }
  if (few_points) {
    for (uint32_t i = 0; i < npts; i++) points[i] = points[i % 3];
    points[1] = Affine<Fp>::inf();
  }
  std::vector<uint32_t> counts(nb), offsets(nb), task_off(nb), entries, task_bucket;
  uint32_t E = 0, ntasks = 0;
  for (uint32_t b = 0; b < nb; b++) {
    uint32_t c = rng.below(2 * mean + 1);
    if (shape == 1) c = (b == nb / 3) ? mean * nb : rng.below(4);
    if (shape == 2) c = rng.below(8) == 0 ? rng.below(2 * mean + 1) : 0;
    counts[b] = c; offsets[b] = E; task_off[b] = ntasks;
    for (uint32_t k = 0; k < c; k++) entries.push_back((rng.below(npts) << 1) | (rng.next() & 1));
    for (uint32_t k = 0; k < (c + seg - 1) / seg; k++) task_bucket.push_back(b);
    E += c;
    ntasks += (c + seg - 1) / seg;
  }
  std::vector<uint32_t> totals(16, 0);
  totals[0] = E; totals[1] = ntasks; totals[4] = seg;
  const uint32_t T = grid * (uint32_t)AFF_THREADS;
  AffArgs<Fp> A;
  memset(&A, 0, sizeof(A));
  A.bases = to_dev(points); A.entries = to_dev(entries); A.task_bucket = to_dev(task_bucket); A.offsets = to_dev(offsets);
  A.counts = to_dev(counts); A.task_off = to_dev(task_off); A.totals = to_dev(totals);
  CK(cudaMalloc(&A.partials, (ntasks + 1) * sizeof(XYZZ<Fp>)));
  CK(cudaMemset(A.partials, 0x5a, (ntasks + 1) * sizeof(XYZZ<Fp>)));   // stale junk, as in the library's reused buffers
  for (int l = 1; l <= AFF_LEVELS_MAX; l++) {
    const size_t sz = ((size_t)(E >> l) + ntasks + T + 16) * sizeof(Affine<Fp>);
    CK(cudaMalloc(&A.lvl[l], sz));
    CK(cudaMemset(A.lvl[l], 0x5a, sz));
  }
  CK(cudaMalloc(&A.desc, (ntasks + T + 16) * sizeof(AffDesc)));
  CK(cudaMalloc(&A.spill, T * sizeof(XYZZ<Fp>)));
  CK(cudaMalloc(&A.spill_task, T * sizeof(uint32_t)));
  CK(cudaMemset(A.spill_task, 0x5a, T * sizeof(uint32_t)));
  A.max_levels = levels; A.min_pairs = min_pairs;
  k_accumulate_affine<Fp><<<grid, AFF_THREADS>>>(A);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  k_aff_fixup<Fp><<<(T + 127) / 128, 128>>>(A.partials, A.spill, A.spill_task, T);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<XYZZ<Fp>> partials(ntasks + 1);
  CK(cudaMemcpy(partials.data(), A.partials, ntasks * sizeof(XYZZ<Fp>), cudaMemcpyDeviceToHost));
  int bad = 0;
  for (uint32_t b = 0; b < nb; b++) {
    XYZZ<Fp> want = XYZZ<Fp>::inf(), got = XYZZ<Fp>::inf();
    for (uint32_t k = 0; k < counts[b]; k++) {
      const uint32_t e = entries[offsets[b] + k];
      Affine<Fp> p = points[e >> 1];
      if (e & 1) p.y = Fp::neg(p.y);
      want.madd(p);
    }
    for (uint32_t k = 0; k < (counts[b] + seg - 1) / seg; k++) got.add(partials[task_off[b] + k]);
    const Affine<Fp> w = want.to_affine(), h = got.to_affine();
    if (!(w.x == h.x) || !(w.y == h.y)) {
      if (bad < 4) printf("    bucket %u (count %u, first task %u, entries from %u) differs\n", b, counts[b], task_off[b], offsets[b]);
      bad++;
    }
  }
  printf("case nb=%u mean=%u shape=%d seg=%u T=%u levels=%d min_pairs=%u few=%d: E=%u tasks=%u  bad buckets=%d\n", nb, mean,
         shape, seg, T, levels, min_pairs, (int)few_points, E, ntasks, bad);
  cudaFree((void*)A.bases); cudaFree((void*)A.entries); cudaFree((void*)A.task_bucket); cudaFree((void*)A.offsets);
  cudaFree((void*)A.counts); cudaFree((void*)A.task_off); cudaFree((void*)A.totals); cudaFree(A.partials);
  for (int l = 1; l <= AFF_LEVELS_MAX; l++) cudaFree(A.lvl[l]);
  cudaFree(A.desc); cudaFree(A.spill); cudaFree(A.spill_task);
  return bad;
}

int main() {
  int bad = 0;
  // fewer entries than threads (one-entry shares, empty threads between the spills of a task): no levels run
  bad += run_case(300, 1376, 31, 0, 32, 444, 4, 1, 1, false);
  bad += run_case(300, 40, 20, 0, 32, 444, 4, 1, 2, false);
  bad += run_case(300, 40, 20, 0, 32, 1, 4, 1, 3, false);
  // levels 1..4 on small shares
  for (int levels = 1; levels <= 4; levels++) {
    bad += run_case(300, 2000, 400, 0, 500, 444, levels, 1, 10 + levels, false);
    bad += run_case(300, 3000, 300, 1, 256, 444, levels, 1, 20 + levels, false);
    bad += run_case(300, 4000, 800, 2, 1000, 444, levels, 1, 30 + levels, false);
    bad += run_case(300, 2000, 400, 0, 500, 444, levels, 1, 40 + levels, true);
  }
  printf(bad ? "FAILED\n" : "ALL OK\n");
  return bad ? 1 : 0;
}
