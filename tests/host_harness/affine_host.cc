// Test harness (NOT part of libb200g16): compiles gnark_whir_b200/csrc/msm_affine.cuh for the host with g++ and
// runs the per-thread routine of k_accumulate_affine / k_aff_fixup for every "thread" in turn on synthetic sorted
// bucket lists, then compares every bucket's sum with a direct mixed-addition chain.  Pins the pair tree (piece
// descriptors, slot formula, odd leftovers, shares cut inside tasks, spills, empty threads, tangent / cancel /
// infinity cases) before it reaches a GPU.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "msm_affine.cuh"

using namespace b200;

namespace {
struct Lcg {
  uint64_t s;
  uint32_t next() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 33); }
  uint32_t below(uint32_t n) { return n ? next() % n : 0; }
};

// shape: 0 = every bucket 0..2*mean entries, 1 = one huge bucket + small ones, 2 = mostly empty buckets
template <class F>
int run_case(const Affine<F>* points, uint32_t npts, uint32_t nb, uint32_t mean, int shape, uint32_t seg, uint32_t T,
             int levels, uint32_t min_pairs, uint64_t seed) {
  Lcg rng{seed * 2 + 1};
  std::vector<uint32_t> counts(nb), offsets(nb), task_off(nb), entries, task_bucket;
  uint32_t E = 0, ntasks = 0;
  for (uint32_t b = 0; b < nb; b++) {
    uint32_t c = rng.below(2 * mean + 1);
    if (shape == 1) c = (b == nb / 3) ? mean * nb : rng.below(4);
    if (shape == 2) c = rng.below(8) == 0 ? rng.below(2 * mean + 1) : 0;
    counts[b] = c;
    offsets[b] = E;
    task_off[b] = ntasks;
    for (uint32_t k = 0; k < c; k++) entries.push_back((rng.below(npts) << 1) | (rng.next() & 1));
    for (uint32_t k = 0; k < (c + seg - 1) / seg; k++) task_bucket.push_back(b);
    E += c;
    ntasks += (c + seg - 1) / seg;
  }
  uint32_t totals[16] = {0};
  totals[0] = E; totals[1] = ntasks; totals[4] = seg;
  entries.push_back(0);
  task_bucket.push_back(0);
  std::vector<XYZZ<F>> partials(ntasks + 1), spill(T);
  std::vector<uint32_t> spill_task(T, 12345u);
  std::vector<AffDesc> desc(ntasks + T + 16);
  std::vector<std::vector<Affine<F>>> lvl(AFF_LEVELS_MAX + 1);
  AffArgs<F> A;
  memset(&A, 0, sizeof(A));
  A.bases = points; A.entries = entries.data(); A.task_bucket = task_bucket.data(); A.offsets = offsets.data();
  A.counts = counts.data(); A.task_off = task_off.data(); A.totals = totals; A.partials = partials.data();
  for (int l = 1; l <= AFF_LEVELS_MAX; l++) {
    lvl[l].assign((size_t)(E >> l) + ntasks + T + 16, Affine<F>{F::one(), F::one()});   // junk, never a valid result
    A.lvl[l] = lvl[l].data();
  }
  A.desc = desc.data(); A.spill = spill.data(); A.spill_task = spill_task.data();
  A.max_levels = levels; A.min_pairs = min_pairs;
  // The GPU runs the threads concurrently: no scratch slot (level buffers, descriptors) may be written by two of
  // them.  Sequential emulation would hide that, so diff the scratch after every thread and keep an owner map.
  std::vector<std::vector<int>> owner(AFF_LEVELS_MAX + 2);
  std::vector<std::vector<Affine<F>>> snap(lvl);
  std::vector<AffDesc> dsnap(desc);
  for (int l = 1; l <= AFF_LEVELS_MAX; l++) owner[l].assign(lvl[l].size(), -1);
  owner[0].assign(desc.size(), -1);
  int collisions = 0;
  for (uint32_t g = 0; g < T; g++) {
    aff_thread<F>(A, g, T);
    if (T > 4096) continue;   // (the diff is O(T * scratch))
    for (int l = 1; l <= AFF_LEVELS_MAX; l++)
      for (size_t k = 0; k < lvl[l].size(); k++)
        if (memcmp(&lvl[l][k], &snap[l][k], sizeof(Affine<F>)) != 0) {
          if (owner[l][k] >= 0 && owner[l][k] != (int)g) collisions++;
          owner[l][k] = (int)g;
          snap[l][k] = lvl[l][k];
        }
    for (size_t k = 0; k < desc.size(); k++)
      if (memcmp(&desc[k], &dsnap[k], sizeof(AffDesc)) != 0) {
        if (owner[0][k] >= 0 && owner[0][k] != (int)g) collisions++;
        owner[0][k] = (int)g;
        dsnap[k] = desc[k];
      }
  }
  if (collisions) return 1000000 + collisions;
  for (uint32_t g = 0; g < T; g++) aff_fixup_thread<F>(partials.data(), spill.data(), spill_task.data(), T, g);
  int bad = 0;
  for (uint32_t b = 0; b < nb; b++) {
    XYZZ<F> want = XYZZ<F>::inf(), got = XYZZ<F>::inf();
    for (uint32_t k = 0; k < counts[b]; k++) {
      const uint32_t e = entries[offsets[b] + k];
      Affine<F> p = points[e >> 1];
      if (e & 1) p.y = F::neg(p.y);
      want.madd(p);
    }
    for (uint32_t k = 0; k < (counts[b] + seg - 1) / seg; k++) got.add(partials[task_off[b] + k]);
    const Affine<F> w = want.to_affine(), h = got.to_affine();
    if (!(w.x == h.x) || !(w.y == h.y)) bad++;
  }
  return bad;
}
}  // namespace

extern "C" {
int affine_host_g1(const uint64_t* points, uint32_t npts, uint32_t nb, uint32_t mean, int shape, uint32_t seg, uint32_t T,
                   int levels, uint32_t min_pairs, uint64_t seed) {
  return run_case<Fp>(reinterpret_cast<const Affine<Fp>*>(points), npts, nb, mean, shape, seg, T, levels, min_pairs, seed);
}
int affine_host_g2(const uint64_t* points, uint32_t npts, uint32_t nb, uint32_t mean, int shape, uint32_t seg, uint32_t T,
                   int levels, uint32_t min_pairs, uint64_t seed) {
  return run_case<Fp2>(reinterpret_cast<const Affine<Fp2>*>(points), npts, nb, mean, shape, seg, T, levels, min_pairs, seed);
}
}
