// Batched-affine bucket accumulation: the alternative to k_accumulate (msm_impl.cuh) for large inputs.
//
// Same job as gnark-crypto's batch-affine bucket path (ecc/bn254/multiexp_affine.go processChunkG1BatchAffine,
// reached through MultiExp from /root/reference/mt.go:496): an affine + affine addition costs
// lambda = (y2 - y1) / (x2 - x1), x3 = lambda^2 - x1 - x2, y3 = lambda (x1 - x3) - y1 — one product, one square,
// one product — once the inverse of x2 - x1 is known, and Montgomery's trick turns n inversions into ONE plus
// 3 (n - 1) products: ~5.8 products per addition against 9.06 for the mixed XYZZ addition.  The decomposition
// is B200-first, not gnark's (which queues additions to distinct buckets on one core):
//
//  * SIMT has no cheap way to share one inversion between lanes (32 lanes inverting cost what one does), so
//    every THREAD amortises its own inversion over a batch of >= ~200 independent additions.  A batch that size
//    does not fit on chip (64 B of operands + 32 B of prefix product per pending addition, 512 threads per SM),
//    so it streams through HBM: 180 GB and ~6.5 TB/s are what make this workable here.
//  * The independent additions come from a PAIR TREE over the sorted bucket lists.  A thread owns an equal
//    share [x0, x1) of the sorted entry array (cut anywhere, also in the middle of a task: perfect balance for
//    any scalar distribution) and reduces the pieces (task ∩ share) level by level: level l adds elements
//    2i and 2i + 1 of every piece (an odd last element is copied), so level l is one batch of ~share / 2^(l+1)
//    additions — forward pass: differences and running prefix products (parked in the output slot the sum will
//    take), ONE inversion, backward pass: the sums.  Element i of a piece that starts at entry s lives at
//    ceil(s / 2^l) + t + g + i of the level-l buffer (t = task, g = thread: strictly increasing along the
//    pieces, so slots never collide and nothing has to be compacted or counted).
//  * After `levels` halvings (chosen per launch so the deepest batch still pays for its inversion) the
//    remaining elements of a piece are summed by the usual mixed XYZZ chain -> partials[t].  A piece that does
//    not start its task goes to spill[g]; k_aff_fixup adds the spills to their task's partial.  Everything
//    after that (merge of split buckets, bucket reduction) is unchanged.
//  * Special cases cannot be ignored (real keys repeat bases; 0/1-heavy witnesses put the same point into a
//    bucket many times): equal x -> doubling (denominator 2y, numerator 3x^2) or P + (-P) = infinity;
//    infinity operands are copied.  They take a slow path; a zero never enters the running product.
//  * A thread's pieces are described once (first entry, length) in a private strip of a descriptor array; the
//    passes walk that strip with the next descriptor already in registers, so a piece boundary costs a dozen
//    ALU instructions — lanes of a warp cross their boundaries at different steps, and anything slower there
//    (the task tables are four dependent loads away) stalls the other 31 lanes every time.
//
// The per-thread routine is plain C++ over ec.cuh (no CUDA types), so tests/host_harness/affine_host.cc runs the
// same code on the CPU against a direct bucket sum.
#pragma once
#include "ec.cuh"

namespace b200 {

constexpr int AFF_LEVELS_MAX = 4;
constexpr uint32_t AFF_NONE = 0xffffffffu;   // spill_task: the thread's first piece starts its task (nothing to fix up)
constexpr uint32_t AFF_EMPTY = 0xfffffffeu;  // spill_task: the thread has no share at all
constexpr uint32_t AFF_STARTS = 0x80000000u; // descriptor flag: the piece begins its task

struct AffDesc {
  uint32_t s;  // first entry of the piece
  uint32_t m;  // length | AFF_STARTS
};

template <class F>
struct AffArgs {
  const Affine<F>* bases;
  const uint32_t* entries;
  const uint32_t* task_bucket;
  const uint32_t* offsets;
  const uint32_t* counts;
  const uint32_t* task_off;
  const uint32_t* totals;
  XYZZ<F>* partials;
  Affine<F>* lvl[AFF_LEVELS_MAX + 1];  // lvl[l], l = 1 .. max_levels: inputs of level l (= outputs of level l - 1)
  AffDesc* desc;                       // #tasks + #threads descriptors; thread g's strip starts at t0 + g
  XYZZ<F>* spill;                      // one per thread
  uint32_t* spill_task;                // task the spill belongs to, AFF_NONE / AFF_EMPTY if none
  int max_levels;
  uint32_t min_pairs;                  // a level is only run if the thread's share yields at least this many pairs
};

// NV 128-bit words of a base point: read-only path with a 64-byte L2 fetch granule on the device (a G1 point is one
// granule and its neighbours in the table are never wanted: the default promotion would double the DRAM traffic)
template <int NV>
B200_HD void aff_gather(void* dst, const void* src) {
#if defined(__CUDA_ARCH__)
  uint4* d = reinterpret_cast<uint4*>(dst);
  const uint4* s = reinterpret_cast<const uint4*>(src);
#pragma unroll
  for (int k = 0; k < NV; k++)
    asm volatile("ld.global.nc.L2::64B.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(d[k].x), "=r"(d[k].y), "=r"(d[k].z), "=r"(d[k].w) : "l"(s + k));
#else
  const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
  uint32_t* d = reinterpret_cast<uint32_t*>(dst);
  for (int k = 0; k < 4 * NV; k++) d[k] = s[k];
#endif
}

// first entry of task t
B200_HD uint32_t aff_task_start(uint32_t t, const uint32_t* task_bucket, const uint32_t* task_off, const uint32_t* offsets,
                                uint32_t seg) {
  const uint32_t b = task_bucket[t];
  return offsets[b] + (t - task_off[b]) * seg;
}
// smallest t in [0, ntasks] whose first entry is > x (tasks tile [0, E) in task order)
B200_HD uint32_t aff_upper_bound(uint32_t x, uint32_t ntasks, const uint32_t* task_bucket, const uint32_t* task_off,
                                 const uint32_t* offsets, uint32_t seg) {
  uint32_t lo = 0, hi = ntasks;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (aff_task_start(mid, task_bucket, task_off, offsets, seg) > x) hi = mid;
    else lo = mid + 1;
  }
  return lo;
}

B200_HD uint32_t ceil_shr(uint32_t x, int l) { return (x + ((1u << l) - 1u)) >> l; }

// the piece of task t inside the share [x0, x1)
template <class F>
B200_HD AffDesc aff_piece(const AffArgs<F>& A, uint32_t t, uint32_t x0, uint32_t x1, uint32_t seg) {
  const uint32_t b = A.task_bucket[t];
  const uint32_t ob = A.offsets[b];
  const uint32_t ts = ob + (t - A.task_off[b]) * seg;
  const uint32_t be = ob + A.counts[b];
  const uint32_t te = ts + seg < be ? ts + seg : be;
  AffDesc p;
  p.s = ts > x0 ? ts : x0;
  p.m = ((te < x1 ? te : x1) - p.s) | (ts >= x0 ? AFF_STARTS : 0u);
  return p;
}

// element idx of a piece at level l: level 0 = the signed base point behind sorted entry s + idx
template <class F>
B200_HD Affine<F> aff_load(const AffArgs<F>& A, int l, uint32_t s, uint32_t inb, uint32_t idx) {
  if (l == 0) {
    const uint32_t e = A.entries[s + idx];
    Affine<F> p;
    aff_gather<sizeof(Affine<F>) / 16>(&p, A.bases + (e >> 1));
    if (e & 1) p.y = F::neg(p.y);
    return p;
  }
  return A.lvl[l][inb + idx];
}
template <class F>
B200_HD F aff_load_x(const AffArgs<F>& A, int l, uint32_t s, uint32_t inb, uint32_t idx) {
  if (l == 0) {
    F r;
    aff_gather<sizeof(F) / 16>(&r, &A.bases[A.entries[s + idx] >> 1].x);
    return r;
  }
  return A.lvl[l][inb + idx].x;
}

// What a + b is.  0: chord (den = bx - ax, num = by - ay); 1: tangent (den = 2 ay, num = 3 ax^2);
// 2: a (b is infinity); 3: b (a is infinity); 4: infinity.
template <class F>
B200_HD_NOINLINE int aff_classify(const Affine<F>& a, const Affine<F>& b, F& den, F& num) {
  if (a.is_inf()) return 3;
  if (b.is_inf()) return 2;
  den = F::sub(b.x, a.x);
  if (!den.is_zero()) {
    num = F::sub(b.y, a.y);
    return 0;
  }
  if (!(a.y == b.y) || a.y.is_zero()) return 4;
  den = F::dbl(a.y);
  F x2 = F::sqr(a.x);
  num = F::add(F::dbl(x2), x2);
  return 1;
}

template <class F>
B200_HD Affine<F> aff_finish(const Affine<F>& a, const F& bx, const F& num, const F& inv_den) {
  const F lam = F::mul(num, inv_den);
  Affine<F> r;
  r.x = F::sub(F::sub(F::sqr(lam), a.x), bx);
  r.y = F::sub(F::mul(lam, F::sub(a.x, r.x)), a.y);
  return r;
}

// One level of the pair tree over the P pieces of thread g (descriptor strip D, first task t0) of share [x0, x1).
template <class F>
B200_HD void aff_level(const AffArgs<F>& A, int l, uint32_t g, uint32_t t0, uint32_t P) {
  const AffDesc* const D = A.desc + t0 + g;
  Affine<F>* const out = A.lvl[l + 1];
  // ---- forward: running product of the denominators; the value BEFORE pair q is parked in out[q].x
  F run = F::one();
  {
    uint32_t j = 0, i = 0, np = 0, c = 0, s = 0, inb = 0, outb = 0;
    bool open = false;
    AffDesc nd = D[0];
    for (;;) {
      // piece boundary as an INNER loop (it also skips pieces without a pair): the lanes of a warp reach their
      // boundaries at different steps, and a `continue` here lets the compiler keep the two paths apart for good —
      // the warp then runs the body below once per group of lanes (measured: 22.8 of 32 lanes active per instruction,
      // 30.7 with the inner loop)
      bool more = true;
      while (i >= np) {
        if (open && (c & 1)) out[outb + np] = aff_load(A, l, s, inb, c - 1);  // odd element: passes through
        open = false;
        if (j >= P) { more = false; break; }
        const AffDesc d = nd;
        const uint32_t tg = t0 + j + g;
        j++;
        if (j < P) nd = D[j];
        s = d.s;
        c = ceil_shr(d.m & ~AFF_STARTS, l);
        np = c >> 1;
        i = 0;
        inb = ceil_shr(s, l) + tg;
        outb = ceil_shr(s, l + 1) + tg;
        open = true;
      }
      if (!more) break;
      const F ax = aff_load_x(A, l, s, inb, 2 * i), bx = aff_load_x(A, l, s, inb, 2 * i + 1);
      F den = F::sub(bx, ax);
      bool live = true;
      if (ax.is_zero() || bx.is_zero() || den.is_zero()) {  // rare: infinity operand, doubling or cancellation
        F num;
        const int cls = aff_classify(aff_load(A, l, s, inb, 2 * i), aff_load(A, l, s, inb, 2 * i + 1), den, num);
        live = cls <= 1;
      }
      if (live) {
        out[outb + i].x = run;
        run = F::mul(run, den);
      }
      i++;
    }
  }
  F inv = F::inv(run);
  // ---- backward: inverse of each denominator from the running inverse and the parked prefix, then the sum
  {
    uint32_t j = P, np = 0, s = 0, inb = 0, outb = 0;
    int32_t i = -1;
    AffDesc nd = D[P - 1];
    for (;;) {
      bool more = true;
      while (i < 0) {
        if (j == 0) { more = false; break; }
        const AffDesc d = nd;
        j--;
        const uint32_t tg = t0 + j + g;
        if (j > 0) nd = D[j - 1];
        s = d.s;
        np = ceil_shr(d.m & ~AFF_STARTS, l) >> 1;
        i = (int32_t)np - 1;
        inb = ceil_shr(s, l) + tg;
        outb = ceil_shr(s, l + 1) + tg;
      }
      if (!more) break;
      const Affine<F> a = aff_load(A, l, s, inb, 2 * (uint32_t)i), b = aff_load(A, l, s, inb, 2 * (uint32_t)i + 1);
      F den = F::sub(b.x, a.x), num;
      int cls = 0;
      if (a.x.is_zero() || b.x.is_zero() || den.is_zero()) cls = aff_classify(a, b, den, num);
      else num = F::sub(b.y, a.y);
      Affine<F> r;
      if (cls <= 1) {
        const F inv_den = F::mul(inv, out[outb + i].x);
        inv = F::mul(inv, den);
        r = aff_finish(a, cls == 0 ? b.x : a.x, num, inv_den);
      } else {
        r = cls == 2 ? a : (cls == 3 ? b : Affine<F>::inf());
      }
      out[outb + i] = r;
      i--;
    }
  }
}

// Everything thread g of T does: its share of the E sorted entries -> partials / spill.
template <class F>
B200_HD void aff_thread(const AffArgs<F>& A, uint32_t g, uint32_t T) {
  const uint32_t E = A.totals[0], ntasks = A.totals[1], seg = A.totals[4];
  const uint32_t x0 = (uint32_t)(((uint64_t)g * E) / T), x1 = (uint32_t)(((uint64_t)(g + 1) * E) / T);
  A.spill_task[g] = x0 >= x1 ? AFF_EMPTY : AFF_NONE;
  if (x0 >= x1) return;
  const uint32_t t0 = aff_upper_bound(x0, ntasks, A.task_bucket, A.task_off, A.offsets, seg) - 1;
  const uint32_t P = aff_upper_bound(x1 - 1, ntasks, A.task_bucket, A.task_off, A.offsets, seg) - t0;
  AffDesc* const D = A.desc + t0 + g;
  for (uint32_t j = 0; j < P; j++) D[j] = aff_piece(A, t0 + j, x0, x1, seg);
  // levels: as long as the share still yields min_pairs additions per inversion (same count in every thread)
  const uint32_t M = E / T;
  int L = 0;
  while (L < A.max_levels && (M >> (L + 1)) >= A.min_pairs) L++;
  for (int l = 0; l < L; l++) aff_level<F>(A, l, g, t0, P);
  // ---- what is left of every piece: mixed XYZZ chain
  uint32_t j = 0, i = 0, c = 0, s = 0, inb = 0, cur = 0;
  bool open = false, starts = false;
  XYZZ<F> acc = XYZZ<F>::inf();
  AffDesc nd = D[0];
  for (;;) {
    bool more = true;
    while (i >= c) {
      if (open) {
        if (starts) A.partials[cur] = acc;
        else { A.spill[g] = acc; A.spill_task[g] = cur; }
      }
      open = false;
      if (j >= P) { more = false; break; }
      const AffDesc d = nd;
      cur = t0 + j;
      j++;
      if (j < P) nd = D[j];
      s = d.s;
      starts = (d.m & AFF_STARTS) != 0;
      c = ceil_shr(d.m & ~AFF_STARTS, L);
      i = 0;
      inb = ceil_shr(s, L) + cur + g;
      acc = XYZZ<F>::inf();
      open = true;
    }
    if (!more) break;
    acc.madd(aff_load(A, L, s, inb, i));
    i++;
  }
}

// spill[g] (a piece that continues a task begun by an earlier thread) -> partials[task].  The first spill of a
// task collects the following ones (a task longer than a share spans several threads).
template <class F>
B200_HD void aff_fixup_thread(XYZZ<F>* partials, const XYZZ<F>* spill, const uint32_t* spill_task, uint32_t T, uint32_t g) {
  const uint32_t t = spill_task[g];
  if (t == AFF_NONE || t == AFF_EMPTY) return;
  // threads without a share (fewer entries than threads) sit between the spills of a task: AFF_EMPTY
  for (uint32_t h = g; h > 0;) {
    const uint32_t u = spill_task[--h];
    if (u == t) return;          // an earlier thread leads this task's spills
    if (u != AFF_EMPTY) break;
  }
  XYZZ<F> acc = partials[t];
  for (uint32_t h = g; h < T; h++) {
    const uint32_t u = spill_task[h];
    if (u == AFF_EMPTY) continue;
    if (u != t) break;
    acc.add(spill[h]);
  }
  partials[t] = acc;
}

#if defined(__CUDACC__)
// G1: 128 registers -> 4 CTAs of 128 threads per SM (3 CTAs with the ~140 registers the compiler would like: 10.5 ms
// against 9.9 at 2^22); G2: 2.  Software prefetch of the next operands into L2 / L1 (prefetch.global) was measured
// too: slower at every distance (13.1 ms), the address needs the next sorted entry — one more dependent load.
constexpr int AFF_THREADS = 128;
template <class F>
__global__ void __launch_bounds__(AFF_THREADS, (sizeof(F) > 32) ? 2 : 4) k_accumulate_affine(const AffArgs<F> A) {
  aff_thread<F>(A, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

template <class F>
__global__ void __launch_bounds__(128) k_aff_fixup(XYZZ<F>* partials, const XYZZ<F>* spill, const uint32_t* spill_task,
                                                   uint32_t T) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < T) aff_fixup_thread<F>(partials, spill, spill_task, T, g);
}
#endif

}  // namespace b200
