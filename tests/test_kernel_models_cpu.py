"""Index algebra of three kernels, restated in Python and checked on the CPU (no GPU needed).

The CUDA code cannot run here; what CAN be pinned before it reaches a GPU is the arithmetic on lane numbers, array
slots and tree levels that those kernels are built on — the part where an off-by-one silently produces a wrong but
plausible result.  Each model follows its kernel line by line (same names) and is compared with an independent
definition:

  * k_merkle_paths_warp / keccak_lane_ctx (csrc/keccak.cu): one warp lane per Keccak state lane; theta from the two
    neighbouring COLUMNS, rho on the lane's own value, pi + chi fetched from the pre-pi lanes (s0, s1, s2)
        == oracle.keccak.keccak_f (itself pinned by hashlib, tests/test_oracle_cpu.py)
  * k_reduce_first / _level / _tail (csrc/msm_impl.cuh): in-place halving schedule -> G and U_m, S = G + sum 2^m U_m
        == sum_b (b + 1) B_b over the integers
  * k_tasks / k_merge_pass / k_merge_mid / k_merge_heavy (csrc/msm_common.cu, msm_impl.cuh): a bucket cut into nt tasks
    is merged by one fan-in-4 pass, then by one thread (5..16 tasks) or a halving tree over the first-level sums
        == the plain sum of the bucket's partials
The GPU suite checks the kernels themselves (tests/test_gpu_keccak.py, tests/test_gpu_msm.py)."""
import random

import pytest

from oracle import keccak as ok

MASK64 = (1 << 64) - 1
# FIPS-202 rho offsets by lane index x + 5y and the round constants, as in csrc/keccak.cu
K_ROT = [0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14]
K_RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B,
        0x0000000080000001, 0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088,
        0x0000000080008009, 0x000000008000000A, 0x000000008000808B, 0x800000000000008B, 0x8000000000008089,
        0x8000000000008003, 0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
        0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]


def _rol(v, n):
    n %= 64
    return ((v << n) | (v >> (64 - n))) & MASK64 if n else v


def _lane_ctx(lane):
    """keccak_lane_ctx: per-lane constants (lanes 25..31 mirror lane 0 and are never read)."""
    l = lane if lane < 25 else 0
    x, y = l % 5, l // 5
    cm0, cp0 = (x + 4) % 5, (x + 1) % 5

    def pi_src(X, Y):           # B[y', 2x' + 3y'] = A[x', y']: lane (X, Y) receives from y' = X, x' = 3 (Y - 3 y') mod 5
        yp = X
        xp = ((Y - 3 * yp) % 5 + 5) * 3 % 5
        return xp + 5 * yp
    return dict(cm0=cm0, cp0=cp0, rot=K_ROT[l], s0=pi_src(x, y), s1=pi_src((x + 1) % 5, y), s2=pi_src((x + 2) % 5, y))


def _keccak_f_warp_model(state):
    """keccak_f1600_warp<SMEM = true>: 32 lanes in lockstep, two shared arrays per round."""
    ctx = [_lane_ctx(lane) for lane in range(32)]
    a = list(state) + [0] * 7
    for rnd in range(24):
        sA = list(a)                                              # sA[lane] = a; __syncwarp
        r = [0] * 32
        for lane in range(32):
            c = ctx[lane]
            cm = cp = 0
            for k in range(5):
                cm ^= sA[c["cm0"] + 5 * k]
                cp ^= sA[c["cp0"] + 5 * k]
            r[lane] = _rol(a[lane] ^ cm ^ _rol(cp, 1), c["rot"])  # theta, then rho on the lane's own value
        sB = r                                                    # sB[lane] = r; __syncwarp
        for lane in range(32):
            c = ctx[lane]
            b, b1, b2 = sB[c["s0"]], sB[c["s1"]], sB[c["s2"]]
            a[lane] = (b ^ (~b1 & b2) ^ (K_RC[rnd] if lane == 0 else 0)) & MASK64
    return a[:25]


def test_keccak_warp_lane_layout_matches_the_permutation():
    rng = random.Random(1600)
    for _ in range(4):
        st = [rng.getrandbits(64) for _ in range(25)]
        assert _keccak_f_warp_model(st) == ok.keccak_f(list(st))
    assert _keccak_f_warp_model([0] * 25) == ok.keccak_f([0] * 25)
    # the three chi operands of a lane come from three different lanes, and every lane is somebody's B[x, y]
    s0 = sorted(_lane_ctx(lane)["s0"] for lane in range(25))
    assert s0 == list(range(25))


# ------------------------------------------------------------------------------------------- bucket-reduction tree
def _reduce_tree_model(buckets, tail_threads=8):
    """k_reduce_first + k_reduce_level (grid-wide levels) + k_reduce_tail (remaining levels, weights, final sum) on one
    window; group = the integers.  Returns sum_b (b + 1) * buckets[b] as the kernels compute it."""
    nbw = len(buckets)
    n = nbw.bit_length() - 1
    assert 1 << n == nbw and n >= 1
    A = [0] * nbw
    half = nbw >> 1
    for i in range(half):                                 # k_reduce_first (level 1): copy the upper half, fold onto the lower
        lo, hi = buckets[i], buckets[i + half]
        A[i + half] = hi
        A[i] = lo + hi
    l0 = 1                                                # msm_enqueue: first level the one-CTA tail can take
    while l0 <= n and l0 * (nbw >> l0) > tail_threads:
        l0 += 1
    l0 = max(l0, 2)

    def level(l):                                         # k_reduce_level / the loop of k_reduce_tail
        s = nbw >> l
        for g in range(l * s):
            j, i = g // s, g % s
            p = (nbw >> j if j else 0) + i
            A[p] += A[p + s]
    for l in range(2, min(l0, n + 1)):
        level(l)
    for l in range(l0 if l0 <= n else n + 1, n + 1):
        level(l)
    for t in range(1, n):                                 # weights: U_t sits at A[2^t]; position 1 = U_0 needs no doubling
        A[1 << t] <<= t

    def slot(idx):                                        # term idx of n + 1: G, U_0, U_1, ...
        return 0 if idx == 0 else 1 << (idx - 1)
    length = n + 1
    while length > 1:                                     # halving tree over the term list
        h = (length + 1) >> 1
        for t in range(length - h):
            A[slot(t)] += A[slot(t + h)]
        length = h
    return A[0]


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 11])
def test_bucket_reduction_tree_weights_every_bucket_by_its_index(n):
    rng = random.Random(n)
    buckets = [rng.randrange(1 << 40) for _ in range(1 << n)]
    want = sum((b + 1) * v for b, v in enumerate(buckets))
    for tail_threads in (1, 8, 512):                      # every split between grid-wide levels and the tail
        assert _reduce_tree_model(buckets, tail_threads) == want
    # a single occupied bucket picks out exactly its weight
    for b in (0, 1, (1 << n) - 1, (1 << n) // 2):
        one = [0] * (1 << n)
        one[b] = 1
        assert _reduce_tree_model(one) == b + 1


# ------------------------------------------------------------------------------------------- split-bucket merge
MERGE_FANIN = 4


def _merge_model(partials, nt):
    """One bucket cut into nt tasks (partials[0..nt)): k_merge_pass(stride 1), then the class k_tasks put it in."""
    p = list(partials)
    for j in range(0, nt, MERGE_FANIN):                   # k_merge_pass: local index j % 4 == 0 sums up to 4 partials
        acc = p[j]
        for q in range(1, MERGE_FANIN):
            if j + q >= nt:
                break
            acc += p[j + q]
        p[j] = acc
    if nt > 16:                                           # k_merge_heavy: halving tree over the first-level sums
        m = (nt + MERGE_FANIN - 1) // MERGE_FANIN
        while m > 1:
            half = (m + 1) >> 1
            for i in range(m - half):
                p[MERGE_FANIN * i] += p[MERGE_FANIN * (i + half)]
            m = half
    elif nt > 4:                                          # k_merge_mid: one thread adds the <= 4 first-level sums
        acc = p[0]
        for i in range(MERGE_FANIN, nt, MERGE_FANIN):
            acc += p[i]
        p[0] = acc
    return p[0]


def test_split_bucket_merge_classes_cover_every_task_count():
    rng = random.Random(44)
    for nt in list(range(1, 70)) + [255, 256, 257, 1000, 5001]:
        partials = [rng.randrange(1 << 50) for _ in range(nt)]
        assert _merge_model(partials, nt) == sum(partials), nt


def test_task_ordering_by_length_is_a_permutation():
    """k_task_hist / k_task_scan / k_task_order: per-CTA shared histogram, descending exclusive scan over the bins,
    rank inside the CTA + base of the CTA's range -> every task gets one slot, longer tasks first."""
    rng = random.Random(7)
    TASK_BINS, TASK_CTA = 4096, 1024
    lens = [rng.choice([32, 32, 32, 5, 17, 4095, 5000, 1]) for _ in range(5000)]
    bins = [min(x, TASK_BINS - 1) for x in lens]
    hist = [0] * TASK_BINS
    for b in bins:                                        # k_task_hist (CTA-local counts flushed to the global histogram)
        hist[b] += 1
    cursor, e = [0] * TASK_BINS, 0
    for b in range(TASK_BINS - 1, -1, -1):                # k_task_scan: exclusive scan in DESCENDING length order
        cursor[b] = e
        e += hist[b]
    order = [None] * len(lens)
    for cta in range(0, len(lens), TASK_CTA):             # k_task_order
        h = {}
        rank = []
        for t in range(cta, min(cta + TASK_CTA, len(lens))):
            rank.append(h.get(bins[t], 0))
            h[bins[t]] = rank[-1] + 1
        base = {}
        for b, cnt in h.items():
            base[b] = cursor[b]
            cursor[b] += cnt
        for k, t in enumerate(range(cta, min(cta + TASK_CTA, len(lens)))):
            order[base[bins[t]] + rank[k]] = t
    assert sorted(order) == list(range(len(lens)))
    by_len = [bins[t] for t in order]
    assert by_len == sorted(by_len, reverse=True)
