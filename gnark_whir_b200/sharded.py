"""Point-range sharding of one MSM across the GPUs of a box (one process per GPU).

BASELINE.json north_star: "MSMs are partitioned across the 8 GPUs of one 8xB200 box by point-range
shards plus a final 8-point reduction over NVLink/NCCL".  The reference itself is single-process
(gnark's MultiExp splits work across goroutines); the only data that has to cross
GPUs is one partial group element per rank:

    rank g:  partial_g = MSM(bases[lo_g:hi_g], scalars[lo_g:hi_g])      (local Pippenger, no traffic)
    all ranks: all_gather(partial_g)   -> G x 64 B (G1) or G x 128 B (G2) over NCCL / NVLink
    result = sum_g partial_g            (G-1 affine additions on the host, b200g16_g1_add/g2_add)

The exchange is latency-bound (512 B at G = 8), so it is a plain torch.distributed all_gather on
whatever backend the process group uses (nccl on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np

from . import lib


def shard_range(n, rank, world):
    """Contiguous, balanced [lo, hi) for `rank` of `world` (sizes differ by at most one)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_range_weighted(n, rank, weights):
    """Contiguous [lo, hi) for `rank` with shard sizes proportional to `weights` (one per rank)."""
    if not (0 <= rank < len(weights)):
        raise ValueError("rank out of range")
    tot = float(sum(weights))
    cuts = [0]
    acc = 0.0
    for w in weights:
        acc += w
        cuts.append(min(n, int(round(n * acc / tot))))
    cuts[-1] = n
    return cuts[rank], cuts[rank + 1]


def prove_weights(world, h_ranks_share=None):
    """Relative MSM shard sizes for the sharded prove.  With world >= 3 the first three ranks also run
    two of computeH's seven transforms each, so they get smaller point shards: share x < 1 such that
    (transforms + x * msm) on those ranks ~ msm on the others.  h_ranks_share: x, or None for the
    default tuned on the synthetic WHIR-shaped prove (bench.py --workload prove)."""
    if world < 3:
        return [1.0] * world
    x = h_ranks_share if h_ranks_share is not None else {4: 0.78, 8: 0.55}.get(world, max(0.3, 1.0 - 3.6 / world))
    return [x] * 3 + [1.0] * (world - 3)


def combine_partials(partials, group=1):
    """Sum of affine partial results (iterable of uint64[8] / uint64[16]) -> affine."""
    add = lib.g1_add if group == 1 else lib.g2_add
    acc = np.zeros(8 if group == 1 else 16, dtype=np.uint64)
    for p in partials:
        acc = add(acc, np.asarray(p, dtype=np.uint64))
    return acc


def exchange_and_combine(local_partial, group=1, device=None, pg=None):
    """all_gather the per-rank partial points and add them.  Works without an initialised process
    group (world of one).  `device`: torch device for the exchange buffer ('cuda:k' for nccl)."""
    import torch
    import torch.distributed as dist

    local = np.ascontiguousarray(local_partial, dtype=np.uint64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(pg) == 1:
        return combine_partials([local], group)
    t = torch.from_numpy(local.view(np.int64).copy())
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size(pg))]
    dist.all_gather(out, t, group=pg)
    parts = [o.cpu().numpy().view(np.uint64) for o in out]
    return combine_partials(parts, group)


class PointExchange:
    """all_gather of small per-rank host records (partial points) with every buffer allocated once: a pinned
    host staging pair and a device pair, so a step costs two small async copies, one NCCL all_gather and the
    host additions — no allocation, no pageable copy, no numpy <-> torch conversion per step."""

    def __init__(self, words, device=None, pg=None):
        import torch
        import torch.distributed as dist
        self.pg = pg
        self.world = dist.get_world_size(pg) if (dist.is_available() and dist.is_initialized()) else 1
        self.words, self.device = words, device
        self.cuda = device is not None and torch.device(device).type == "cuda"
        self.h_in = torch.empty(words, dtype=torch.int64)
        self.h_out = torch.empty(self.world * words, dtype=torch.int64)
        if self.cuda:
            self.h_in, self.h_out = self.h_in.pin_memory(), self.h_out.pin_memory()
            self.d_in = torch.empty(words, dtype=torch.int64, device=device)
            self.d_out = torch.empty(self.world * words, dtype=torch.int64, device=device)
        self.np_in = self.h_in.numpy().view(np.uint64)
        self.np_out = self.h_out.numpy().view(np.uint64)

    def gather(self, record):
        """record: uint64[words] -> (world, words) uint64 view (valid until the next call)"""
        import torch
        import torch.distributed as dist
        self.np_in[:] = np.asarray(record, dtype=np.uint64).reshape(self.words)
        if self.world == 1:
            self.np_out[:] = self.np_in
        elif self.cuda:
            self.d_in.copy_(self.h_in, non_blocking=True)
            dist.all_gather_into_tensor(self.d_out, self.d_in, group=self.pg)
            self.h_out.copy_(self.d_out, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        else:
            dist.all_gather_into_tensor(self.h_out, self.h_in, group=self.pg)
        return self.np_out.reshape(self.world, self.words)

    def combine(self, local_partial, group=1):
        """Sum over ranks of one partial point."""
        return combine_partials(list(self.gather(local_partial)), group)


PROOF_PARTS = (("msm_a", 8), ("msm_b1", 8), ("msm_k", 8), ("msm_z", 8), ("msm_b2", 16))


def upload_pk_shard(ctx, pk, rank, world):
    """Upload this rank's point-range shard of a groth16.ProvingKey (host arrays) -> pk handle.
    Every vector is cut with shard_range(len, rank, world); flags stay whole."""
    spans = [shard_range(len(v), rank, world) for v in (pk.G1_A, pk.G1_B, pk.G1_K, pk.G1_Z)]
    (a0, a1), (b0, b1), (k0, k1), (z0, z1) = spans
    return ctx.pk_upload(pk.log2_domain, len(pk.InfinityA), pk.G1_A[a0:a1], pk.G1_B[b0:b1], pk.G1_K[k0:k1],
                         pk.G1_Z[z0:z1], pk.G2_B[b0:b1], pk.G1_Alpha, pk.G1_Beta, pk.G1_Delta, pk.G2_Beta,
                         pk.G2_Delta, pk.InfinityA, pk.InfinityB, pk.k_skip, partial=True,
                         offsets=(a0, b0, k0, z0))


def pack_partials(proof_dict):
    return np.concatenate([np.asarray(proof_dict[k], dtype=np.uint64).reshape(w) for k, w in PROOF_PARTS])


def sum_partials(packed_list):
    """packed_list: per-rank 48-word vectors -> dict of the five complete MSM results."""
    out = {}
    o = 0
    for k, w in PROOF_PARTS:
        out[k] = combine_partials([p[o:o + w] for p in packed_list], 1 if w == 8 else 2)
        o += w
    return out


def prove_sharded(ctx, pk_shard, wires, a, b, c, r, s, device=None, pg=None):
    """One rank's part of a sharded Groth16 prove: every rank runs computeH (replicated) and the
    five MSMs on its shard, the 5 partial points (384 B) are all-gathered, and every rank
    finishes the proof on the host.  Returns the same dict as Context.prove."""
    import torch
    import torch.distributed as dist
    part, _ = ctx.prove(pk_shard, wires, a, b, c, r, s)
    packed = pack_partials(part)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(pg) > 1:
        t = torch.from_numpy(packed.view(np.int64).copy())
        if device is not None:
            t = t.to(device)
        outs = [torch.empty_like(t) for _ in range(dist.get_world_size(pg))]
        dist.all_gather(outs, t, group=pg)
        packed_list = [o.cpu().numpy().view(np.uint64) for o in outs]
    else:
        packed_list = [packed]
    sums = sum_partials(packed_list)
    return ctx.prove_finish(pk_shard, sums["msm_a"], sums["msm_b1"], sums["msm_k"], sums["msm_z"], sums["msm_b2"], r, s)


def h_vector_owner(v, world):
    """Rank that transforms vector v (0 = a, 1 = b, 2 = c) in compute_h_distributed."""
    return v % min(world, 3)


def compute_h_distributed(ctx, t_a, t_b, t_c, log2n, pg=None):
    """computeH spread over the ranks of a process group.  t_a, t_b, t_c: torch int64 tensors (N, 4) on
    this rank's GPU holding the same zero-padded a, b, c on every rank; h is left in t_a on every rank.

    gnark runs the FFTInverse + coset-FFT pairs of a, b and c in three goroutines; here up to three
    ranks take one vector each (2 of the 7 transforms), broadcast the coset evaluations over NCCL /
    NVLink (N x 32 B each), and every rank finishes with the pointwise step and the last transform.
    Critical path: 2 + 1 transforms + 3 broadcasts instead of 7 transforms.  The result is the same
    field elements in the same order as b200g16_compute_h_dev (bit-exact; tests/test_sharded_cpu.py
    checks the orchestration over gloo, tests/test_gpu_prove.py the arithmetic)."""
    import torch.distributed as dist

    world = dist.get_world_size(pg) if (dist.is_available() and dist.is_initialized()) else 1
    if world == 1:
        ctx.compute_h_dev(t_a.data_ptr(), t_b.data_ptr(), t_c.data_ptr(), log2n)
        return t_a
    rank = dist.get_rank(pg)
    vecs = (t_a, t_b, t_c)
    for v, t in enumerate(vecs):
        if h_vector_owner(v, world) == rank:
            ctx.ntt_dev(t.data_ptr(), log2n, inverse=True, decimation=lib.DIF)
            ctx.ntt_dev(t.data_ptr(), log2n, coset=True, decimation=lib.DIT)
    for v, t in enumerate(vecs):
        src = h_vector_owner(v, world)
        dist.broadcast(t, src=dist.get_global_rank(pg, src) if pg is not None else src, group=pg)
    ctx.h_pointwise_dev(t_a.data_ptr(), t_b.data_ptr(), t_c.data_ptr(), log2n)
    ctx.ntt_dev(t_a.data_ptr(), log2n, inverse=True, coset=True, decimation=lib.DIF)
    return t_a


class DistributedH:
    """computeH split over the 2 / 4 / 8 ranks of a process group (b200g16_dist_h_*): every rank keeps only its
    positions [rank M, (rank + 1) M) of a, b, c; the cross-GPU butterfly levels run over CUDA-IPC peer memory, the
    four phases are separated by barriers, and the rank's slice of h (= its Z shard) stays on its GPU.
    Bit-identical to b200g16_compute_h_dev (tests/test_gpu_group.py, and the gate of bench.py's sharded prove)."""

    def __init__(self, ctx, log2n, pg=None):
        import torch.distributed as dist
        self.ctx, self.L, self.pg = ctx, log2n, pg
        self.world, self.rank = dist.get_world_size(pg), dist.get_rank(pg)
        if self.world not in (2, 4, 8):
            raise ValueError("DistributedH needs 2, 4 or 8 ranks")
        self.M = (1 << log2n) // self.world
        mine = ctx.dist_h_init(log2n, self.world, self.rank).tobytes()
        outs = [None] * self.world
        dist.all_gather_object(outs, mine, group=pg)       # 192 bytes per rank; works on nccl and gloo groups alike
        ctx.dist_h_open(np.frombuffer(b"".join(outs), dtype=np.uint8))
        dist.barrier(group=pg)

    def slice_ptr(self, which):
        return self.ctx.dist_h_slice(which)

    def load(self, t_a, t_b, t_c):
        """t_*: torch int64 tensors (M, 4) on this GPU: this rank's zero-padded slices of a, b, c."""
        import torch
        for t in (t_a, t_b, t_c):
            assert t.is_cuda and t.is_contiguous() and t.numel() == 4 * self.M
        torch.cuda.current_stream().synchronize()          # the slices were produced on torch's stream
        self.ctx.dist_h_load(t_a.data_ptr(), t_b.data_ptr(), t_c.data_ptr())

    def run(self):
        """All four phases; returns the device pointer to pass as d_h (so that d_h + 32 off_z is this rank's slice)."""
        import torch.distributed as dist
        for phase in range(4):
            dist.barrier(group=self.pg)
            self.ctx.dist_h_phase(phase)
        return self.slice_ptr(0) - 32 * self.rank * self.M

    def close(self):
        import torch.distributed as dist
        dist.barrier(group=self.pg)
        self.ctx.dist_h_close()


def prove_distributed(ctx, pk_shard, t_wires, t_a, t_b, t_c, log2n, r, s, device=None, pg=None):
    """One rank's part of a sharded prove with computeH spread over the ranks and overlapped with the
    witness MSMs.  Tensors are torch int64 (n, 4) on this rank's GPU; a, b, c zero-padded, identical on all
    ranks, clobbered.  Ranks that own a vector (h_vector_owner) transform it first, so the broadcasts can
    start early, then enqueue their (smaller, see prove_weights) MSM shards; the other ranks enqueue their
    MSMs first and meet the broadcast while those run.  Every rank then finishes h and runs Z.
    Returns the finished proof dict (all ranks)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(pg) if (dist.is_available() and dist.is_initialized()) else 1
    if world == 1:
        return ctx.prove_dev(pk_shard, t_wires.data_ptr(), t_a.data_ptr(), t_b.data_ptr(), t_c.data_ptr(), r, s)
    rank = dist.get_rank(pg)
    vecs = (t_a, t_b, t_c)
    mine = [v for v in range(3) if h_vector_owner(v, world) == rank]
    if not mine:
        ctx.prove_begin_dev(pk_shard, t_wires.data_ptr())           # MSMs run while the owners transform
    for v in mine:
        ctx.ntt_dev(vecs[v].data_ptr(), log2n, inverse=True, decimation=lib.DIF)
        ctx.ntt_dev(vecs[v].data_ptr(), log2n, coset=True, decimation=lib.DIT)
    for v, t in enumerate(vecs):
        src = h_vector_owner(v, world)
        dist.broadcast(t, src=dist.get_global_rank(pg, src) if pg is not None else src, group=pg)
    if mine:
        ctx.prove_begin_dev(pk_shard, t_wires.data_ptr())
    if t_a.is_cuda:
        torch.cuda.current_stream().synchronize()                   # broadcasts landed before the library reads them
    ctx.h_pointwise_dev(t_a.data_ptr(), t_b.data_ptr(), t_c.data_ptr(), log2n)
    ctx.ntt_dev(t_a.data_ptr(), log2n, inverse=True, coset=True, decimation=lib.DIF)
    part = ctx.prove_end_dev(pk_shard, t_a.data_ptr(), r, s)
    packed = pack_partials(part)
    t = torch.from_numpy(packed.view(np.int64).copy())
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=pg)
    sums = sum_partials([o.cpu().numpy().view(np.uint64) for o in outs])
    return ctx.prove_finish(pk_shard, sums["msm_a"], sums["msm_b1"], sums["msm_k"], sums["msm_z"], sums["msm_b2"], r, s)


class ShardedBases:
    """This rank's slice of a point vector, resident on this rank's GPU."""

    def __init__(self, ctx, bases, lo, hi, n_total, group=1):
        self.ctx, self.bases, self.lo, self.hi, self.n_total, self.group = ctx, bases, lo, hi, n_total, group

    @classmethod
    def from_host(cls, ctx, points, rank, world, group=1):
        pts = np.asarray(points, dtype=np.uint64)
        n = pts.shape[0]
        lo, hi = shard_range(n, rank, world)
        up = ctx.upload_g1 if group == 1 else ctx.upload_g2
        return cls(ctx, up(pts[lo:hi]), lo, hi, n, group)

    def msm(self, scalars_full_or_slice, device=None, pg=None, sliced=False):
        """scalars: the full (n_total,4) vector (this rank uses rows lo..hi) or, with sliced=True,
        just this rank's rows.  Returns the full MSM result on every rank."""
        sc = np.asarray(scalars_full_or_slice, dtype=np.uint64).reshape(-1, 4)
        mine = sc if sliced else sc[self.lo:self.hi]
        partial = self.ctx.msm(self.bases, np.ascontiguousarray(mine))
        return exchange_and_combine(partial, self.group, device=device, pg=pg)

    def free(self):
        self.bases.free()
