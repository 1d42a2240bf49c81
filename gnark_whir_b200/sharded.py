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
