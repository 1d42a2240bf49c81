"""CPU test of the N>1 path: world_size-2 gloo processes exchange per-shard partial MSM results
and combine them with the library's host-side adder.  The partial MSMs themselves are computed
by the oracle here (no GPU in this container); on the GPU box the same exchange_and_combine runs
over NCCL with partials from b200g16_msm_g1 (see bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from gnark_whir_b200 import sharded
from oracle import bn254 as bn
from oracle.bn254 import R


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 1000, (1 << 24) - 1):
        for world in (1, 2, 3, 4, 8):
            spans = [sharded.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharded.shard_range(10, 2, 2)


def test_combine_partials_matches_oracle(rng):
    pts = [bn.g1_mul(bn.G1_GEN, rng.randrange(R)) for _ in range(5)] + [None]
    got = sharded.combine_partials(bn.g1_to_array(pts), 1)
    assert bn.g1_from_array(got)[0] == bn.g1_sum(pts)
    p2 = [bn.g2_mul(bn.G2_GEN, rng.randrange(R)) for _ in range(3)]
    assert bn.g2_from_array(sharded.combine_partials(bn.g2_to_array(p2), 2))[0] == bn.g2_sum(p2)


def _worker(rank, world, port, n, q):
    import random

    import torch.distributed as dist
    from oracle import cport
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = random.Random(99)                                   # same inputs on every rank
    k0, d = rng.randrange(R), rng.randrange(R)
    pts = cport.g1_progression(bn.fr_to_mont_array([k0]), bn.fr_to_mont_array([d]), n)
    ss = [rng.randrange(R) for _ in range(n)]
    sc = bn.fr_to_mont_array(ss)
    lo, hi = sharded.shard_range(n, rank, world)
    partial = cport.msm_g1(pts[lo:hi], sc[lo:hi], 1)          # stands in for the per-GPU MSM
    total = sharded.exchange_and_combine(partial, group=1)
    exp = bn.g1_mul(bn.G1_GEN, sum(s * (k0 + i * d) for i, s in enumerate(ss)) % R)
    q.put((rank, bn.g1_from_array(total)[0] == exp))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_exchange_and_combine():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 201, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


# ---- computeH spread over ranks (sharded.compute_h_distributed): orchestration over gloo, with the C oracle
# standing in for the per-GPU transforms
class _OracleCtx:
    """Duck-types the three lib.Context calls compute_h_distributed makes, on CPU tensors."""

    def __init__(self):
        self.t = {}

    def reg(self, *tensors):
        for t in tensors:
            self.t[t.data_ptr()] = t

    def _arr(self, ptr):
        return self.t[ptr].numpy().view(np.uint64)

    def ntt_dev(self, ptr, log2n, batch=1, inverse=False, coset=False, decimation=0):
        from oracle import cport
        a = self._arr(ptr)
        a[:] = cport.ntt(a, inverse=inverse, coset=coset, decimation=decimation)

    def h_pointwise_dev(self, pa, pb, pc, log2n):
        a, b, c = (bn.fr_from_mont_array(self._arr(p)) for p in (pa, pb, pc))
        den = pow(pow(5, 1 << log2n, R) - 1, -1, R)
        self._arr(pa)[:] = bn.fr_to_mont_array([(x * y - z) * den % R for x, y, z in zip(a, b, c)])


def _h_worker(rank, world, port, logn, q):
    import random

    import torch
    import torch.distributed as dist
    from oracle import cport
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = random.Random(7)                                    # same inputs on every rank
    n = 1 << logn
    a, b, c = ([rng.randrange(R) for _ in range(n)] for _ in range(3))
    arrs = [bn.fr_to_mont_array(v) for v in (a, b, c)]
    exp = cport.compute_h(arrs[0], arrs[1], arrs[2], logn)
    ts = [torch.from_numpy(x.view(np.int64).copy()) for x in arrs]
    ctx = _OracleCtx()
    ctx.reg(*ts)
    h = sharded.compute_h_distributed(ctx, ts[0], ts[1], ts[2], logn)
    q.put((rank, bool(np.array_equal(h.numpy().view(np.uint64), exp))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_compute_h_distributed_over_gloo(world):
    assert [sharded.h_vector_owner(v, 8) for v in range(3)] == [0, 1, 2]
    assert [sharded.h_vector_owner(v, 2) for v in range(3)] == [0, 1, 0]
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_h_worker, args=(r, world, port, 6, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == [(r, True) for r in range(world)]
