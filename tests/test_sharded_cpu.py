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


def test_weighted_shards_partition_and_follow_weights():
    for n in (0, 1, 7, 1000, (1 << 20) - 1):
        for weights in ([1.0], [1, 1], [0.55, 0.55, 0.55, 1, 1, 1, 1, 1], [0.78, 0.78, 0.78, 1.0]):
            cuts = [sharded.shard_range_weighted(n, r, weights) for r in range(len(weights))]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(len(cuts) - 1))
            assert all(lo <= hi for lo, hi in cuts)
    sizes = [hi - lo for lo, hi in (sharded.shard_range_weighted(1 << 20, r, [0.5, 0.5, 0.5, 1, 1, 1, 1, 1]) for r in range(8))]
    assert abs(sizes[0] / sizes[7] - 0.5) < 0.01
    assert sharded.prove_weights(2) == [1.0, 1.0] and len(sharded.prove_weights(8)) == 8
    assert sharded.prove_weights(8)[0] < 1.0 == sharded.prove_weights(8)[7]


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
    xch = sharded.PointExchange(8)                            # the preallocated exchange bench.py steps through
    same = all(np.array_equal(xch.combine(partial, 1), total) for _ in range(3))
    recs = xch.gather(np.full(8, rank + 1, dtype=np.uint64))
    same = same and recs.shape == (world, 8) and all((recs[r] == r + 1).all() for r in range(world))
    q.put((rank, bn.g1_from_array(total)[0] == exp and same))
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


class _OracleProveCtx(_OracleCtx):
    """Adds the two prove halves: the "MSMs" are faked as rank-dependent multiples of the generator, h's
    first element is folded into msm_z so a wrong or stale h shows up in the result."""

    def __init__(self, rank):
        super().__init__()
        self.rank, self.calls = rank, []

    def ntt_dev(self, *a, **k):
        self.calls.append("ntt")
        super().ntt_dev(*a, **k)

    def h_pointwise_dev(self, *a, **k):
        self.calls.append("pointwise")
        super().h_pointwise_dev(*a, **k)

    def prove_begin_dev(self, pk, d_wires):
        self.calls.append("begin")

    def prove_end_dev(self, pk, d_h, r, s):
        self.calls.append("end")
        h0 = bn.fr_from_mont_array(self._arr(d_h)[:1])[0]
        g1 = lambda k: bn.g1_to_array([bn.g1_mul(bn.G1_GEN, k)])[0]
        return {"msm_a": g1(self.rank + 1), "msm_b1": g1(2 * self.rank + 1), "msm_k": g1(5), "msm_z": g1(h0 % 1000 + 1),
                "msm_b2": bn.g2_to_array([bn.g2_mul(bn.G2_GEN, self.rank + 3)])[0]}

    def prove_finish(self, pk, a, b1, k, z, b2, r, s):
        return {"msm_a": a, "msm_b1": b1, "msm_k": k, "msm_z": z, "msm_b2": b2}


def _prove_worker(rank, world, port, logn, q):
    import random

    import torch
    import torch.distributed as dist
    from oracle import cport
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = random.Random(11)
    n = 1 << logn
    arrs = [bn.fr_to_mont_array([rng.randrange(R) for _ in range(n)]) for _ in range(3)]
    h = cport.compute_h(arrs[0], arrs[1], arrs[2], logn)
    h0 = bn.fr_from_mont_array(h[:1])[0]
    ts = [torch.from_numpy(x.view(np.int64).copy()) for x in arrs]
    wires = torch.zeros((4, 4), dtype=torch.int64)
    ctx = _OracleProveCtx(rank)
    ctx.reg(*ts)
    out = sharded.prove_distributed(ctx, None, wires, ts[0], ts[1], ts[2], logn, None, None)
    exp_a = bn.g1_to_array([bn.g1_mul(bn.G1_GEN, sum(r + 1 for r in range(world)))])[0]
    exp_z = bn.g1_to_array([bn.g1_mul(bn.G1_GEN, world * (h0 % 1000 + 1))])[0]
    owner = any(sharded.h_vector_owner(v, world) == rank for v in range(3))
    order_ok = (ctx.calls.index("begin") > ctx.calls.index("ntt")) if owner else ctx.calls[0] == "begin"
    order_ok = order_ok and ctx.calls[-1] == "end" and ctx.calls.index("pointwise") > ctx.calls.index("begin")
    q.put((rank, bool(np.array_equal(out["msm_a"], exp_a) and np.array_equal(out["msm_z"], exp_z) and order_ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_prove_distributed_orchestration_over_gloo():
    world = 4                                     # ranks 0-2 own a, b, c; rank 3 only runs MSMs
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_prove_worker, args=(r, world, port, 5, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == [(r, True) for r in range(world)]


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
