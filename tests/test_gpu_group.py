"""GPU parity of the one-process multi-GPU C-ABI (b200g16_group_*): the single-process call site of the reference
(/root/reference/mt.go:496) spread over several devices.  A group may name a device more than once, so the
sharding, the peer copies of the coset evaluations / h slices and the host-side combination are exercised on a
single-GPU box too ([0, 0], [0, 0, 0], [0] * 5); with >= 2 GPUs the same tests also run on distinct devices."""
import random

import numpy as np
import pytest
import torch

from gnark_whir_b200 import groth16 as g16
from gnark_whir_b200 import lib
from oracle import bn254 as bn
from oracle import cport, synth
from oracle import groth16 as og
from oracle.bn254 import R

pytestmark = pytest.mark.gpu


def _device_lists():
    out = [[0], [0, 0], [0, 0, 0], [0] * 4, [0] * 5, [0] * 8]     # 2 / 4 / 8: computeH split over all devices
    n = torch.cuda.device_count()
    if n >= 2:
        out.append([0, 1])
    if n >= 4:
        out.append([0, 1, 2, 3])
    return out


@pytest.mark.parametrize("devices", _device_lists(), ids=lambda d: "dev" + "_".join(map(str, d)))
def test_group_msm_matches_closed_form(ctx, devices):
    rs = np.random.Generator(np.random.PCG64(100 + len(devices)))
    n = 50_000 + 7
    ks = synth.rand_fr(rs, n)
    p1 = ctx.fixed_base_mul(synth.G1, ks, group=1)
    p2 = ctx.fixed_base_mul(synth.G2, ks[:3001], group=2)
    sc = synth.whir_mix(rs, n)
    with lib.Group(devices) as g:
        assert len(g) == len(devices)
        for pre in (False, True):
            b1 = g.upload(p1, group=1, precompute=pre)
            assert np.array_equal(g.msm(b1, sc), cport.g1_gen_mul(cport.fr_dot(ks, sc)))
            g.bases_free(b1)
        b2 = g.upload(p2, group=2)
        assert np.array_equal(g.msm(b2, sc[:3001]), cport.g2_gen_mul(cport.fr_dot(ks[:3001], sc[:3001])))
        with pytest.raises(lib.B200Error):          # length must match the sharded vector
            g.msm(b2, sc[:3000])
        g.bases_free(b2)


@pytest.mark.parametrize("devices", _device_lists(), ids=lambda d: "dev" + "_".join(map(str, d)))
def test_group_prove_equals_single_gpu_prove_and_oracle(ctx, devices):
    """A real (tau-derived) key from the oracle's Setup with points at infinity in A / B, ragged shards."""
    rng = random.Random(500 + len(devices))
    r1cs, w = og.synthetic_r1cs(300, 4, rng)
    tw = og.ToxicWaste(*[rng.randrange(1, R) for _ in range(5)], sigma=rng.randrange(1, R))
    opk, ovk = og.setup(r1cs, tw)
    r, s = rng.randrange(R), rng.randrange(R)
    oproof, aux = og.prove(r1cs, opk, w, r, s)
    k_skip = np.ones(r1cs.nb_wires, np.uint8)
    k_skip[opk.k_wires] = 0
    a, b, c = og.solve_abc(r1cs, w)
    f = bn.fr_to_mont_array
    with lib.Group(devices) as g:
        for pre in (False, True):
            pk = g.pk_upload(opk.domain.logn, r1cs.nb_wires, bn.g1_to_array(opk.A), bn.g1_to_array(opk.B), bn.g1_to_array(opk.K),
                             bn.g1_to_array(opk.Z), bn.g2_to_array(opk.B2), bn.g1_to_array([opk.alpha1])[0],
                             bn.g1_to_array([opk.beta1])[0], bn.g1_to_array([opk.delta1])[0], bn.g2_to_array([opk.beta2])[0],
                             bn.g2_to_array([opk.delta2])[0], np.array(opk.infinity_a, np.uint8), np.array(opk.infinity_b, np.uint8),
                             k_skip, precompute=pre)
            for it in range(2):
                got, h = g.prove(pk, f(w), f(a), f(b), f(c), f([r])[0], f([s])[0], want_h=True, log2_domain=opk.domain.logn)
                assert np.array_equal(h, f(aux["h"]))
                assert np.array_equal(got["ar"], bn.g1_to_array([oproof.Ar])[0])
                assert np.array_equal(got["bs"], bn.g2_to_array([oproof.Bs])[0])
                assert np.array_equal(got["krs"], bn.g1_to_array([oproof.Krs])[0])
                assert np.array_equal(got["msm_z"], bn.g1_to_array([aux["krs2"]])[0])
                assert np.array_equal(got["bs1"], bn.g1_to_array([aux["bs1"]])[0])
            g.pk_free(pk)
        with pytest.raises(lib.B200Error):
            g.prove(None, f(w), f(a), f(b), f(c), f([r])[0], f([s])[0])


@pytest.mark.parametrize("n_dev", [3, 4, 8])
def test_group_prove_2p18_known_dlog_key(ctx, n_dev):
    """2^18 constraints, window tables, 3 shards (a, b, c on three devices, h on the root) and 4 / 8 shards (computeH
    split over all devices, cross-GPU levels over peer memory): every MSM output, h and the proof against the closed
    forms."""
    L, N = 18, 1 << 18
    rs = np.random.Generator(np.random.PCG64(1818))
    ks = {"a": synth.rand_fr(rs, N), "b": synth.rand_fr(rs, N), "k": synth.rand_fr(rs, N - 1), "z": synth.rand_fr(rs, N - 1),
          "b2": synth.rand_fr(rs, N)}
    small = synth.rand_fr(rs, 5)
    pts = {n: ctx.fixed_base_mul(synth.G1, ks[n], group=1) for n in ("a", "b", "k", "z")}
    pts["b2"] = ctx.fixed_base_mul(synth.G2, ks["b2"], group=2)
    g1s, g2s = ctx.fixed_base_mul(synth.G1, small[:3], group=1), ctx.fixed_base_mul(synth.G2, small[3:], group=2)
    k_skip = np.zeros(N, np.uint8)
    k_skip[0] = 1
    zeros = np.zeros(N, np.uint8)
    wires = synth.whir_mix(rs, N)
    a, b, c = synth.rand_fr(rs, N - 9), synth.rand_fr(rs, N - 9), synth.rand_fr(rs, N - 9)
    r, s = synth.rand_fr(rs, 1)[0], synth.rand_fr(rs, 1)[0]
    # closed forms (same arithmetic as oracle.synth.KnownDlogKey.expected, on host-side discrete logs)
    key = synth.KnownDlogKey.__new__(synth.KnownDlogKey)
    key.L, key.N, key.k, key.small = L, N, ks, small
    exp, h_exp = key.expected(wires, a, b, c, r, s)
    nd = torch.cuda.device_count()
    devices = list(range(n_dev)) if nd >= n_dev else [0] * n_dev
    lib.host_register(wires)                        # the Go shim's b200g16_host_register path
    try:
        with lib.Group(devices) as g:
            pk = g.pk_upload(L, N, pts["a"], pts["b"], pts["k"], pts["z"], pts["b2"], g1s[0], g1s[1], g1s[2], g2s[0], g2s[1],
                             zeros, zeros, k_skip, precompute=True)
            for it in range(3):
                got, h = g.prove(pk, wires, a, b, c, r, s, want_h=(it == 0), log2_domain=L)
                assert synth.check_proof(got, exp) == [], f"pass {it}"
                if it == 0:
                    assert np.array_equal(h, h_exp)
            g.pk_free(pk)
    finally:
        lib.host_unregister(wires)


# ---- one PROCESS per rank: the CUDA-IPC path of sharded.DistributedH, two ranks sharing GPU 0 over a gloo group
def _dist_h_worker(rank, world, port, L, q):
    import os

    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = False
    try:
        from gnark_whir_b200 import sharded
        torch.cuda.set_device(0)
        N = 1 << L
        M = N // world
        rs = np.random.Generator(np.random.PCG64(77))                    # same inputs on every rank
        a, b, c = synth.rand_fr(rs, N), synth.rand_fr(rs, N), synth.rand_fr(rs, N)
        a[N - 37:] = 0                                                    # zero padding reaches into the last slice
        want = cport.compute_h(a, b, c, L, 2)
        with lib.Context(0) as c0:
            dh = sharded.DistributedH(c0, L)
            got = []
            for it in range(2):                                           # twice: buffers and events are reused
                sl = [torch.from_numpy(v[rank * M:(rank + 1) * M].view(np.int64).copy()).cuda() for v in (a, b, c)]
                dh.load(*sl)
                d_h = dh.run()
                assert d_h + 32 * rank * M == dh.slice_ptr(0)
                from cuda import cudart
                host = np.empty((M, 4), dtype=np.uint64)
                err, = cudart.cudaMemcpy(host.ctypes.data, dh.slice_ptr(0), 32 * M, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost)
                got.append(int(err) == 0 and np.array_equal(host, want[rank * M:(rank + 1) * M]))
                dist.barrier()
            dh.close()
            ok = all(got)
    finally:
        q.put((rank, ok))
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_distributed_h_over_cuda_ipc_two_processes(world):
    import socket

    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mctx = mp.get_context("spawn")
    q = mctx.Queue()
    procs = [mctx.Process(target=_dist_h_worker, args=(r, world, port, 14, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
    assert res == [(r, True) for r in range(world)]
