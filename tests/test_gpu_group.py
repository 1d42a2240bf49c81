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
    out = [[0], [0, 0], [0, 0, 0], [0] * 5]
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


def test_group_prove_2p18_known_dlog_key(ctx):
    """2^18 constraints, window tables, 3 shards: every MSM output, h and the proof against the closed forms."""
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
    devices = [0, 1, 2] if torch.cuda.device_count() >= 3 else [0, 0, 0]
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
