"""GPU parity: groth16 Setup / Prove through the host mirror + C-ABI vs the oracle's restatement
of gnark v0.11.0 (reference call sites mt.go:448,496,497).  With toxic waste and r, s fixed,
every pk/vk element, every intermediate MSM output, H, and the final proof must be identical;
the proof must verify under the oracle's independent pairing and equal the closed form."""
import random

import numpy as np
import pytest

from gnark_whir_b200 import groth16 as g16
from oracle import bn254 as bn
from oracle import groth16 as og
from oracle.bn254 import R

pytestmark = pytest.mark.gpu


def _case(seed, nb_constraints, nb_public, with_commitment):
    rng = random.Random(seed)
    r1cs, w = og.synthetic_r1cs(nb_constraints, nb_public, rng, with_commitment=with_commitment)
    tw = og.ToxicWaste(*[rng.randrange(1, R) for _ in range(5)], sigma=rng.randrange(1, R))
    r, s = rng.randrange(R), rng.randrange(R)
    return r1cs, w, tw, r, s


@pytest.mark.parametrize("nb_constraints,nb_public,with_commitment",
                         [(1, 1, False), (7, 2, False), (24, 3, False), (24, 3, True), (100, 5, True), (257, 9, False)])
def test_setup_prove_match_oracle_and_verify(ctx, nb_constraints, nb_public, with_commitment):
    r1cs, w, tw, r, s = _case(nb_constraints * 7 + nb_public, nb_constraints, nb_public, with_commitment)
    opk, ovk = og.setup(r1cs, tw)
    pk, vk = g16.Setup(ctx, r1cs, g16.ToxicWaste(tw.tau, tw.alpha, tw.beta, tw.gamma, tw.delta, tw.sigma))
    try:
        # ---- Setup parity: every element of pk / vk
        assert pk.log2_domain == opk.domain.logn
        assert np.array_equal(pk.G1_A, bn.g1_to_array(opk.A))
        assert np.array_equal(pk.G1_B, bn.g1_to_array(opk.B))
        assert np.array_equal(pk.G1_Z, bn.g1_to_array(opk.Z))
        assert np.array_equal(pk.G1_K, bn.g1_to_array(opk.K))
        assert np.array_equal(pk.G2_B, bn.g2_to_array(opk.B2))
        assert np.array_equal(pk.G1_Alpha, bn.g1_to_array([opk.alpha1])[0])
        assert np.array_equal(pk.G1_Beta, bn.g1_to_array([opk.beta1])[0])
        assert np.array_equal(pk.G1_Delta, bn.g1_to_array([opk.delta1])[0])
        assert np.array_equal(pk.G2_Beta, bn.g2_to_array([opk.beta2])[0])
        assert np.array_equal(pk.G2_Delta, bn.g2_to_array([opk.delta2])[0])
        assert list(pk.InfinityA.astype(bool)) == opk.infinity_a
        assert list(pk.InfinityB.astype(bool)) == opk.infinity_b
        assert np.array_equal(vk.G1_K, bn.g1_to_array(ovk.K))
        assert np.array_equal(vk.G2_Gamma, bn.g2_to_array([ovk.gamma2])[0])
        if with_commitment:
            assert np.array_equal(pk.CommitmentKeys[0].Basis, bn.g1_to_array(opk.ped_basis))
            assert np.array_equal(pk.CommitmentKeys[0].BasisExpSigma, bn.g1_to_array(opk.ped_basis_exp_sigma))
            assert np.array_equal(vk.PedersenGSigmaNeg, bn.g2_to_array([ovk.ped_g_sigma_neg])[0])

        # ---- Prove parity
        ow = list(w)
        ocom = og.finalize_witness(r1cs, opk, ow)

        def resolve(wit):                      # stands in for the solver finishing after the hint
            L, Rr, O = r1cs.constraints[-1]
            wit[O[0][0]] = og.lc_eval(L, wit) * og.lc_eval(Rr, wit) % R
        proof = g16.Prove(ctx, r1cs, pk, w, r=r, s=s, resolve=resolve if with_commitment else None, want_h=True)
        assert proof.debug["witness"] == ow
        oproof, aux = og.prove(r1cs, opk, ow, r, s, commitment=ocom)
        n = opk.domain.n
        assert np.array_equal(proof.debug["h"], bn.fr_to_mont_array(aux["h"]))
        wa = [ow[i] for i in range(r1cs.nb_wires) if not opk.infinity_a[i]]
        wb = [ow[i] for i in range(r1cs.nb_wires) if not opk.infinity_b[i]]
        wk = [ow[i] for i in opk.k_wires]
        assert np.array_equal(proof.debug["msm_a"], bn.g1_to_array([bn.g1_msm(opk.A, wa)])[0])
        assert np.array_equal(proof.debug["msm_b1"], bn.g1_to_array([bn.g1_msm(opk.B, wb)])[0])
        assert np.array_equal(proof.debug["msm_k"], bn.g1_to_array([bn.g1_msm(opk.K, wk)])[0])
        assert np.array_equal(proof.debug["msm_z"], bn.g1_to_array([aux["krs2"]])[0])
        assert np.array_equal(proof.debug["msm_b2"], bn.g2_to_array([bn.g2_msm(opk.B2, wb)])[0])
        assert np.array_equal(proof.debug["bs1"], bn.g1_to_array([aux["bs1"]])[0])
        assert np.array_equal(proof.Ar, bn.g1_to_array([oproof.Ar])[0])
        assert np.array_equal(proof.Bs, bn.g2_to_array([oproof.Bs])[0])
        assert np.array_equal(proof.Krs, bn.g1_to_array([oproof.Krs])[0])
        if with_commitment:
            assert np.array_equal(proof.Commitments[0], bn.g1_to_array([oproof.commitments[0]])[0])
            assert np.array_equal(proof.CommitmentPok, bn.g1_to_array([oproof.commitment_pok])[0])

        # ---- the GPU proof verifies (independent pairing) and equals the closed form
        gp = og.Proof(bn.g1_from_array(proof.Ar)[0], bn.g2_from_array(proof.Bs)[0], bn.g1_from_array(proof.Krs)[0],
                      [bn.g1_from_array(c)[0] for c in proof.Commitments],
                      bn.g1_from_array(proof.CommitmentPok)[0] if with_commitment else None)
        assert og.verify(gp, ovk, ow[:r1cs.nb_public])
        assert og.closed_form_proof(r1cs, tw, ow, r, s, aux["h"]) == (gp.Ar, gp.Bs, gp.Krs)
        bad = list(ow[:r1cs.nb_public])
        if len(bad) > 1:
            bad[1] = (bad[1] + 1) % R
            assert not og.verify(gp, ovk, bad)
    finally:
        pk.free()


def test_prove_rejects_mismatched_inputs(ctx):
    from gnark_whir_b200 import lib
    r1cs, w, tw, r, s = _case(3, 8, 2, False)
    pk, _ = g16.Setup(ctx, r1cs, g16.ToxicWaste(tw.tau, tw.alpha, tw.beta, tw.gamma, tw.delta, tw.sigma))
    try:
        h = pk.device_handle(ctx)
        a, b, c = g16.solve_abc(r1cs, w)
        with pytest.raises(lib.B200Error):      # wrong witness length
            ctx.prove(h, g16.fr_array(w[:-1]), g16.fr_array(a), g16.fr_array(b), g16.fr_array(c),
                      g16.fr_array([r])[0], g16.fr_array([s])[0])
        with pytest.raises(lib.B200Error):      # more constraints than the domain
            big = g16.fr_array(a * 3)
            ctx.prove(h, g16.fr_array(w), big, big, big, g16.fr_array([r])[0], g16.fr_array([s])[0])
    finally:
        pk.free()


def test_prove_with_window_tables_is_identical(ctx):
    """pk uploaded with precompute=1 (window tables on A, B1, K, Z, B2): every MSM output and the
    proof must equal the table-free prove bit for bit."""
    r1cs, w, tw, r, s = _case(91, 300, 4, False)
    pk, _ = g16.Setup(ctx, r1cs, g16.ToxicWaste(tw.tau, tw.alpha, tw.beta, tw.gamma, tw.delta, tw.sigma))
    try:
        a, b, c = g16.solve_abc(r1cs, w)
        args = (g16.fr_array(w), g16.fr_array(a), g16.fr_array(b), g16.fr_array(c), g16.fr_array([r])[0], g16.fr_array([s])[0])
        plain, _ = ctx.prove(pk.device_handle(ctx), *args)
        pk.free()
        tabled, _ = ctx.prove(pk.device_handle(ctx, precompute=True), *args)
        for k in plain:
            assert np.array_equal(plain[k], tabled[k]), k
    finally:
        pk.free()


def test_staged_compute_h_and_prove_h_match_prove(ctx):
    """The multi-GPU prove runs computeH as separate stages (b200g16_ntt_dev x2 per vector on three ranks,
    broadcast, b200g16_h_pointwise_dev, last transform) and then b200g16_prove_h_dev: on one GPU the stages
    must reproduce b200g16_compute_h / b200g16_prove bit for bit."""
    import torch
    from gnark_whir_b200 import lib, sharded
    r1cs, w, tw, r, s = _case(123, 200, 3, False)
    pk, _ = g16.Setup(ctx, r1cs, g16.ToxicWaste(tw.tau, tw.alpha, tw.beta, tw.gamma, tw.delta, tw.sigma))
    try:
        a, b, c = g16.solve_abc(r1cs, w)
        L, N = pk.log2_domain, 1 << pk.log2_domain
        fr = [g16.fr_array(w), g16.fr_array(a), g16.fr_array(b), g16.fr_array(c), g16.fr_array([r])[0], g16.fr_array([s])[0]]
        full, h = ctx.prove(pk.device_handle(ctx), *fr, want_h=True, log2_domain=L)

        def dev(x):
            t = torch.zeros((N, 4), dtype=torch.int64, device="cuda")
            t[:x.shape[0]] = torch.from_numpy(x.view(np.int64)).cuda()
            return t
        ta, tb, tc = dev(fr[1]), dev(fr[2]), dev(fr[3])
        for t in (ta, tb, tc):
            ctx.ntt_dev(t.data_ptr(), L, inverse=True, decimation=lib.DIF)
            ctx.ntt_dev(t.data_ptr(), L, coset=True, decimation=lib.DIT)
        ctx.h_pointwise_dev(ta.data_ptr(), tb.data_ptr(), tc.data_ptr(), L)
        ctx.ntt_dev(ta.data_ptr(), L, inverse=True, coset=True, decimation=lib.DIF)
        assert np.array_equal(ta.cpu().numpy().view(np.uint64), h)
        # world of one: compute_h_distributed is the plain library call
        ta2, tb2, tc2 = dev(fr[1]), dev(fr[2]), dev(fr[3])
        sharded.compute_h_distributed(ctx, ta2, tb2, tc2, L)
        assert torch.equal(ta, ta2)
        wires = torch.from_numpy(fr[0].view(np.int64)).cuda()
        staged = ctx.prove_h_dev(pk.device_handle(ctx), wires.data_ptr(), ta.data_ptr(), fr[4], fr[5])
        for k in full:
            assert np.array_equal(full[k], staged[k]), k
        # the two-halves form: witness MSMs enqueued first, h supplied later (other library calls in between)
        ctx.prove_begin_dev(pk.device_handle(ctx), wires.data_ptr())
        tb3 = dev(fr[2])
        ctx.ntt_dev(tb3.data_ptr(), L, inverse=True, decimation=lib.DIF)      # queues behind the MSMs
        halves = ctx.prove_end_dev(pk.device_handle(ctx), ta.data_ptr(), fr[4], fr[5])
        for k in full:
            assert np.array_equal(full[k], halves[k]), k
        with pytest.raises(lib.B200Error):                                    # end without begin
            ctx.prove_end_dev(pk.device_handle(ctx), ta.data_ptr(), fr[4], fr[5])
    finally:
        pk.free()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_prove_equals_single_gpu_prove(ctx, world):
    """Point-range shards of the proving key (config 5): the shards' partial MSM sums, added and
    finished on the host, must give the identical proof.  The ranks are emulated one after the
    other on this GPU (same library calls a multi-process run makes; the exchange itself is
    covered over gloo in test_sharded_cpu.py and over NCCL by bench.py --gpus N)."""
    from gnark_whir_b200 import sharded
    r1cs, w, tw, r, s = _case(77, 60, 4, False)
    pk, _ = g16.Setup(ctx, r1cs, g16.ToxicWaste(tw.tau, tw.alpha, tw.beta, tw.gamma, tw.delta, tw.sigma))
    try:
        a, b, c = g16.solve_abc(r1cs, w)
        args = (g16.fr_array(w), g16.fr_array(a), g16.fr_array(b), g16.fr_array(c), g16.fr_array([r])[0], g16.fr_array([s])[0])
        full, _ = ctx.prove(pk.device_handle(ctx), *args)
        packed, shard0 = [], None
        for rank in range(world):
            h = sharded.upload_pk_shard(ctx, pk, rank, world)
            part, _ = ctx.prove(h, *args)
            assert not part["ar"].any()                     # a shard never claims a finished proof
            packed.append(sharded.pack_partials(part))
            if rank == 0:
                shard0 = h
            else:
                ctx.pk_free(h)
        sums = sharded.sum_partials(packed)
        for k in ("msm_a", "msm_b1", "msm_k", "msm_z", "msm_b2"):
            assert np.array_equal(sums[k], full[k]), k
        fin = ctx.prove_finish(shard0, sums["msm_a"], sums["msm_b1"], sums["msm_k"], sums["msm_z"], sums["msm_b2"], args[4], args[5])
        ctx.pk_free(shard0)
        for k in ("ar", "bs", "krs", "bs1"):
            assert np.array_equal(fin[k], full[k]), k
        # world of one through the convenience wrapper
        one = sharded.prove_sharded(ctx, sharded.upload_pk_shard(ctx, pk, 0, 1), *args)
        assert np.array_equal(one["krs"], full["krs"])
    finally:
        pk.free()
