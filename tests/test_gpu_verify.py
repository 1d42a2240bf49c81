"""GPU parity: groth16.Verify / PairingCheck / Pair through the C-ABI (reference call site mt.go:497)
against the oracle's independent pairing and verifier."""
import random

import numpy as np
import pytest

from gnark_whir_b200 import groth16 as g16
from gnark_whir_b200 import lib
from oracle import bn254 as bn
from oracle import groth16 as og
from oracle.bn254 import R
from tests.test_pairing_cpu import COFACTOR, gt_from_gnark_layout

pytestmark = pytest.mark.gpu


def test_pairing_check_bilinearity(ctx):
    rng = random.Random(21)
    ka, kb = rng.randrange(1, R), rng.randrange(1, R)
    Pa, Qb = bn.g1_mul(bn.G1_GEN, ka), bn.g2_mul(bn.G2_GEN, kb)
    g2 = bn.g2_to_array([Qb, bn.G2_GEN])
    assert ctx.pairing_check(bn.g1_to_array([Pa, bn.g1_neg(bn.g1_mul(bn.G1_GEN, ka * kb % R))]), g2)
    assert not ctx.pairing_check(bn.g1_to_array([Pa, bn.g1_neg(bn.g1_mul(bn.G1_GEN, (ka * kb + 1) % R))]), g2)
    # empty product and products with infinity are 1
    assert ctx.pairing_check(np.zeros((0, 8), np.uint64), np.zeros((0, 16), np.uint64))
    assert ctx.pairing_check(np.zeros((1, 8), np.uint64), bn.g2_to_array([Qb]))
    # more pairs than threads in the CTA: prod_i e(a_i G, Q) * e(-(sum a_i) G, Q) == 1
    ks = [rng.randrange(1, R) for _ in range(40)]
    pts = [bn.g1_mul(bn.G1_GEN, k) for k in ks] + [bn.g1_neg(bn.g1_mul(bn.G1_GEN, sum(ks) % R))]
    assert ctx.pairing_check(bn.g1_to_array(pts), bn.g2_to_array([Qb] * 41))
    # a G1 point off the curve is an error, not "false"
    off = bn.g1_to_array([(Pa[0], (Pa[1] + 1) % bn.P)])
    with pytest.raises(lib.B200Error):
        ctx.pairing_check(off, bn.g2_to_array([Qb]))


def test_pair_equals_oracle_pairing_with_gnark_cofactor(ctx):
    rng = random.Random(22)
    Pa, Qb = bn.g1_mul(bn.G1_GEN, rng.randrange(1, R)), bn.g2_mul(bn.G2_GEN, rng.randrange(1, R))
    gt = ctx.pair(bn.g1_to_array([Pa]), bn.g2_to_array([Qb]))
    assert gt_from_gnark_layout(gt) == bn.f12_pow(bn.pairing(Pa, Qb), COFACTOR)


@pytest.mark.parametrize("nb_constraints,nb_public,with_commitment", [(7, 1, False), (24, 3, False), (40, 4, True)])
def test_verify_accepts_gpu_proofs_and_rejects_tampering(ctx, nb_constraints, nb_public, with_commitment):
    rng = random.Random(nb_constraints)
    r1cs, w = og.synthetic_r1cs(nb_constraints, nb_public, rng, with_commitment=with_commitment)
    tw = og.ToxicWaste(*[rng.randrange(1, R) for _ in range(5)], sigma=rng.randrange(1, R))
    pk, vk = g16.Setup(ctx, r1cs, g16.ToxicWaste(tw.tau, tw.alpha, tw.beta, tw.gamma, tw.delta, tw.sigma))
    try:
        def resolve(wit):
            L, Rr, O = r1cs.constraints[-1]
            wit[O[0][0]] = og.lc_eval(L, wit) * og.lc_eval(Rr, wit) % R
        proof = g16.Prove(ctx, r1cs, pk, w, r=rng.randrange(R), s=rng.randrange(R),
                          resolve=resolve if with_commitment else None)
        wit = proof.debug["witness"]
        public = wit[1:r1cs.nb_public]                      # witness.Public(): without the one wire
        g16.Verify(ctx, proof, vk, public)                     # raises if invalid
        # the oracle's verifier agrees on the same proof
        _, ovk = og.setup(r1cs, tw)
        gp = og.Proof(bn.g1_from_array(proof.Ar)[0], bn.g2_from_array(proof.Bs)[0], bn.g1_from_array(proof.Krs)[0],
                      [bn.g1_from_array(c)[0] for c in proof.Commitments],
                      bn.g1_from_array(proof.CommitmentPok)[0] if with_commitment else None)
        assert og.verify(gp, ovk, wit[:r1cs.nb_public])
        # tampered public input
        if public:
            bad = list(public)
            bad[0] = (bad[0] + 1) % R
            with pytest.raises(ValueError):
                g16.Verify(ctx, proof, vk, bad)
        # tampered proof element (still a valid curve point)
        forged = g16.Proof(bn.g1_to_array([bn.g1_mul(bn.g1_from_array(proof.Ar)[0], 2)])[0], proof.Krs, proof.Bs,
                           proof.Commitments, proof.CommitmentPok)
        with pytest.raises(ValueError):
            g16.Verify(ctx, forged, vk, public)
        if with_commitment:                                   # forged proof of knowledge
            forged = g16.Proof(proof.Ar, proof.Krs, proof.Bs, proof.Commitments,
                               bn.g1_to_array([bn.g1_mul(bn.g1_from_array(proof.CommitmentPok)[0], 3)])[0])
            with pytest.raises(ValueError):
                g16.Verify(ctx, forged, vk, public)
        # wrong witness size is an error (gnark: "invalid witness size")
        with pytest.raises(lib.B200Error):
            g16.Verify(ctx, proof, vk, list(public) + [1])
    finally:
        pk.free()
