"""Small pass over every kernel family for compute-sanitizer (memcheck / racecheck); run under gpurun:
   compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import random
import sys

import numpy as np

sys.path.insert(0, ".")
from gnark_whir_b200 import groth16 as g16  # noqa: E402
from gnark_whir_b200 import lib  # noqa: E402
from oracle import bn254 as bn  # noqa: E402
from oracle import groth16 as og  # noqa: E402
from oracle.bn254 import R  # noqa: E402

rng = random.Random(3)
rs = np.random.Generator(np.random.PCG64(3))


def rand_fr(n):
    a = rs.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    return a


with lib.Context(0) as ctx:
    n = 3000
    ks, sc = rand_fr(n), rand_fr(n)
    sc[:500] = 0
    sc[500:1500, 1:] = 0
    sc[500:1500, 0] = bn.fr_to_mont_array([1])[0][0]          # a heavy bucket (scalar 1), exercises the merge tree
    sc[500:1500] = bn.fr_to_mont_array([1])[0]
    for group in (1, 2):
        gen = g16.g1_point(g16.G1_GEN) if group == 1 else g16.g2_point(g16.G2_GEN)
        bases = ctx.fixed_base_mul(gen, ks, group=group, resident=True)
        plain = ctx.msm(bases, sc)
        bases.precompute(9)
        assert np.array_equal(ctx.msm(bases, sc), plain)
        bases.free()
    a = rand_fr(1 << 12)
    back = ctx.ntt(ctx.ntt(a, coset=True), inverse=True, coset=True, decimation=lib.DIT)
    assert np.array_equal(a, back)
    ctx.compute_h(rand_fr(1000), rand_fr(1000), rand_fr(1000), 10)
    ctx.keccak_f_batch(rs.integers(0, 1 << 63, size=(70, 25), dtype=np.uint64))
    ctx.keccak_sponge_batch(rs.integers(0, 256, size=(33, 300), dtype=np.uint8), 64)
    r1cs, w = og.synthetic_r1cs(40, 3, rng, with_commitment=True)
    pk, vk = g16.Setup(ctx, r1cs)

    def resolve(wit):
        L, Rr, O = r1cs.constraints[-1]
        wit[O[0][0]] = og.lc_eval(L, wit) * og.lc_eval(Rr, wit) % R
    proof = g16.Prove(ctx, r1cs, pk, w, resolve=resolve)
    g16.Verify(ctx, proof, vk, proof.debug["witness"][1:r1cs.nb_public])
    data = g16.proof_write_to(ctx, proof)
    g16.proof_read_from(ctx, data)
    pk.free()
print("sanitize smoke ok")
