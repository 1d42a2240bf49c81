"""Setup -> Prove -> Verify -> wire formats on one B200, through the host mirror of gnark's groth16 API
(gnark_whir_b200/groth16.py over the C-ABI).  Mirrors the three calls of the reference's flow
(/root/reference/mt.go:448,496,497) on a synthetic byte/bit-heavy R1CS with one BSB22 commitment.

    python examples/prove_verify.py [nb_constraints]
"""
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnark_whir_b200 import groth16, lib  # noqa: E402


def synthetic_circuit(nb_constraints, nb_public, rng):
    """(lc) * (lc) = fresh wire, 70% small values — the shape of a verifier circuit's witness."""
    R = groth16.R_MOD
    w = [1] + [rng.randrange(256) for _ in range(nb_public - 1)] + [rng.randrange(R) for _ in range(4)]
    commitment_wire = len(w)
    w.append(0)
    cons = []
    pick = lambda: rng.randrange(2) if rng.random() < 0.7 else rng.randrange(R)
    for _ in range(nb_constraints - 1):
        L = [(rng.randrange(len(w)), pick()) for _ in range(2)]
        Rr = [(rng.randrange(len(w)), pick()) for _ in range(2)]
        L = [(i, c) for i, c in L if i != commitment_wire] or [(0, 1)]
        Rr = [(i, c) for i, c in Rr if i != commitment_wire] or [(0, 1)]
        cons.append((L, Rr, [(len(w), 1)]))
        w.append(groth16._lc(L, w) * groth16._lc(Rr, w) % R)
    cons.append(([(commitment_wire, 1)], [(1, 1)], [(len(w), 1)]))      # uses the challenge: solved after the commitment
    w.append(0)
    r1cs = groth16.R1CS(len(w), nb_public, cons, private_committed=[nb_public, nb_public + 1], public_committed=[1],
                        commitment_wire=commitment_wire)
    return r1cs, w


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    rng = random.Random(1)
    r1cs, witness = synthetic_circuit(n, 4, rng)
    with lib.Context(0) as ctx:
        t0 = time.perf_counter()
        pk, vk = groth16.Setup(ctx, r1cs)
        t1 = time.perf_counter()

        def resolve(w):                                    # what the solver does once the challenge is known
            L, Rr, O = r1cs.constraints[-1]
            w[O[0][0]] = groth16._lc(L, w) * groth16._lc(Rr, w) % groth16.R_MOD
        proof = groth16.Prove(ctx, r1cs, pk, witness, resolve=resolve)
        t2 = time.perf_counter()
        public = proof.debug["witness"][1:r1cs.nb_public]
        groth16.Verify(ctx, proof, vk, public)
        t3 = time.perf_counter()
        blob = groth16.proof_write_to(ctx, proof)
        vk_blob = groth16.vk_write_to(ctx, vk)
        groth16.Verify(ctx, groth16.proof_read_from(ctx, blob), groth16.vk_read_from(ctx, vk_blob), public)
        print(f"{n} constraints: setup {1e3 * (t1 - t0):.1f} ms (python host loops dominate), prove {1e3 * (t2 - t1):.1f} ms, "
              f"verify {1e3 * (t3 - t2):.1f} ms; proof {len(blob)} bytes, vk {len(vk_blob)} bytes; round trip verified")
        pk.free()


if __name__ == "__main__":
    main()
