"""Regenerates tests/golden/*.json from the python oracle (run from the repo root).

The reference ships no vectors; these fixtures freeze the oracle's outputs so that a later
change to the oracle cannot silently move the target the CUDA path is compared against.
Keccak cases are cross-checked against hashlib before being written."""
import hashlib
import json
import os
import random
import sys

sys.path.insert(0, ".")
from oracle import groth16 as og  # noqa: E402
from oracle import keccak as ok  # noqa: E402
from oracle.bn254 import R  # noqa: E402

OUT = os.path.join("tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    rnd = random.Random(20261018)
    # hashlib pin first
    for n in (0, 1, 136, 200):
        m = bytes(rnd.randrange(256) for _ in range(n))
        assert ok.sha3_like(m, 136, 0x06, 32) == hashlib.sha3_256(m).digest()
    sponge = []
    for n_in, n_out in ((0, 32), (1, 32), (8, 8), (64, 32), (135, 32), (136, 32), (137, 32), (272, 32),
                        (512, 32), (512, 200), (33, 137)):
        m = bytes(rnd.randrange(256) for _ in range(n_in))
        sponge.append({"in": m.hex(), "out_len": n_out, "out": ok.sponge_hash(m, n_out).hex()})
    kf = []
    for st in ([0] * 25, [rnd.randrange(1 << 64) for _ in range(25)]):
        kf.append({"in": st, "out": ok.keccak_f(st)})
    with open(os.path.join(OUT, "keccak_sponge.json"), "w") as f:
        json.dump({"sponge": sponge, "keccak_f": kf}, f, indent=1)

    seed, nc, npub = 424242, 12, 3
    rng = random.Random(seed)
    r1cs, w = og.synthetic_r1cs(nc, npub, rng, with_commitment=False)
    tw = og.ToxicWaste(*[rng.randrange(1, R) for _ in range(5)], sigma=rng.randrange(1, R))
    pk, vk = og.setup(r1cs, tw)
    r, s = rng.randrange(R), rng.randrange(R)
    proof, aux = og.prove(r1cs, pk, w, r, s)
    assert og.verify(proof, vk, w[:npub])
    with open(os.path.join(OUT, "groth16_small.json"), "w") as f:
        json.dump({"seed": seed, "nb_constraints": nc, "nb_public": npub,
                   "Ar": [hex(v) for v in proof.Ar], "Bs": [hex(v) for c in proof.Bs for v in c],
                   "Krs": [hex(v) for v in proof.Krs], "h": [hex(v) for v in aux["h"]]}, f, indent=1)
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
