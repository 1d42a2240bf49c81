"""CPU tests of the ProveKit input readers (gnark_whir_b200/provekit.py; the reference parses the same files at
/root/reference/main.go:94-150 and expands them at mt.go:229-304, 358-402) and of the ProvingKey / Domain wire layout
of the oracle.  The reference ships no sample files, so fixtures are built here: the layout is arkworks'
CanonicalSerialize as documented in provekit.py, the decoding rules are the reference's own Go code restated in
oracle/keccak.py (prefix_decode_paths) and compared with the product's implementation."""
import base64
import json
import random

import numpy as np
import pytest

from gnark_whir_b200 import provekit as pkit
from oracle import keccak as ok
from oracle.bn254 import R


def _digest(rng):
    return bytes(rng.randrange(256) for _ in range(32))


def _element(rng, n_leaves, height, leaf_len):
    """A multipath the way arkworks compresses it: each path stored as (shared prefix length with the previous path,
    remaining suffix)."""
    full = [[_digest(rng) for _ in range(height)] for _ in range(n_leaves)]
    for j in range(1, n_leaves):                       # make neighbours share a random-length prefix
        k = rng.randrange(height + 1)
        full[j][:k] = full[j - 1][:k]
    prefix, suffixes = [0], [full[0]]
    for j in range(1, n_leaves):
        k = 0
        while k < height and full[j][k] == full[j - 1][k]:
            k += 1
        if k == height:                                # identical path: arkworks still stores a (possibly empty) suffix
            k = rng.randrange(height)
        prefix.append(k)
        suffixes.append(full[j][k:])
    mp = pkit.MultiPath([_digest(rng) for _ in range(n_leaves)], np.array(prefix, np.uint64), suffixes,
                        np.array([rng.randrange(1 << (height + 1)) for _ in range(n_leaves)], np.uint64))
    leaves = [[rng.randrange(1 << 256) for _ in range(leaf_len)] for _ in range(n_leaves)]
    return pkit.ProofElement(mp, leaves), full


def test_ark_proof_round_trip_and_layout():
    rng = random.Random(7)
    e0, _ = _element(rng, 5, 6, 4)
    e1, _ = _element(rng, 3, 4, 2)
    proof = pkit.ProofObject([e0], [e1, e0], [rng.randrange(R) for _ in range(3)])
    data = pkit.write_proof(proof)
    # layout spot checks: u64 little-endian vector lengths, 32-byte digests, little-endian field elements
    assert data[:8] == (1).to_bytes(8, "little") and data[8:16] == (5).to_bytes(8, "little")
    assert data[16:48] == e0.A.LeafSiblingHashes[0]
    assert data[-32:] == proof.StatementValuesAtRandomPoint[-1].to_bytes(32, "little")
    back = pkit.read_proof(data)
    assert pkit.write_proof(back) == data
    assert back.MerklePaths[0].B == e1.B and list(back.FirstRoundPaths[0].A.LeafIndexes) == list(e0.A.LeafIndexes)
    for cut in (0, 7, 100, len(data) - 1):
        with pytest.raises(ValueError):
            pkit.read_proof(data[:cut])
    with pytest.raises(ValueError):
        pkit.read_proof(data + b"\x00")
    with pytest.raises(ValueError):                    # a length field larger than the input
        pkit.read_proof((1 << 40).to_bytes(8, "little") + data[8:])


def test_parse_paths_object_matches_the_reference_rules():
    rng = random.Random(11)
    e, full = _element(rng, 9, 7, 3)
    dec = pkit.parse_paths_object([e])[0]
    # utilities.PrefixDecodePath + utilities.Reverse, as restated by the oracle
    want = ok.prefix_decode_paths(e.A.AuthPathsSuffixes, [int(x) for x in e.A.AuthPathsPrefixLengths])
    assert dec.AuthPaths == want and [p[::-1] for p in dec.AuthPaths] == full
    assert all(len(p) == 7 for p in dec.AuthPaths)
    assert dec.Leaves == [[v % R for v in leaf] for leaf in e.B]
    leaves, sib, auth, idx = pkit.merkle_batch(dec)
    assert leaves.shape == (9, 96) and sib.shape == (9, 32) and auth.shape == (9, 7, 32) and idx.dtype == np.uint64
    assert bytes(leaves[2][32:64]) == dec.Leaves[2][1].to_bytes(32, "little")
    assert bytes(auth[4][0]) == dec.AuthPaths[4][0] and bytes(sib[8]) == e.A.LeafSiblingHashes[8]
    bad = pkit.ProofElement(pkit.MultiPath(e.A.LeafSiblingHashes, e.A.AuthPathsPrefixLengths, e.A.AuthPathsSuffixes[:-1],
                                           e.A.LeafIndexes), e.B)
    with pytest.raises(ValueError):
        pkit.parse_paths_object([bad])


def test_params_and_r1cs_json():
    cfg = {"log_num_constraints": 20, "n_rounds": 3, "n_vars": 20, "folding_factor": [4, 4, 4], "ood_samples": [2, 2, 2],
           "num_queries": [80, 50, 30], "pow_bits": [10, 8, 6], "final_queries": 20, "final_pow_bits": 4,
           "final_folding_pow_bits": 2, "domain_generator": "5", "rate": 1, "io_pattern": "x", "transcript_len": 3,
           "transcript": base64.b64encode(b"\x01\x02\x03").decode(), "statement_evaluations": ["7"], "ignored": 1}
    got = pkit.read_params(json.dumps(cfg))
    assert got["num_queries"] == [80, 50, 30] and got["transcript"] == b"\x01\x02\x03" and "ignored" not in got
    assert pkit.read_params("{}")["folding_factor"] == [] and pkit.read_params("{}")["n_rounds"] == 0
    rng = random.Random(13)
    interner = [rng.randrange(1 << 256) for _ in range(6)]
    hexval = ((6).to_bytes(8, "little") + b"".join(v.to_bytes(32, "little") for v in interner)).hex()
    mat = {"rows": 3, "cols": 4, "row_indices": [0, 2, 2], "col_indices": [1, 3, 0, 2], "values": [5, 0, 3, 3]}
    r1 = pkit.read_r1cs(json.dumps({"public_inputs": 1, "witnesses": 4, "constraints": 3, "interner": {"values": hexval},
                                    "a": mat, "b": mat, "c": mat}))
    assert r1.interner == interner and r1.constraints == 3
    # mt.go:358-372: row i owns [row_indices[i], row_indices[i+1]) and the last row runs to the end
    assert pkit.matrix_cells(r1.a, r1.interner) == [(0, 1, interner[5] % R), (0, 3, interner[0] % R),
                                                    (2, 0, interner[3] % R), (2, 2, interner[3] % R)]
    broken = dict(mat, values=[5, 0, 3, 9])
    with pytest.raises(ValueError):
        pkit.read_r1cs(json.dumps({"interner": {"values": hexval}, "a": broken, "b": mat, "c": mat}))
