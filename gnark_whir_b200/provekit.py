"""ProveKit inputs of the reference's `go run .` flow (/root/reference/main.go:15-150), read on the host:

  prover/proof   ark-serialize (CanonicalSerialize, uncompressed) of ProofObject      main.go:15-39, 94-105
  prover/params  JSON Config                                                          main.go:41-58, 107-118
  r1cs.json      JSON R1CS: three CSR matrices over an interned value table, the table itself hex-encoded
                 ark-serialize of Interner{Values []Fp256}                            main.go:60-90, 128-150

plus what `verify_circuit` derives from them before the circuit is built: the CSR -> cell expansion
(mt.go:358-402) and the Merkle multipath decoding (ParsePathsObject, mt.go:229-304; PrefixDecodePath,
utilities.go:67-78).  `merkle_batch` turns one decoded ProofElement into the flat arrays
b200g16_keccak_merkle_paths takes, so the recompute of every opened path of a round is one GPU call.

ark-serialize layout (arkworks CanonicalSerialize, mirrored by reilabs/go-ark-serialize; neither source is on disk —
go.mod pins the module — so this follows the published format): integers little-endian, u64 / usize as 8 bytes,
Vec<T> as a u64 length followed by the elements, [u8; 32] as 32 raw bytes, a prime-field element as its canonical
value in 32 little-endian bytes (== the four little-endian u64 limbs of Fp256, typeConverters.go:26-44), struct
fields in declaration order.
"""
from __future__ import annotations

import json
import struct
from dataclasses import dataclass, field

import numpy as np

R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001


# ---------------------------------------------------------------- ark-serialize primitives
class _Reader:
    def __init__(self, data):
        self.d, self.o = memoryview(bytes(data)), 0

    def take(self, n):
        if self.o + n > len(self.d):
            raise ValueError("ark-serialize: unexpected end of input")
        v = self.d[self.o:self.o + n]
        self.o += n
        return v

    def u64(self):
        return struct.unpack("<Q", self.take(8))[0]

    def vec(self, elem):
        n = self.u64()
        if n > len(self.d) - self.o:                 # every element takes at least one byte
            raise ValueError("ark-serialize: vector length exceeds the input")
        return [elem() for _ in range(n)]

    def digest(self):
        return bytes(self.take(32))

    def fp256(self):
        return int.from_bytes(self.take(32), "little")

    def u64s(self):
        n = self.u64()
        if 8 * n > len(self.d) - self.o:
            raise ValueError("ark-serialize: vector length exceeds the input")
        return np.frombuffer(self.take(8 * n), dtype="<u8").astype(np.uint64)


def _w_u64(v):
    return struct.pack("<Q", int(v))


def _w_vec(items, enc):
    return _w_u64(len(items)) + b"".join(enc(x) for x in items)


def _w_fp(v):
    return int(v).to_bytes(32, "little")


# ---------------------------------------------------------------- ProofObject (main.go:15-39)
@dataclass
class MultiPath:                       # arkworks crypto-primitives MultiPath, Digest = KeccakDigest
    LeafSiblingHashes: list            # [bytes32]
    AuthPathsPrefixLengths: np.ndarray
    AuthPathsSuffixes: list            # [[bytes32]]
    LeafIndexes: np.ndarray


@dataclass
class ProofElement:
    A: MultiPath
    B: list                            # [[int]]: the opened leaves, field elements


@dataclass
class ProofObject:
    FirstRoundPaths: list
    MerklePaths: list
    StatementValuesAtRandomPoint: list = field(default_factory=list)


def _read_multipath(r):
    return MultiPath(r.vec(r.digest), r.u64s(), r.vec(lambda: r.vec(r.digest)), r.u64s())


def _read_element(r):
    return ProofElement(_read_multipath(r), r.vec(lambda: r.vec(r.fp256)))


def read_proof(data):
    """go_ark_serialize.CanonicalDeserializeWithMode(proofFile, &proof, false, false)  (main.go:100-105)"""
    r = _Reader(data)
    p = ProofObject(r.vec(lambda: _read_element(r)), r.vec(lambda: _read_element(r)), r.vec(r.fp256))
    if r.o != len(r.d):
        raise ValueError(f"ark-serialize: {len(r.d) - r.o} trailing bytes after the proof")
    return p


def _write_multipath(m):
    return (_w_vec(m.LeafSiblingHashes, bytes) + _w_vec(list(m.AuthPathsPrefixLengths), _w_u64)
            + _w_vec(m.AuthPathsSuffixes, lambda s: _w_vec(s, bytes)) + _w_vec(list(m.LeafIndexes), _w_u64))


def _write_element(e):
    return _write_multipath(e.A) + _w_vec(e.B, lambda leaf: _w_vec(leaf, _w_fp))


def write_proof(p):
    return (_w_vec(p.FirstRoundPaths, _write_element) + _w_vec(p.MerklePaths, _write_element)
            + _w_vec(p.StatementValuesAtRandomPoint, _w_fp))


# ---------------------------------------------------------------- params (main.go:41-58)
CONFIG_KEYS = ("log_num_constraints", "n_rounds", "n_vars", "folding_factor", "ood_samples", "num_queries", "pow_bits",
               "final_queries", "final_pow_bits", "final_folding_pow_bits", "domain_generator", "rate", "io_pattern",
               "transcript", "transcript_len", "statement_evaluations")


def read_params(data):
    """json.Unmarshal(configFile, &config): unknown keys ignored, missing keys zero-valued like Go's decoder;
    `transcript` is a Go []byte, i.e. base64 in JSON."""
    import base64
    raw = json.loads(data)
    zero = {"folding_factor": [], "ood_samples": [], "num_queries": [], "pow_bits": [], "statement_evaluations": [],
            "domain_generator": "", "io_pattern": "", "transcript": b""}
    cfg = {k: raw.get(k, zero.get(k, 0)) for k in CONFIG_KEYS}
    if isinstance(cfg["transcript"], str):
        cfg["transcript"] = base64.b64decode(cfg["transcript"])
    elif isinstance(cfg["transcript"], list):
        cfg["transcript"] = bytes(cfg["transcript"])
    return cfg


# ---------------------------------------------------------------- r1cs.json (main.go:60-90, 128-150)
@dataclass
class SparseMatrix:
    rows: int
    cols: int
    row_indices: np.ndarray            # start offset of every row in col_indices / values
    col_indices: np.ndarray
    values: np.ndarray                 # indices into the interner


@dataclass
class InternedR1CS:
    public_inputs: int
    witnesses: int
    constraints: int
    interner: list                     # field elements (canonical ints)
    a: SparseMatrix
    b: SparseMatrix
    c: SparseMatrix


def read_interner(data):
    r = _Reader(data)
    vals = r.vec(r.fp256)
    if r.o != len(r.d):
        raise ValueError("ark-serialize: trailing bytes after the interner")
    return vals


def read_r1cs(data):
    raw = json.loads(data)

    def mat(m):
        u = lambda k: np.asarray(m.get(k, []), dtype=np.uint64)
        sm = SparseMatrix(int(m.get("rows", 0)), int(m.get("cols", 0)), u("row_indices"), u("col_indices"), u("values"))
        if len(sm.col_indices) != len(sm.values):
            raise ValueError("r1cs.json: col_indices and values differ in length")
        if len(sm.row_indices) and (np.diff(sm.row_indices.astype(np.int64)) < 0).any():
            raise ValueError("r1cs.json: row_indices must be non-decreasing")
        return sm
    interner = read_interner(bytes.fromhex(raw["interner"]["values"]))
    out = InternedR1CS(int(raw.get("public_inputs", 0)), int(raw.get("witnesses", 0)), int(raw.get("constraints", 0)),
                       interner, mat(raw["a"]), mat(raw["b"]), mat(raw["c"]))
    for m in (out.a, out.b, out.c):
        if len(m.values) and int(m.values.max()) >= len(interner):
            raise ValueError("r1cs.json: value index outside the interner")
    return out


def matrix_cells(m, interner):
    """mt.go:358-372: CSR -> (row, column, value mod r) per stored cell, in storage order."""
    nnz = len(m.values)
    starts = m.row_indices.astype(np.int64)
    rows = np.zeros(nnz, dtype=np.int64)
    if len(starts):
        ends = np.append(starts[1:], nnz)
        for i, (s, e) in enumerate(zip(starts, ends)):
            rows[s:e] = i
    table = [v % R_MOD for v in interner]
    return [(int(rows[j]), int(m.col_indices[j]), table[int(m.values[j])]) for j in range(nnz)]


# ---------------------------------------------------------------- Merkle multipaths (mt.go:229-304, utilities.go:58-78)
def prefix_decode_path(prev, prefix_len, suffix):
    """utilities.PrefixDecodePath"""
    return list(suffix) if prefix_len == 0 else list(prev[:prefix_len]) + list(suffix)


@dataclass
class DecodedPaths:
    AuthPaths: list                    # per leaf: tree-height digests, leaf side first (after utilities.Reverse)
    Leaves: list                       # per leaf: field elements mod r
    LeafSiblingHashes: list
    LeafIndexes: np.ndarray


def parse_paths_object(elements):
    """ParsePathsObject (mt.go:229-304) for every ProofElement."""
    out = []
    for e in elements:
        n = len(e.A.LeafIndexes)
        if n == 0:
            out.append(DecodedPaths([], [], [], np.zeros(0, np.uint64)))
            continue
        if not (len(e.A.AuthPathsSuffixes) == len(e.A.AuthPathsPrefixLengths) == len(e.A.LeafSiblingHashes) == len(e.B) == n):
            raise ValueError("multipath: vectors of different lengths")
        height = len(e.A.AuthPathsSuffixes[0])
        prev = list(e.A.AuthPathsSuffixes[0])
        paths = [prev[::-1]]
        for j in range(1, n):
            prev = prefix_decode_path(prev, int(e.A.AuthPathsPrefixLengths[j]), e.A.AuthPathsSuffixes[j])
            if len(prev) != height:
                raise ValueError("multipath: decoded path has the wrong height")
            paths.append(prev[::-1])
        out.append(DecodedPaths(paths, [[v % R_MOD for v in leaf] for leaf in e.B], list(e.A.LeafSiblingHashes),
                                np.asarray(e.A.LeafIndexes, dtype=np.uint64)))
    return out


def merkle_batch(decoded, byteorder="little"):
    """Flat arrays for Context.keccak_merkle_paths / b200g16_keccak_merkle_paths: (leaves u8[n, 32 k], siblings
    u8[n, 32], auth_paths u8[n, height, 32], indexes u64[n]).  Leaf bytes = the leaf's field elements, 32 bytes each
    (`byteorder`: ark-serialize writes little-endian; BASELINE configs[3] hashes them with the Keccak duplex)."""
    n = len(decoded.Leaves)
    k = len(decoded.Leaves[0]) if n else 0
    leaves = np.zeros((n, 32 * k), dtype=np.uint8)
    for i, leaf in enumerate(decoded.Leaves):
        if len(leaf) != k:
            raise ValueError("merkle_batch: ragged leaves")
        leaves[i] = np.frombuffer(b"".join(int(v).to_bytes(32, byteorder) for v in leaf), dtype=np.uint8)
    sib = np.frombuffer(b"".join(decoded.LeafSiblingHashes), dtype=np.uint8).reshape(n, 32).copy() if n else np.zeros((0, 32), np.uint8)
    height = len(decoded.AuthPaths[0]) if n else 0
    auth = (np.frombuffer(b"".join(b"".join(p) for p in decoded.AuthPaths), dtype=np.uint8).reshape(n, height, 32).copy()
            if n else np.zeros((0, 0, 32), np.uint8))
    return leaves, sib, auth, decoded.LeafIndexes.copy()
