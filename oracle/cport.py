"""ctypes binding of oracle/c/oracle.c (the C restatement; TEST INFRASTRUCTURE, see its header).

Build with `make -C oracle` (done by __graft_entry__.build()).  Array conventions are the
gnark-crypto memory layout used everywhere else: (n,4)/(n,8) uint64 Montgomery limbs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None
_vp = C.c_void_p


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = C.CDLL(LIB_PATH)
        lib.oracle_msm_g1.restype = C.c_int
        lib.oracle_msm_g1.argtypes = [_vp, _vp, C.c_size_t, C.c_int, _vp]
        lib.oracle_msm_window.restype = C.c_int
        lib.oracle_msm_window.argtypes = [C.c_size_t]
        lib.oracle_g1_progression.restype = C.c_int
        lib.oracle_g1_progression.argtypes = [_vp, _vp, C.c_size_t, _vp]
        lib.oracle_g1_gen_mul.restype = None
        lib.oracle_g1_gen_mul.argtypes = [_vp, _vp]
        lib.oracle_fr_dot.restype = None
        lib.oracle_fr_dot.argtypes = [_vp, _vp, C.c_size_t, C.c_int, _vp]
        lib.oracle_fr_dot_progression.restype = None
        lib.oracle_fr_dot_progression.argtypes = [_vp, _vp, _vp, C.c_size_t, _vp]
        lib.oracle_ntt.restype = C.c_int
        lib.oracle_ntt.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        lib.oracle_compute_h.restype = C.c_int
        lib.oracle_compute_h.argtypes = [_vp, _vp, _vp, C.c_int, C.c_int]
        lib.oracle_keccak_f_batch.restype = None
        lib.oracle_keccak_f_batch.argtypes = [_vp, C.c_size_t, C.c_int]
        lib.oracle_sponge.restype = None
        lib.oracle_sponge.argtypes = [_vp, C.c_size_t, _vp, C.c_size_t]
        lib.oracle_merkle_paths.restype = None
        lib.oracle_merkle_paths.argtypes = [_vp, C.c_size_t, _vp, _vp, _vp, C.c_uint, C.c_size_t, _vp, C.c_int]
        lib.oracle_threads.restype = C.c_int
        lib.oracle_msm_g2.restype = C.c_int
        lib.oracle_msm_g2.argtypes = [_vp, _vp, C.c_size_t, C.c_int, _vp]
        lib.oracle_g1_mul.restype = None
        lib.oracle_g1_mul.argtypes = [_vp, _vp, _vp]
        lib.oracle_g2_mul.restype = None
        lib.oracle_g2_mul.argtypes = [_vp, _vp, _vp]
        lib.oracle_g2_gen_mul.restype = None
        lib.oracle_g2_gen_mul.argtypes = [_vp, _vp]
        lib.oracle_groth16_prove.restype = C.c_int
        lib.oracle_groth16_prove.argtypes = ([C.c_int, C.c_size_t, _vp, C.c_size_t, _vp, C.c_size_t, _vp, C.c_size_t]
                                             + [_vp] * 14 + [C.c_size_t, _vp, _vp, C.c_int, _vp, _vp])
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(_vp)


def _u64(a, cols):
    return np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, cols)


def threads():
    return int(load().oracle_threads())


def msm_g1(points, scalars, nthreads=0):
    pts, sc = _u64(points, 8), _u64(scalars, 4)
    out = np.zeros(8, dtype=np.uint64)
    if load().oracle_msm_g1(_p(pts), _p(sc), sc.shape[0], nthreads, _p(out)):
        raise MemoryError("oracle_msm_g1")
    return out


def g1_progression(k0_mont, d_mont, n):
    k0, d = _u64(k0_mont, 4), _u64(d_mont, 4)
    out = np.zeros((n, 8), dtype=np.uint64)
    if load().oracle_g1_progression(_p(k0), _p(d), n, _p(out)):
        raise MemoryError("oracle_g1_progression")
    return out


def g1_gen_mul(k_mont):
    k = _u64(k_mont, 4)
    out = np.zeros(8, dtype=np.uint64)
    load().oracle_g1_gen_mul(_p(k), _p(out))
    return out


def fr_dot(a, b, nthreads=0):
    a, b = _u64(a, 4), _u64(b, 4)
    out = np.zeros(4, dtype=np.uint64)
    load().oracle_fr_dot(_p(a), _p(b), a.shape[0], nthreads, _p(out))
    return out


def fr_dot_progression(a, k0_mont, d_mont):
    a, k0, d = _u64(a, 4), _u64(k0_mont, 4), _u64(d_mont, 4)
    out = np.zeros(4, dtype=np.uint64)
    load().oracle_fr_dot_progression(_p(a), _p(k0), _p(d), a.shape[0], _p(out))
    return out


def ntt(data, inverse=False, coset=False, decimation=0, nthreads=0):
    a = _u64(data, 4).copy()
    n = a.shape[0]
    if load().oracle_ntt(_p(a), n.bit_length() - 1, int(inverse), int(coset), int(decimation), nthreads):
        raise MemoryError("oracle_ntt")
    return a


def compute_h(a, b, c, log2n, nthreads=0):
    n = 1 << log2n
    bufs = []
    for v in (a, b, c):
        v = _u64(v, 4)
        p = np.zeros((n, 4), dtype=np.uint64)
        p[:v.shape[0]] = v
        bufs.append(p)
    if load().oracle_compute_h(_p(bufs[0]), _p(bufs[1]), _p(bufs[2]), log2n, nthreads):
        raise MemoryError("oracle_compute_h")
    return bufs[0]


def keccak_f_batch(states, nthreads=0):
    st = np.ascontiguousarray(states, dtype=np.uint64).reshape(-1, 25).copy()
    load().oracle_keccak_f_batch(_p(st), st.shape[0], nthreads)
    return st


def sponge(data, out_len):
    d = np.frombuffer(bytes(data), dtype=np.uint8).copy() if len(data) else np.zeros(1, dtype=np.uint8)
    out = np.zeros(max(out_len, 1), dtype=np.uint8)
    load().oracle_sponge(_p(d), len(data), _p(out), out_len)
    return bytes(out[:out_len])


def merkle_paths(leaves, siblings, auth_paths, indexes, nthreads=0):
    leaves = np.ascontiguousarray(leaves, dtype=np.uint8)
    siblings = np.ascontiguousarray(siblings, dtype=np.uint8)
    auth_paths = np.ascontiguousarray(auth_paths, dtype=np.uint8)
    indexes = np.ascontiguousarray(indexes, dtype=np.uint64)
    n, leaf_len = leaves.shape
    height = auth_paths.shape[1] + 1
    roots = np.zeros((n, 32), dtype=np.uint8)
    load().oracle_merkle_paths(_p(leaves), leaf_len, _p(siblings), _p(auth_paths), _p(indexes), height, n, _p(roots),
                               nthreads)
    return roots


def msm_g2(points, scalars, nthreads=0):
    pts, sc = _u64(points, 16), _u64(scalars, 4)
    out = np.zeros(16, dtype=np.uint64)
    if load().oracle_msm_g2(_p(pts), _p(sc), sc.shape[0], nthreads, _p(out)):
        raise MemoryError("oracle_msm_g2")
    return out


def g1_mul(point, k_mont):
    p, k = _u64(point, 8), _u64(k_mont, 4)
    out = np.zeros(8, dtype=np.uint64)
    load().oracle_g1_mul(_p(p), _p(k), _p(out))
    return out


def g2_mul(point, k_mont):
    p, k = _u64(point, 16), _u64(k_mont, 4)
    out = np.zeros(16, dtype=np.uint64)
    load().oracle_g2_mul(_p(p), _p(k), _p(out))
    return out


def g2_gen_mul(k_mont):
    k = _u64(k_mont, 4)
    out = np.zeros(16, dtype=np.uint64)
    load().oracle_g2_gen_mul(_p(k), _p(out))
    return out


PROVE_OUT = (("ar", 8), ("bs", 16), ("krs", 8), ("msm_a", 8), ("msm_b1", 8), ("msm_k", 8), ("msm_z", 8), ("msm_b2", 16))


def groth16_prove(log2n, A, B1, K, Z, B2, alpha, beta, delta, beta2, delta2, infinity_a, infinity_b, k_skip,
                  wires, a, b, c, r, s, nthreads=0, want_h=False):
    """The C restatement of gnark's Prove after Solve (oracle_groth16_prove).  Arrays as in
    gnark_whir_b200.lib.Context.pk_upload / prove.  Returns (dict of affine points, h or None)."""
    A, B1, K, Z, B2 = _u64(A, 8), _u64(B1, 8), _u64(K, 8), _u64(Z, 8), _u64(B2, 16)
    small = [_u64(alpha, 8), _u64(beta, 8), _u64(delta, 8), _u64(beta2, 16), _u64(delta2, 16)]
    flags = [np.ascontiguousarray(f, dtype=np.uint8) for f in (infinity_a, infinity_b, k_skip)]
    wires, a, b, c = _u64(wires, 4), _u64(a, 4), _u64(b, 4), _u64(c, 4)
    r, s = _u64(r, 4), _u64(s, 4)
    n = 1 << log2n
    if Z.shape[0] != n - 1 or B2.shape[0] != B1.shape[0]:
        raise ValueError("groth16_prove: Z must have N-1 points, B2 as many as B1")
    out = np.zeros(80, dtype=np.uint64)
    h = np.zeros((n, 4), dtype=np.uint64) if want_h else None
    st = load().oracle_groth16_prove(log2n, wires.shape[0], _p(A), A.shape[0], _p(B1), B1.shape[0], _p(K), K.shape[0],
                                     _p(Z), _p(B2), *[_p(x) for x in small], *[_p(f) for f in flags], _p(wires),
                                     _p(a), _p(b), _p(c), a.shape[0], _p(r), _p(s), nthreads, _p(out),
                                     _p(h) if want_h else None)
    if st:
        raise RuntimeError(f"oracle_groth16_prove: status {st}")
    res, o = {}, 0
    for k, w in PROVE_OUT:
        res[k] = out[o:o + w].copy()
        o += w
    return res, h
