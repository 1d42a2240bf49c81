"""ctypes binding of libb200g16.so — the same C-ABI a cgo shim binds (include/b200g16.h).

This module is plumbing for tests / bench / the Python host mirror; it contains no
arithmetic and NO fallback: if the shared library (built by __graft_entry__.build()) is
missing, or no sm_100 device is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200g16.so")

DIF, DIT = 0, 1


class B200Error(RuntimeError):
    pass


_lib = None

_u64p = C.POINTER(C.c_uint64)
_vp = C.c_void_p
_sz = C.c_size_t

# name -> (restype, argtypes); kept in one table so tests can check it against the header
SIGNATURES = {
    "b200g16_version": (C.c_int, []),
    "b200g16_last_error": (C.c_char_p, []),
    "b200g16_init": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "b200g16_destroy": (None, [_vp]),
    "b200g16_launch_count": (C.c_uint64, [_vp]),
    "b200g16_last_timings": (C.c_int, [_vp, C.POINTER(C.c_float), C.c_int]),
    "b200g16_set_msm_window": (C.c_int, [_vp, C.c_int]),
    "b200g16_set_msm_batch_affine": (C.c_int, [_vp, C.c_int, C.c_int, C.c_uint]),
    "b200g16_bases_upload_g1": (C.c_int, [_vp, _vp, _sz, C.POINTER(_vp)]),
    "b200g16_bases_upload_g2": (C.c_int, [_vp, _vp, _sz, C.POINTER(_vp)]),
    "b200g16_bases_free": (None, [_vp]),
    "b200g16_bases_len": (_sz, [_vp]),
    "b200g16_bases_precompute": (C.c_int, [_vp, _vp, C.c_int]),
    "b200g16_bases_window": (C.c_int, [_vp]),
    "b200g16_msm_plan": (C.c_int, [_vp, _vp, _sz, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "b200g16_bases_download": (C.c_int, [_vp, _sz, _sz, _vp]),
    "b200g16_fixed_base_mul_g1": (C.c_int, [_vp, _vp, _vp, _sz, _vp]),
    "b200g16_fixed_base_mul_g2": (C.c_int, [_vp, _vp, _vp, _sz, _vp]),
    "b200g16_bases_from_scalars_g1": (C.c_int, [_vp, _vp, _vp, _sz, C.POINTER(_vp)]),
    "b200g16_bases_from_scalars_g2": (C.c_int, [_vp, _vp, _vp, _sz, C.POINTER(_vp)]),
    "b200g16_modmul_probe": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_float)]),
    "b200g16_pipe_probe": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_float)]),
    "b200g16_fp52_probe": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_float)]),
    "b200g16_msm_g1": (C.c_int, [_vp, _vp, _sz, _vp, _sz, _vp]),
    "b200g16_msm_g2": (C.c_int, [_vp, _vp, _sz, _vp, _sz, _vp]),
    "b200g16_msm_g1_dev": (C.c_int, [_vp, _vp, _sz, _vp, _sz, _vp]),
    "b200g16_msm_g1_begin": (C.c_int, [_vp, _vp, _sz, _vp, _sz, C.POINTER(C.c_int)]),
    "b200g16_msm_g1_begin_dev": (C.c_int, [_vp, _vp, _sz, _vp, _sz, C.POINTER(C.c_int)]),
    "b200g16_msm_g1_end": (C.c_int, [_vp, C.c_int, _vp]),
    "b200g16_msm_g2_dev": (C.c_int, [_vp, _vp, _sz, _vp, _sz, _vp]),
    "b200g16_ntt": (C.c_int, [_vp, _vp, C.c_uint, C.c_int, C.c_int, C.c_int]),
    "b200g16_ntt_dev": (C.c_int, [_vp, _vp, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_int]),
    "b200g16_compute_h": (C.c_int, [_vp, _vp, _vp, _vp, _sz, C.c_uint, _vp]),
    "b200g16_compute_h_dev": (C.c_int, [_vp, _vp, _vp, _vp, C.c_uint]),
    "b200g16_keccak_f_batch": (C.c_int, [_vp, _vp, _sz]),
    "b200g16_keccak_f_batch_dev": (C.c_int, [_vp, _vp, _sz]),
    "b200g16_keccak_sponge_batch": (C.c_int, [_vp, _vp, _sz, _sz, _vp, _sz]),
    "b200g16_keccak_merkle_paths": (C.c_int, [_vp, _vp, _sz, _vp, _vp, _vp, C.c_uint, _sz, _vp, _vp, _vp]),
    "b200g16_keccak_merkle_paths_dev": (C.c_int, [_vp, _vp, _sz, _vp, _vp, _vp, C.c_uint, _sz, _vp, _vp, _vp]),
    "b200g16_pk_upload": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "b200g16_pk_free": (None, [_vp]),
    "b200g16_prove": (C.c_int, [_vp, _vp, _vp, _sz, _vp, _vp, _vp, _sz, _vp, _vp, _vp, _vp]),
    "b200g16_prove_dev": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200g16_prove_h_dev": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200g16_h_pointwise_dev": (C.c_int, [_vp, _vp, _vp, _vp, C.c_uint]),
    "b200g16_prove_begin_dev": (C.c_int, [_vp, _vp, _vp]),
    "b200g16_prove_end_dev": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "b200g16_prove_finish": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200g16_pairing_check": (C.c_int, [_vp, _vp, _vp, _sz, C.POINTER(C.c_int)]),
    "b200g16_pair": (C.c_int, [_vp, _vp, _vp, _sz, _vp]),
    "b200g16_verify": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, C.POINTER(C.c_int)]),
    "b200g16_g1_decode": (C.c_int, [_vp, _vp, _sz, C.c_int, _vp, _vp]),
    "b200g16_g2_decode": (C.c_int, [_vp, _vp, _sz, C.c_int, C.c_int, _vp, _vp]),
    "b200g16_g1_encode": (C.c_int, [_vp, _vp, _sz, C.c_int, _vp]),
    "b200g16_g2_encode": (C.c_int, [_vp, _vp, _sz, C.c_int, _vp]),
    "b200g16_group_init": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(_vp)]),
    "b200g16_group_destroy": (None, [_vp]),
    "b200g16_group_size": (C.c_int, [_vp]),
    "b200g16_group_ctx": (_vp, [_vp, C.c_int]),
    "b200g16_host_register": (C.c_int, [_vp, _sz]),
    "b200g16_host_unregister": (C.c_int, [_vp]),
    "b200g16_group_bases_upload_g1": (C.c_int, [_vp, _vp, _sz, C.POINTER(_vp)]),
    "b200g16_group_bases_upload_g2": (C.c_int, [_vp, _vp, _sz, C.POINTER(_vp)]),
    "b200g16_group_bases_precompute": (C.c_int, [_vp, _vp, C.c_int]),
    "b200g16_group_bases_free": (None, [_vp]),
    "b200g16_group_msm_g1": (C.c_int, [_vp, _vp, _vp, _sz, _vp]),
    "b200g16_group_msm_g2": (C.c_int, [_vp, _vp, _vp, _sz, _vp]),
    "b200g16_group_pk_upload": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "b200g16_group_pk_free": (None, [_vp]),
    "b200g16_group_prove": (C.c_int, [_vp, _vp, _vp, _sz, _vp, _vp, _vp, _sz, _vp, _vp, _vp, _vp]),
    "b200g16_dist_h_init": (C.c_int, [_vp, C.c_uint, C.c_int, C.c_int, _vp]),
    "b200g16_dist_h_open": (C.c_int, [_vp, _vp]),
    "b200g16_dist_h_load": (C.c_int, [_vp, _vp, _vp, _vp]),
    "b200g16_dist_h_slice": (_vp, [_vp, C.c_int]),
    "b200g16_dist_h_phase": (C.c_int, [_vp, C.c_int]),
    "b200g16_dist_h_close": (None, [_vp]),
    "b200g16_g1_add": (C.c_int, [_vp, _vp, _vp]),
    "b200g16_g2_add": (C.c_int, [_vp, _vp, _vp]),
    "b200g16_g1_scalar_mul": (C.c_int, [_vp, _vp, _vp]),
    "b200g16_g2_scalar_mul": (C.c_int, [_vp, _vp, _vp]),
}


def load():
    """dlopen the library (raises if it has not been built — there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200Error(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


# b200g16_set_msm_batch_affine(mode, levels, min_pairs) as a fresh ctx has it
MSM_BATCH_AFFINE_DEFAULT = (0, 3, 192)


def _check(status):
    if status != 0:
        raise B200Error(f"libb200g16 error {status}: {load().b200g16_last_error().decode()}")


def _dp(dev_ptr):
    """A raw device pointer for the library.  The library works on its OWN (non-blocking) streams and cannot know which
    stream produced the data, so device-side work the caller still has in flight on torch's streams (a clone, a random
    fill) is drained first — the C-ABI's contract for device pointers (include/b200g16.h) is "the memory is ready"."""
    import sys
    torch = sys.modules.get("torch")
    if torch is not None and torch.cuda.is_initialized():
        torch.cuda.current_stream().synchronize()   # the producer's stream only: an open b200g16_msm_g1_begin keeps running
    return _vp(int(dev_ptr))


def _ptr(a):
    return a.ctypes.data_as(_vp)


def _u64(a, cols=None):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if cols is not None:
        a = a.reshape(-1, cols)
    return a


# ---- host-only helpers (no GPU needed) -----------------------------------------------------
def g1_add(a, b):
    out = np.zeros(8, dtype=np.uint64)
    _check(load().b200g16_g1_add(_ptr(_u64(a)), _ptr(_u64(b)), _ptr(out)))
    return out


def g2_add(a, b):
    out = np.zeros(16, dtype=np.uint64)
    _check(load().b200g16_g2_add(_ptr(_u64(a)), _ptr(_u64(b)), _ptr(out)))
    return out


def g1_scalar_mul(p, k):
    out = np.zeros(8, dtype=np.uint64)
    _check(load().b200g16_g1_scalar_mul(_ptr(_u64(p)), _ptr(_u64(k)), _ptr(out)))
    return out


def g2_scalar_mul(p, k):
    out = np.zeros(16, dtype=np.uint64)
    _check(load().b200g16_g2_scalar_mul(_ptr(_u64(p)), _ptr(_u64(k)), _ptr(out)))
    return out


class PkDesc(C.Structure):
    """struct b200g16_pk_desc (include/b200g16.h)"""
    _fields_ = [
        ("log2_domain", C.c_uint), ("n_wires", _sz),
        ("g1_a", _vp), ("g1_b", _vp), ("g1_k", _vp), ("g1_z", _vp), ("g2_b", _vp),
        ("res_a", _vp), ("res_b", _vp), ("res_k", _vp), ("res_z", _vp), ("res_b2", _vp),
        ("n_a", _sz), ("n_b", _sz), ("n_k", _sz), ("n_z", _sz),
        ("g1_alpha", _vp), ("g1_beta", _vp), ("g1_delta", _vp), ("g2_beta", _vp), ("g2_delta", _vp),
        ("infinity_a", _vp), ("infinity_b", _vp), ("k_skip", _vp),
        ("partial", C.c_int), ("off_a", _sz), ("off_b", _sz), ("off_k", _sz), ("off_z", _sz),
        ("precompute", C.c_int),
    ]


class VkDesc(C.Structure):
    """struct b200g16_vk_desc (include/b200g16.h)"""
    _fields_ = [("g1_alpha", _vp), ("g2_beta", _vp), ("g2_gamma", _vp), ("g2_delta", _vp), ("g1_k", _vp), ("n_k", _sz),
                ("ped_g", _vp), ("ped_g_sigma_neg", _vp)]


class ProofOut(C.Structure):
    """struct b200g16_proof (include/b200g16.h)"""
    _fields_ = [
        ("ar", C.c_uint64 * 8), ("bs", C.c_uint64 * 16), ("krs", C.c_uint64 * 8),
        ("msm_a", C.c_uint64 * 8), ("msm_b1", C.c_uint64 * 8), ("msm_k", C.c_uint64 * 8),
        ("msm_z", C.c_uint64 * 8), ("msm_b2", C.c_uint64 * 16), ("bs1", C.c_uint64 * 8),
    ]

    def as_dict(self):
        return {name: np.array(getattr(self, name)[:], dtype=np.uint64) for name, _ in self._fields_}


def _pk_desc(log2_domain, n_wires, A, B, K, Z, B2, alpha, beta, delta, beta2, delta2, infinity_a, infinity_b, k_skip,
             partial=False, offsets=(0, 0, 0, 0), precompute=False):
    """struct b200g16_pk_desc from numpy arrays / Bases handles -> (desc, buffers to keep alive during the call)"""
    keep = []

    def vec(v, cols):
        if isinstance(v, Bases):
            return None, v.handle, v.n
        a = _u64(v, cols)
        keep.append(a)
        return _ptr(a), None, a.shape[0]

    d = PkDesc()
    d.log2_domain, d.n_wires = log2_domain, n_wires
    d.partial = int(bool(partial))
    d.precompute = int(bool(precompute))
    d.off_a, d.off_b, d.off_k, d.off_z = [int(x) for x in offsets]
    d.g1_a, d.res_a, d.n_a = vec(A, 8)
    d.g1_b, d.res_b, d.n_b = vec(B, 8)
    d.g1_k, d.res_k, d.n_k = vec(K, 8)
    d.g1_z, d.res_z, d.n_z = vec(Z, 8)
    d.g2_b, d.res_b2, nb2 = vec(B2, 16)
    if nb2 != d.n_b:
        raise B200Error("pk_upload: len(G2.B) != len(G1.B)")
    for name, val, w in (("g1_alpha", alpha, 8), ("g1_beta", beta, 8), ("g1_delta", delta, 8),
                         ("g2_beta", beta2, 16), ("g2_delta", delta2, 16)):
        a = _u64(val).reshape(w)
        keep.append(a)
        setattr(d, name, _ptr(a))
    for name, val in (("infinity_a", infinity_a), ("infinity_b", infinity_b), ("k_skip", k_skip)):
        a = np.ascontiguousarray(val, dtype=np.uint8)
        if a.shape[0] != n_wires:
            raise B200Error(f"pk_upload: {name} must have n_wires entries")
        keep.append(a)
        setattr(d, name, _ptr(a))
    return d, keep


def host_register(arr):
    """Page-lock a numpy array in place (b200g16_host_register); returns the array."""
    _check(load().b200g16_host_register(_ptr(arr), arr.nbytes))
    return arr


def host_unregister(arr):
    _check(load().b200g16_host_unregister(_ptr(arr)))


class Group:
    """Several GPUs driven from this one process (b200g16_group_*): point-range shards of the proving key, peer
    copies for h, partial points added on the host.  devices may repeat a device (several shards on one GPU)."""

    def __init__(self, devices):
        arr = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = _vp()
        _check(load().b200g16_group_init(arr, len(devices), C.byref(h)))
        self.h, self.devices = h, list(devices)

    def close(self):
        if self.h:
            load().b200g16_group_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __len__(self):
        return int(load().b200g16_group_size(self.h))

    def ctx(self, i):
        """Borrowed Context of the i-th device (do not close it)."""
        c = Context.__new__(Context)
        c.h, c.device = _vp(load().b200g16_group_ctx(self.h, i)), self.devices[i]
        return c

    def upload(self, points, group=1, precompute=False):
        pts = _u64(points, 8 if group == 1 else 16)
        h = _vp()
        fn = load().b200g16_group_bases_upload_g1 if group == 1 else load().b200g16_group_bases_upload_g2
        _check(fn(self.h, _ptr(pts), pts.shape[0], C.byref(h)))
        if precompute:
            _check(load().b200g16_group_bases_precompute(self.h, h, 0))
        return (h, group, pts.shape[0])

    def bases_free(self, gb):
        load().b200g16_group_bases_free(gb[0])

    def msm(self, gb, scalars):
        h, group, n = gb
        sc = _u64(scalars, 4)
        out = np.zeros(8 if group == 1 else 16, dtype=np.uint64)
        fn = load().b200g16_group_msm_g1 if group == 1 else load().b200g16_group_msm_g2
        _check(fn(self.h, h, _ptr(sc), sc.shape[0], _ptr(out)))
        return out

    def pk_upload(self, log2_domain, n_wires, A, B, K, Z, B2, alpha, beta, delta, beta2, delta2, infinity_a, infinity_b,
                  k_skip, precompute=False):
        """The whole key as host arrays; the library cuts and uploads the shards."""
        d, keep = _pk_desc(log2_domain, n_wires, A, B, K, Z, B2, alpha, beta, delta, beta2, delta2, infinity_a, infinity_b,
                           k_skip, False, (0, 0, 0, 0), precompute)
        h = _vp()
        _check(load().b200g16_group_pk_upload(self.h, C.byref(d), C.byref(h)))
        del keep
        return h

    def pk_free(self, pk):
        load().b200g16_group_pk_free(pk)

    def prove(self, pk, wires, a, b, c, r, s, want_h=False, log2_domain=None):
        wires, a, b, c = _u64(wires, 4), _u64(a, 4), _u64(b, 4), _u64(c, 4)
        r, s = _u64(r).reshape(4), _u64(s).reshape(4)
        out = ProofOut()
        h = np.zeros((1 << log2_domain, 4), dtype=np.uint64) if want_h else None
        _check(load().b200g16_group_prove(self.h, pk, _ptr(wires), wires.shape[0], _ptr(a), _ptr(b), _ptr(c), a.shape[0],
                                          _ptr(r), _ptr(s), C.byref(out), _ptr(h) if want_h else None))
        return out.as_dict(), h


class Bases:
    """Device-resident vector of affine points (one of the proving key's point arrays)."""

    def __init__(self, ctx, handle, group, n):
        self.ctx, self.handle, self.group, self.n = ctx, handle, group, n

    def free(self):
        if self.handle:
            load().b200g16_bases_free(self.handle)
            self.handle = None

    def __len__(self):
        return self.n

    def precompute(self, window_bits=0):
        """Attach a window table (b200g16_bases_precompute); returns the window width used."""
        _check(load().b200g16_bases_precompute(self.ctx.h, self.handle, int(window_bits)))
        return self.window()

    def window(self):
        return int(load().b200g16_bases_window(self.handle))

    def download(self, offset=0, n=None):
        n = self.n - offset if n is None else n
        out = np.zeros((n, 8 if self.group == 1 else 16), dtype=np.uint64)
        _check(load().b200g16_bases_download(self.handle, offset, n, _ptr(out)))
        return out


class Context:
    """One per process per GPU (b200g16_init)."""

    def __init__(self, device=0):
        h = _vp()
        _check(load().b200g16_init(int(device), C.byref(h)))
        self.h = h
        self.device = device

    def close(self):
        if self.h:
            load().b200g16_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- bookkeeping
    def launch_count(self):
        return int(load().b200g16_launch_count(self.h))

    def last_timings(self):
        buf = (C.c_float * 16)()
        n = load().b200g16_last_timings(self.h, buf, 16)
        return [buf[i] for i in range(n)]

    def set_msm_window(self, c):
        _check(load().b200g16_set_msm_window(self.h, int(c)))

    def set_msm_batch_affine(self, mode, levels=0, min_pairs=0):
        """0 = mixed XYZZ bucket additions, 1 = batched-affine pair tree for large inputs, 2 = always."""
        _check(load().b200g16_set_msm_batch_affine(self.h, int(mode), int(levels), int(min_pairs)))

    # -- bases
    def upload_g1(self, points):
        pts = _u64(points, 8)
        h = _vp()
        _check(load().b200g16_bases_upload_g1(self.h, _ptr(pts), pts.shape[0], C.byref(h)))
        return Bases(self, h, 1, pts.shape[0])

    def upload_g2(self, points):
        pts = _u64(points, 16)
        h = _vp()
        _check(load().b200g16_bases_upload_g2(self.h, _ptr(pts), pts.shape[0], C.byref(h)))
        return Bases(self, h, 2, pts.shape[0])

    # -- fixed-base batch scalar multiplication (BatchScalarMultiplicationG1/G2)
    def fixed_base_mul(self, base, scalars, group=1, resident=False):
        lib = load()
        words = 8 if group == 1 else 16
        base = _u64(base).reshape(words)
        sc = _u64(scalars, 4)
        n = sc.shape[0]
        if resident:
            h = _vp()
            fn = lib.b200g16_bases_from_scalars_g1 if group == 1 else lib.b200g16_bases_from_scalars_g2
            _check(fn(self.h, _ptr(base), _ptr(sc), n, C.byref(h)))
            return Bases(self, h, group, n)
        out = np.zeros((n, words), dtype=np.uint64)
        fn = lib.b200g16_fixed_base_mul_g1 if group == 1 else lib.b200g16_fixed_base_mul_g2
        _check(fn(self.h, _ptr(base), _ptr(sc), n, _ptr(out)))
        return out

    def modmul_probe(self, blocks_per_sm=4, chains=4, iters=2000):
        rate = C.c_double()
        ms = C.c_float()
        _check(load().b200g16_modmul_probe(self.h, blocks_per_sm, chains, iters, C.byref(rate), C.byref(ms)))
        return rate.value, ms.value

    def pipe_probe(self, mode, blocks_per_sm=8, iters=2000):
        rate = C.c_double()
        ms = C.c_float()
        _check(load().b200g16_pipe_probe(self.h, mode, blocks_per_sm, iters, C.byref(rate), C.byref(ms)))
        return rate.value, ms.value

    def fp52_probe(self, variant, blocks_per_sm=4, iters=2000):
        rate = C.c_double()
        ms = C.c_float()
        _check(load().b200g16_fp52_probe(self.h, variant, blocks_per_sm, iters, C.byref(rate), C.byref(ms)))
        return rate.value, ms.value

    # -- NTT / computeH (mirror fft.Domain.FFT / FFTInverse and prove.go computeH)
    def ntt(self, data, inverse=False, coset=False, decimation=DIF):
        """data: (N,4) uint64 Montgomery, N a power of two; returns the transformed copy."""
        a = _u64(data, 4).copy()
        n = a.shape[0]
        if n == 0 or n & (n - 1):
            raise B200Error("ntt: length must be a power of two")
        _check(load().b200g16_ntt(self.h, _ptr(a), n.bit_length() - 1, int(inverse), int(coset), int(decimation)))
        return a

    def ntt_dev(self, d_ptr, log2n, batch=1, inverse=False, coset=False, decimation=DIF):
        _check(load().b200g16_ntt_dev(self.h, _dp(d_ptr), log2n, batch, int(inverse), int(coset), int(decimation)))

    def compute_h(self, a, b, c, log2n):
        a, b, c = _u64(a, 4), _u64(b, 4), _u64(c, 4)
        h = np.zeros((1 << log2n, 4), dtype=np.uint64)
        _check(load().b200g16_compute_h(self.h, _ptr(a), _ptr(b), _ptr(c), a.shape[0], log2n, _ptr(h)))
        return h

    def compute_h_dev(self, d_a, d_b, d_c, log2n):
        _check(load().b200g16_compute_h_dev(self.h, _dp(d_a), _dp(d_b), _dp(d_c), log2n))

    # -- Keccak (mirror keccakf.Permute, keccakSponge.Digest, VerifyMerkleTreeProofs)
    def keccak_f_batch(self, states):
        st = np.ascontiguousarray(states, dtype=np.uint64).reshape(-1, 25).copy()
        _check(load().b200g16_keccak_f_batch(self.h, _ptr(st), st.shape[0]))
        return st

    def keccak_f_batch_dev(self, d_ptr, n):
        _check(load().b200g16_keccak_f_batch_dev(self.h, _dp(d_ptr), n))

    def keccak_sponge_batch(self, inputs, out_len):
        """inputs: (n, in_len) uint8 -> (n, out_len) uint8"""
        a = np.ascontiguousarray(inputs, dtype=np.uint8)
        n, in_len = a.shape
        out = np.zeros((n, out_len), dtype=np.uint8)
        _check(load().b200g16_keccak_sponge_batch(self.h, _ptr(a), in_len, n, _ptr(out), out_len))
        return out

    def keccak_merkle_paths(self, leaves, siblings, auth_paths, indexes, expected_root=None):
        """leaves (n, leaf_len) u8; siblings (n, 32) u8; auth_paths (n, height-1, 32) u8;
        indexes (n,) u64 -> (roots (n,32) u8, ok (n,) bool or None)"""
        leaves = np.ascontiguousarray(leaves, dtype=np.uint8)
        siblings = np.ascontiguousarray(siblings, dtype=np.uint8)
        auth_paths = np.ascontiguousarray(auth_paths, dtype=np.uint8)
        indexes = np.ascontiguousarray(indexes, dtype=np.uint64)
        n, leaf_len = leaves.shape
        height = auth_paths.shape[1] + 1
        roots = np.zeros((n, 32), dtype=np.uint8)
        ok = np.zeros(n, dtype=np.uint8) if expected_root is not None else None
        er = np.frombuffer(bytes(expected_root), dtype=np.uint8).copy() if expected_root is not None else None
        _check(load().b200g16_keccak_merkle_paths(
            self.h, _ptr(leaves), leaf_len, _ptr(siblings), _ptr(auth_paths) if auth_paths.size else None,
            _ptr(indexes), height, n, _ptr(er) if er is not None else None, _ptr(roots),
            _ptr(ok) if ok is not None else None))
        return roots, (ok.astype(bool) if ok is not None else None)

    # -- Groth16 prove (pk resident; mirrors groth16_bn254.Prove after Solve)
    def pk_upload(self, log2_domain, n_wires, A, B, K, Z, B2, alpha, beta, delta, beta2, delta2,
                  infinity_a, infinity_b, k_skip, partial=False, offsets=(0, 0, 0, 0), precompute=False):
        """A/B/K/Z/B2: numpy point arrays (host) or Bases (already resident, borrowed).
        partial=True: the vectors are entries [off, off+len) of the full key vectors
        (offsets = (off_a, off_b, off_k, off_z)); prove() then returns partial MSM sums."""
        d, keep = _pk_desc(log2_domain, n_wires, A, B, K, Z, B2, alpha, beta, delta, beta2, delta2, infinity_a, infinity_b,
                           k_skip, partial, offsets, precompute)
        h = _vp()
        _check(load().b200g16_pk_upload(self.h, C.byref(d), C.byref(h)))
        del keep
        return h

    def pk_free(self, pk):
        load().b200g16_pk_free(pk)

    def prove(self, pk, wires, a, b, c, r, s, want_h=False, log2_domain=None):
        wires, a, b, c = _u64(wires, 4), _u64(a, 4), _u64(b, 4), _u64(c, 4)
        r, s = _u64(r).reshape(4), _u64(s).reshape(4)
        out = ProofOut()
        h = np.zeros((1 << log2_domain, 4), dtype=np.uint64) if want_h else None
        _check(load().b200g16_prove(self.h, pk, _ptr(wires), wires.shape[0], _ptr(a), _ptr(b), _ptr(c), a.shape[0],
                                    _ptr(r), _ptr(s), C.byref(out), _ptr(h) if want_h else None))
        return out.as_dict(), h

    def prove_finish(self, pk, msm_a, msm_b1, msm_k, msm_z, msm_b2, r, s):
        """Final assembly from the five complete (summed over shards) MSM results."""
        r, s = _u64(r).reshape(4), _u64(s).reshape(4)
        pts = [_u64(v).reshape(w) for v, w in ((msm_a, 8), (msm_b1, 8), (msm_k, 8), (msm_z, 8), (msm_b2, 16))]
        out = ProofOut()
        _check(load().b200g16_prove_finish(pk, *[_ptr(p) for p in pts], _ptr(r), _ptr(s), C.byref(out)))
        return out.as_dict()

    def prove_dev(self, pk, d_wires, d_a, d_b, d_c, r, s):
        r, s = _u64(r).reshape(4), _u64(s).reshape(4)
        out = ProofOut()
        _check(load().b200g16_prove_dev(self.h, pk, _dp(d_wires), _dp(d_a), _dp(d_b), _dp(d_c),
                                        _ptr(r), _ptr(s), C.byref(out)))
        return out.as_dict()

    # -- pairing / verify (mirror bn254.PairingCheck, bn254.Pair, groth16.Verify)
    def pairing_check(self, g1_points, g2_points):
        p, q = _u64(g1_points, 8), _u64(g2_points, 16)
        if p.shape[0] != q.shape[0]:
            raise B200Error("pairing_check: len(P) != len(Q)")
        ok = C.c_int()
        _check(load().b200g16_pairing_check(self.h, _ptr(p), _ptr(q), p.shape[0], C.byref(ok)))
        return bool(ok.value)

    def pair(self, g1_points, g2_points):
        """-> GT element, (6, 8) uint64: C0.B0, C0.B1, C0.B2, C1.B0, C1.B1, C1.B2 (gnark E12 order)"""
        p, q = _u64(g1_points, 8), _u64(g2_points, 16)
        if p.shape[0] != q.shape[0]:
            raise B200Error("pair: len(P) != len(Q)")
        out = np.zeros((6, 8), dtype=np.uint64)
        _check(load().b200g16_pair(self.h, _ptr(p), _ptr(q), p.shape[0], _ptr(out)))
        return out

    def verify(self, alpha, beta2, gamma2, delta2, K, ar, bs, krs, public_inputs, commitment=None, pok=None,
               ped_g=None, ped_g_sigma_neg=None):
        keep = [_u64(alpha).reshape(8), _u64(beta2).reshape(16), _u64(gamma2).reshape(16), _u64(delta2).reshape(16),
                _u64(K, 8)]
        d = VkDesc()
        d.g1_alpha, d.g2_beta, d.g2_gamma, d.g2_delta, d.g1_k = [_ptr(a) for a in keep]
        d.n_k = keep[4].shape[0]
        opt = [None if v is None else _u64(v).reshape(w) for v, w in ((ped_g, 16), (ped_g_sigma_neg, 16),
                                                                        (commitment, 8), (pok, 8))]
        d.ped_g = _ptr(opt[0]) if opt[0] is not None else None
        d.ped_g_sigma_neg = _ptr(opt[1]) if opt[1] is not None else None
        pts = [_u64(ar).reshape(8), _u64(bs).reshape(16), _u64(krs).reshape(8)]
        pub = _u64(public_inputs, 4)
        ok = C.c_int()
        _check(load().b200g16_verify(self.h, C.byref(d), _ptr(pts[0]), _ptr(pts[1]), _ptr(pts[2]),
                                     _ptr(opt[2]) if opt[2] is not None else None,
                                     _ptr(opt[3]) if opt[3] is not None else None,
                                     _ptr(pub) if pub.shape[0] else None, pub.shape[0], C.byref(ok)))
        return bool(ok.value)

    # -- gnark-crypto point encodings (G1Affine.Bytes / RawBytes / SetBytes and the G2 counterparts)
    def encode_points(self, points, group=1, raw=False):
        """points (n, 8|16) uint64 -> (n, record) uint8, record = 32|64 compressed, 64|128 raw"""
        pts = _u64(points, 8 if group == 1 else 16)
        rec = (32 if group == 1 else 64) * (2 if raw else 1)
        out = np.zeros((pts.shape[0], rec), dtype=np.uint8)
        fn = load().b200g16_g1_encode if group == 1 else load().b200g16_g2_encode
        _check(fn(self.h, _ptr(pts), pts.shape[0], int(raw), _ptr(out)))
        return out

    def decode_points(self, data, group=1, raw=False, subgroup_check=True):
        """data: bytes / uint8 array of n fixed-size records -> (points (n, 8|16) uint64, ok (n,) bool)"""
        rec = (32 if group == 1 else 64) * (2 if raw else 1)
        buf = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8) if isinstance(data, (bytes, bytearray))
                                   else data, dtype=np.uint8).reshape(-1, rec)
        n = buf.shape[0]
        out = np.zeros((n, 8 if group == 1 else 16), dtype=np.uint64)
        ok = np.zeros(n, dtype=np.uint8)
        if group == 1:
            _check(load().b200g16_g1_decode(self.h, _ptr(buf), n, int(raw), _ptr(out), _ptr(ok)))
        else:
            _check(load().b200g16_g2_decode(self.h, _ptr(buf), n, int(raw), int(subgroup_check), _ptr(out), _ptr(ok)))
        return out, ok.astype(bool)

    def msm_plan(self, bases, n=None):
        """(window bits c, digits per scalar W) the library uses for an n-point MSM on `bases`."""
        c, w = C.c_int(), C.c_int()
        _check(load().b200g16_msm_plan(self.h, bases.handle, bases.n if n is None else n, C.byref(c), C.byref(w)))
        return c.value, w.value

    def h_pointwise_dev(self, d_a, d_b, d_c, log2n):
        _check(load().b200g16_h_pointwise_dev(self.h, _dp(d_a), _dp(d_b), _dp(d_c), log2n))

    def prove_h_dev(self, pk, d_wires, d_h, r, s):
        """prove_dev with h already computed (d_h: N elements, bit-reversed order)."""
        r, s = _u64(r).reshape(4), _u64(s).reshape(4)
        out = ProofOut()
        _check(load().b200g16_prove_h_dev(self.h, pk, _dp(d_wires), _dp(d_h), _ptr(r), _ptr(s), C.byref(out)))
        return out.as_dict()

    def prove_begin_dev(self, pk, d_wires):
        """First half of a prove: the four witness MSMs are enqueued; returns without waiting."""
        _check(load().b200g16_prove_begin_dev(self.h, pk, _dp(d_wires)))

    def prove_end_dev(self, pk, d_h, r, s):
        """Second half: Z MSM over h (device pointer), wait, assemble."""
        r, s = _u64(r).reshape(4), _u64(s).reshape(4)
        out = ProofOut()
        _check(load().b200g16_prove_end_dev(self.h, pk, _dp(d_h), _ptr(r), _ptr(s), C.byref(out)))
        return out.as_dict()

    # -- computeH split over 2 / 4 / 8 ranks (CUDA IPC peer memory; see sharded.DistributedH)
    def dist_h_init(self, log2n, n_peers, me):
        """-> uint8[192]: the IPC handles of this rank's a / b / c slices"""
        h = np.zeros(192, dtype=np.uint8)
        _check(load().b200g16_dist_h_init(self.h, log2n, n_peers, me, _ptr(h)))
        return h

    def dist_h_open(self, all_handles):
        a = np.ascontiguousarray(all_handles, dtype=np.uint8)
        _check(load().b200g16_dist_h_open(self.h, _ptr(a)))

    def dist_h_load(self, d_a, d_b, d_c):
        _check(load().b200g16_dist_h_load(self.h, _dp(d_a), _dp(d_b), _dp(d_c)))

    def dist_h_slice(self, which):
        p = load().b200g16_dist_h_slice(self.h, which)
        if not p:
            raise B200Error("dist_h_slice: not initialised")
        return int(p)

    def dist_h_phase(self, phase):
        _check(load().b200g16_dist_h_phase(self.h, phase))

    def dist_h_close(self):
        load().b200g16_dist_h_close(self.h)

    # -- MSM (host scalars: numpy (n,4) uint64 Montgomery; or a device pointer + n)
    def msm(self, bases, scalars, offset=0, n=None):
        lib = load()
        words = 8 if bases.group == 1 else 16
        out = np.zeros(words, dtype=np.uint64)
        if isinstance(scalars, np.ndarray):
            sc = _u64(scalars, 4)
            n = sc.shape[0] if n is None else n
            fn = lib.b200g16_msm_g1 if bases.group == 1 else lib.b200g16_msm_g2
            _check(fn(self.h, bases.handle, offset, _ptr(sc), n, _ptr(out)))
        else:  # raw device pointer (e.g. torch tensor .data_ptr())
            fn = lib.b200g16_msm_g1_dev if bases.group == 1 else lib.b200g16_msm_g2_dev
            _check(fn(self.h, bases.handle, offset, _dp(scalars), n, _ptr(out)))
        return out

    def msm_begin(self, bases, scalars, offset=0, n=None):
        """Enqueue a G1 MSM (b200g16_msm_g1_begin / _begin_dev) and return its ticket; msm_end(ticket) collects the
        result.  Calls made in between run behind it.  The scalars must stay alive and unchanged until msm_end."""
        lib, ticket = load(), C.c_int(-1)
        if isinstance(scalars, np.ndarray):
            sc = _u64(scalars, 4)
            n = sc.shape[0] if n is None else n
            _check(lib.b200g16_msm_g1_begin(self.h, bases.handle, offset, _ptr(sc), n, C.byref(ticket)))
            self._async_keep = getattr(self, "_async_keep", {})
            self._async_keep[ticket.value] = sc
        else:
            _check(lib.b200g16_msm_g1_begin_dev(self.h, bases.handle, offset, _dp(scalars), n, C.byref(ticket)))
        return ticket.value

    def msm_end(self, ticket):
        out = np.zeros(8, dtype=np.uint64)
        _check(load().b200g16_msm_g1_end(self.h, int(ticket), _ptr(out)))
        getattr(self, "_async_keep", {}).pop(int(ticket), None)
        return out
