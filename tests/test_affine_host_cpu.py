"""The batched-affine pair tree (gnark_whir_b200/csrc/msm_affine.cuh) compiled for the host and run thread by thread
against a direct bucket sum: pins the exact routine k_accumulate_affine executes without a GPU (the same way
tests/test_lib_cpu.py pins the device Montgomery product).  gnark-crypto counterpart: ecc/bn254/multiexp_affine.go."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import bn254 as bn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_harness", "affine_host.cc")
OUT = os.path.join(ROOT, "tests", "host_harness", "_build", "affine_host.so")


@pytest.fixture(scope="module")
def harness():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", SRC, "-I",
                    os.path.join(ROOT, "gnark_whir_b200", "csrc"), "-o", OUT], check=True)
    return ctypes.CDLL(OUT)


def _run(fn, pts, **kw):
    a = np.ascontiguousarray(pts)
    return fn(a.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(a.shape[0]), ctypes.c_uint32(kw["nb"]),
              ctypes.c_uint32(kw["mean"]), ctypes.c_int(kw["shape"]), ctypes.c_uint32(kw["seg"]), ctypes.c_uint32(kw["T"]),
              ctypes.c_int(kw["levels"]), ctypes.c_uint32(kw["min_pairs"]), ctypes.c_uint64(kw["seed"]))


@pytest.fixture(scope="module")
def g1_points():
    many = bn.g1_to_array(bn.g1_batch_mul_gen(list(range(1, 200))))
    few = bn.g1_to_array([None] + bn.g1_batch_mul_gen([5, 7, 11]))   # infinity + 3 points: tangents, cancellations
    return many, few


@pytest.mark.parametrize("T", [1, 3, 64, 1000])
@pytest.mark.parametrize("levels", [1, 2, 4])
def test_pair_tree_g1_distinct_points(harness, g1_points, T, levels):
    many, _ = g1_points
    for shape in (0, 1, 2):
        for seg in (3, 32, 100000):
            bad = _run(harness.affine_host_g1, many, nb=40, mean=20, shape=shape, seg=seg, T=T, levels=levels, min_pairs=1,
                       seed=T * 100 + levels * 10 + shape)
            assert bad == 0, (shape, seg)


@pytest.mark.parametrize("T", [1, 7, 200])
def test_pair_tree_g1_repeated_points_and_infinity(harness, g1_points, T):
    _, few = g1_points
    for shape in (0, 1, 2):
        for levels in (1, 3, 4):
            bad = _run(harness.affine_host_g1, few, nb=24, mean=30, shape=shape, seg=16, T=T, levels=levels, min_pairs=1,
                       seed=T + levels)
            assert bad == 0, (shape, levels)


def test_pair_tree_g1_no_levels_is_the_plain_chain(harness, g1_points):
    many, _ = g1_points
    for T in (5, 5000):      # more threads than entries: empty threads between the spills of a task
        assert _run(harness.affine_host_g1, many, nb=16, mean=12, shape=0, seg=8, T=T, levels=1, min_pairs=1 << 30, seed=T) == 0


def test_pair_tree_g2(harness):
    pts = bn.g2_to_array([None] + bn.g2_batch_mul_gen([1, 2, 3, 4, 5, 6, 7, 8, 9]))
    for T, levels, seg in ((1, 3, 1000), (9, 2, 5), (50, 4, 16)):
        assert _run(harness.affine_host_g2, pts, nb=12, mean=25, shape=0, seg=seg, T=T, levels=levels, min_pairs=1, seed=T) == 0
