"""CPU pin of the pairing code the GPU verify kernel runs (gnark_whir_b200/csrc/pairing.cuh, compiled for
the host by tests/host_harness/pairing_host.cc) against the oracle's independent optimal-ate pairing
(oracle/bn254.py: affine Miller loop + naive (p^12-1)/r exponentiation in python big ints).
The harness is test infrastructure; libb200g16 itself only runs this code on the GPU."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import bn254 as bn
from oracle.bn254 import P, R

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_harness", "pairing_host.cc")
OUT = os.path.join(HERE, "host_harness", "_build", "pairing_host.so")
CSRC = os.path.join(HERE, "..", "gnark_whir_b200", "csrc")
U = bn.U_BN
COFACTOR = 2 * U * (6 * U * U + 3 * U + 1)


@pytest.fixture(scope="module")
def H():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", SRC, "-I", CSRC, "-o", OUT], check=True)
    return C.CDLL(OUT)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _flat(f):
    out = np.zeros((6, 8), dtype=np.uint64)
    for i, (a0, a1) in enumerate(f):
        out[i, :4] = bn.fp_to_mont_limbs(a0)
        out[i, 4:] = bn.fp_to_mont_limbs(a1)
    return out


def _unflat(arr):
    return [(bn.fp_from_mont_limbs(arr[i, :4]), bn.fp_from_mont_limbs(arr[i, 4:])) for i in range(6)]


def gt_from_gnark_layout(arr):
    """(6,8) u64 in gnark's E12 order C0.B0,C0.B1,C0.B2,C1.B0,C1.B1,C1.B2 -> oracle's w-power coefficients"""
    arr = np.asarray(arr, dtype=np.uint64).reshape(6, 8)
    f = [None] * 6
    for k, idx in enumerate([0, 2, 4, 1, 3, 5]):
        f[idx] = (bn.fp_from_mont_limbs(arr[k, :4]), bn.fp_from_mont_limbs(arr[k, 4:]))
    return f


def test_fp12_tower_matches_oracle(H):
    rng = random.Random(11)
    out = np.zeros((6, 8), dtype=np.uint64)
    for _ in range(4):
        a = [(rng.randrange(P), rng.randrange(P)) for _ in range(6)]
        b = [(rng.randrange(P), rng.randrange(P)) for _ in range(6)]
        H.host_f12_mul(_ptr(_flat(a)), _ptr(_flat(b)), _ptr(out))
        assert _unflat(out) == bn.f12_mul(a, b)
        H.host_f12_inv(_ptr(_flat(a)), _ptr(out))
        assert bn.f12_mul(_unflat(out), a) == bn.F12_ONE
        H.host_f12_frob2(_ptr(_flat(a)), _ptr(out))
        assert _unflat(out) == bn.f12_pow(a, P * P)
        H.host_f12_sqr(_ptr(_flat(a)), _ptr(out))
        assert _unflat(out) == bn.f12_mul(a, a)
        for k in (1, 2, 3):
            H.host_f12_frob(_ptr(_flat(a)), k, _ptr(out))
            assert _unflat(out) == bn.f12_pow(a, P ** k)


def test_pairing_value_and_cofactor(H):
    rng = random.Random(12)
    ka, kb = rng.randrange(1, R), rng.randrange(1, R)
    Pa, Qb = bn.g1_mul(bn.G1_GEN, ka), bn.g2_mul(bn.G2_GEN, kb)
    g1, g2 = bn.g1_to_array([Pa]), bn.g2_to_array([Qb])
    gt = np.zeros((6, 8), dtype=np.uint64)
    assert H.host_pair(_ptr(g1), _ptr(g2), 1, 0, _ptr(gt)) == 0
    exp = bn.pairing(Pa, Qb)
    assert gt_from_gnark_layout(gt) == exp                      # the reduced pairing itself
    H.host_pair(_ptr(g1), _ptr(g2), 1, 1, _ptr(gt))
    assert gt_from_gnark_layout(gt) == bn.f12_pow(exp, COFACTOR)   # with gnark's final-exp cofactor


def test_pairing_products_and_point_checks(H):
    rng = random.Random(13)
    ka, kb = rng.randrange(1, R), rng.randrange(1, R)
    Pa, Qb = bn.g1_mul(bn.G1_GEN, ka), bn.g2_mul(bn.G2_GEN, kb)
    gt = np.zeros((6, 8), dtype=np.uint64)
    g2 = bn.g2_to_array([Qb, bn.G2_GEN])
    good = bn.g1_to_array([Pa, bn.g1_neg(bn.g1_mul(bn.G1_GEN, ka * kb % R))])
    bad = bn.g1_to_array([Pa, bn.g1_neg(bn.g1_mul(bn.G1_GEN, (ka * kb + 1) % R))])
    assert H.host_pair(_ptr(good), _ptr(g2), 2, 0, _ptr(gt)) == 1
    assert H.host_pair(_ptr(bad), _ptr(g2), 2, 0, _ptr(gt)) == 0
    # infinity on either side contributes 1
    inf1 = np.zeros((1, 8), dtype=np.uint64)
    assert H.host_pair(_ptr(inf1), _ptr(bn.g2_to_array([Qb])), 1, 0, _ptr(gt)) == 1
    assert H.host_g1_on_curve(_ptr(bn.g1_to_array([Pa]))) == 1
    off = bn.g1_to_array([(Pa[0], (Pa[1] + 1) % P)])
    assert H.host_g1_on_curve(_ptr(off)) == 0
    assert H.host_g2_in_subgroup(_ptr(bn.g2_to_array([Qb]))) == 1
    # a twist point outside the r-torsion: pick x until x^3 + b' is a square in Fp2, do NOT clear the cofactor
    x = (5, 1)
    while True:
        rhs = bn.f2_add(bn.f2_mul(bn.f2_sqr(x), x), bn.B2)
        y = _f2_sqrt(rhs)
        if y is not None:
            break
        x = (x[0] + 1, x[1])
    assert bn.g2_on_curve((x, y))
    assert H.host_g2_in_subgroup(_ptr(bn.g2_to_array([(x, y)]))) == 0


def _f2_sqrt(a):
    """square root in Fp2 = Fp[u]/(u^2+1), p = 3 mod 4 (complex method); None if a is not a square"""
    a0, a1 = a
    if a1 == 0:
        r = pow(a0, (P + 1) // 4, P)
        if r * r % P == a0:
            return (r, 0)
        r = pow(-a0 % P, (P + 1) // 4, P)
        return (0, r) if r * r % P == -a0 % P else None
    n = (a0 * a0 + a1 * a1) % P
    s = pow(n, (P + 1) // 4, P)
    if s * s % P != n:
        return None
    for sign in (1, -1):
        t = (a0 + sign * s) * pow(2, -1, P) % P
        x0 = pow(t, (P + 1) // 4, P)
        if x0 * x0 % P == t and x0:
            x1 = a1 * pow(2 * x0, -1, P) % P
            if bn.f2_sqr((x0, x1)) == (a0 % P, a1 % P):
                return (x0, x1)
    return None
