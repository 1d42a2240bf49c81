"""CPU tests of the product's host side: the C-ABI library loads and exports every symbol the
header declares, the host-only group helpers and the host multiplier agree with the oracle, the
device Montgomery algorithm (executed through ptx.cuh's host emulation) is pinned, and the
product never imports the oracle.  No GPU compute is issued here."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from gnark_whir_b200 import groth16 as g16
from gnark_whir_b200 import lib
from oracle import bn254 as bn
from oracle import groth16 as og
from oracle.bn254 import P, R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "b200g16.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200g16_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = lib.load()
    syms = _header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/b200g16.h but not exported"
    # and the ctypes table covers the header (so bindings cannot silently go stale)
    assert set(lib.SIGNATURES) == set(syms)
    assert L.b200g16_version() >= 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lib.B200Error) as e:
        lib.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gnark_whir_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "liboracle" not in txt and "cport" not in txt, f


def test_host_group_helpers_match_oracle(rng):
    a, b = rng.randrange(R), rng.randrange(R)
    A, B = bn.g1_mul(bn.G1_GEN, a), bn.g1_mul(bn.G1_GEN, b)
    enc = lambda p: bn.g1_to_array([p])[0]
    assert bn.g1_from_array(lib.g1_add(enc(A), enc(B)))[0] == bn.g1_add(A, B)
    assert bn.g1_from_array(lib.g1_add(enc(A), enc(A)))[0] == bn.g1_add(A, A)
    assert bn.g1_from_array(lib.g1_add(enc(A), enc(bn.g1_neg(A))))[0] is None
    assert bn.g1_from_array(lib.g1_add(enc(None), enc(A)))[0] == A
    assert bn.g1_from_array(lib.g1_scalar_mul(enc(A), bn.fr_to_mont_array([b])[0]))[0] == bn.g1_mul(A, b)
    A2, B2 = bn.g2_mul(bn.G2_GEN, a), bn.g2_mul(bn.G2_GEN, b)
    enc2 = lambda p: bn.g2_to_array([p])[0]
    assert bn.g2_from_array(lib.g2_add(enc2(A2), enc2(B2)))[0] == bn.g2_add(A2, B2)
    assert bn.g2_from_array(lib.g2_add(enc2(A2), enc2(A2)))[0] == bn.g2_add(A2, A2)
    assert bn.g2_from_array(lib.g2_scalar_mul(enc2(A2), bn.fr_to_mont_array([b])[0]))[0] == bn.g2_mul(A2, b)


@pytest.fixture(scope="module")
def fieldlib(tmp_path_factory):
    """The device field code compiled for the host (g++), PTX primitives emulated bit-exactly."""
    d = tmp_path_factory.mktemp("fieldlib")
    src = d / "f.cpp"
    src.write_text('''
#include "field.cuh"
using namespace b200;
template <class F> static void ld(F& x, const uint32_t* a) { for (int i = 0; i < 8; i++) x.l[i] = a[i]; }
template <class F> static void st(uint32_t* r, const F& x) { for (int i = 0; i < 8; i++) r[i] = x.l[i]; }
extern "C" {
void fp_mul_ptx(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fp x, y; ld(x, a); ld(y, b); st(r, Fp::mul_ptx(x, y)); }
void fr_mul_ptx(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fr x, y; ld(x, a); ld(y, b); st(r, Fr::mul_ptx(x, y)); }
void fp_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fp x, y; ld(x, a); ld(y, b); st(r, Fp::mul(x, y)); }
void fp_sqr_ptx(const uint32_t* a, uint32_t* r) { Fp x; ld(x, a); st(r, Fp::sqr_ptx(x)); }
void fp_dot2_ptx(const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d, uint32_t* r) {
  Fp x, y, z, w; ld(x, a); ld(y, b); ld(z, c); ld(w, d); st(r, Fp::dot2_ptx(x, y, z, w)); }
void fr_sqr_ptx(const uint32_t* a, uint32_t* r) { Fr x; ld(x, a); st(r, Fr::sqr_ptx(x)); }
void fp_add(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fp x, y; ld(x, a); ld(y, b); st(r, Fp::add(x, y)); }
void fp_sub(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fp x, y; ld(x, a); ld(y, b); st(r, Fp::sub(x, y)); }
void fp_neg(const uint32_t* a, uint32_t* r) { Fp x; ld(x, a); st(r, Fp::neg(x)); }
void fp_inv(const uint32_t* a, uint32_t* r) { Fp x; ld(x, a); st(r, Fp::inv(x)); }
void fr_inv(const uint32_t* a, uint32_t* r) { Fr x; ld(x, a); st(r, Fr::inv(x)); }
}
''')
    out = d / "f.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++",
                    "-I", os.path.join(ROOT, "gnark_whir_b200", "csrc"), str(src), "-o", str(out)], check=True)
    return ctypes.CDLL(str(out))


def _limbs(x):
    return (ctypes.c_uint32 * 8)(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)])


def _val(a):
    return sum(int(a[i]) << (32 * i) for i in range(8))


def test_device_montgomery_algorithm_pinned(fieldlib, rng):
    """mul_ptx is the exact mad.lo.cc/madc.hi.cc sequence the GPU runs; check it, the host
    multiplier and add/sub/neg/inv against python ints, edge cases first."""
    rip, rir = pow(1 << 256, -1, P), pow(1 << 256, -1, R)
    edge = [0, 1, 2, P - 1, P - 2, (1 << 254) % P, (1 << 256) % P]
    o = (ctypes.c_uint32 * 8)()
    for it in range(3000):
        if it < 49:
            a, b = edge[it // 7], edge[it % 7]
        else:
            a, b = rng.randrange(P), rng.randrange(P)
        fieldlib.fp_mul_ptx(_limbs(a), _limbs(b), o); assert _val(o) == a * b * rip % P
        fieldlib.fp_mul(_limbs(a), _limbs(b), o); assert _val(o) == a * b * rip % P
        fieldlib.fp_sqr_ptx(_limbs(a), o); assert _val(o) == a * a * rip % P      # the device's dedicated square
        c, d = (P - 1 - it, P - 2) if it < 49 else (rng.randrange(P), rng.randrange(P))
        fieldlib.fp_dot2_ptx(_limbs(a), _limbs(b), _limbs(c), _limbs(d), o)         # the device's fused a*b + c*d
        assert _val(o) == (a * b + c * d) * rip % P
        fieldlib.fr_sqr_ptx(_limbs(b % R), o); assert _val(o) == (b % R) ** 2 * rir % R
        fieldlib.fp_add(_limbs(a), _limbs(b), o); assert _val(o) == (a + b) % P
        fieldlib.fp_sub(_limbs(a), _limbs(b), o); assert _val(o) == (a - b) % P
        fieldlib.fp_neg(_limbs(a), o); assert _val(o) == (-a) % P
        a %= R; b %= R
        fieldlib.fr_mul_ptx(_limbs(a), _limbs(b), o); assert _val(o) == a * b * rir % R
    for a in [(1 << 253), (1 << 253) + 0xffffffff, P >> 1, 0xffffffff, (P - 1) & ~0xffffffff,
              sum(0x80000000 << (32 * k) for k in range(7))]:      # every limb's top bit set: the doubled operand's carries
        fieldlib.fp_sqr_ptx(_limbs(a % P), o); assert _val(o) == (a % P) ** 2 * rip % P
    a = rng.randrange(1, P)
    fieldlib.fp_inv(_limbs(a * (1 << 256) % P), o); assert _val(o) * rip % P == pow(a, -1, P)
    a = rng.randrange(1, R)
    fieldlib.fr_inv(_limbs(a * (1 << 256) % R), o); assert _val(o) * rir % R == pow(a, -1, R)


def test_fp64_pipe_montgomery_product_pinned(tmp_path, rng):
    """fp52.cuh: the 5 x 52-bit DFMA.RZ Montgomery product (R' = 2^260), executed here through its exact 128-bit
    emulation of fma.rz.f64 — the GPU runs the same source with the hardware instruction (b200g16_fp52_probe
    compares the two on the device).  Lazy result: congruent to a b 2^-260 and below a b / 2^260 + M."""
    src = tmp_path / "f52.cpp"
    src.write_text('''
#include <cstring>
#include "fp52.cuh"
using namespace b200;
extern "C" {
void fp52_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fp52::mul(Fp52::from_words(a), Fp52::from_words(b)).to_words(r); }
void fr52_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fr52::mul(Fr52::from_words(a), Fr52::from_words(b)).to_words(r); }
}
''')
    out = tmp_path / "f52.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++",
                    "-I", os.path.join(ROOT, "gnark_whir_b200", "csrc"), str(src), "-o", str(out)], check=True)
    L = ctypes.CDLL(str(out))
    o = (ctypes.c_uint32 * 8)()
    top = (1 << 52) - 1
    for fn, m in ((L.fp52_mul, P), (L.fr52_mul, R)):
        ri = pow(1 << 260, -1, m)
        edge = [0, 1, m - 1, m, 2 * m, (1 << 256) - 1, top, top << 52, sum(top << (52 * k) for k in range(4)) | (0xffffffffffff << 208)]
        cases = [(a, b) for a in edge for b in edge] + [(rng.randrange(1 << 256), rng.randrange(1 << 256)) for _ in range(3000)]
        for a, b in cases:
            fn(_limbs(a), _limbs(b), o)
            v = _val(o)
            assert v % m == a * b * ri % m and v <= (a * b >> 260) + m


def test_host_mirror_layout_and_hash_agree_with_oracle(rng):
    vals = [0, 1, R - 1] + [rng.randrange(R) for _ in range(5)]
    assert np.array_equal(g16.fr_array(vals), bn.fr_to_mont_array(vals))
    pt = bn.g1_mul(bn.G1_GEN, rng.randrange(R))
    assert np.array_equal(g16.g1_point(pt), bn.g1_to_array([pt])[0])
    q = bn.g2_mul(bn.G2_GEN, rng.randrange(R))
    assert np.array_equal(g16.g2_point(q), bn.g2_to_array([q])[0])
    pub = [rng.randrange(R) for _ in range(2)]
    assert g16.commitment_challenge(bn.g1_to_array([pt])[0], pub) == og.commitment_challenge(pt, pub)
    assert g16.commitment_challenge(np.zeros(8, np.uint64), pub) == og.commitment_challenge(None, pub)
    r1cs, w = og.synthetic_r1cs(9, 2, rng)
    assert g16.solve_abc(r1cs, w) == og.solve_abc(r1cs, w)
    # Verify needs a device context: no CPU fallback behind the host mirror
    with pytest.raises(AttributeError):
        g16.Verify(None, g16.Proof(None, None, None), g16.VerifyingKey(None, None, None, None, None), [])


def test_header_is_valid_c_and_library_refuses_to_run_without_a_gpu(tmp_path):
    """include/b200g16.h compiled as plain C11 (what cgo does), linked against the built library, run:
    host-only helpers answer, b200g16_init fails loudly without a device (no CPU fallback)."""
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    libdir = os.path.join(root, "gnark_whir_b200")
    exe = str(tmp_path / "abi_check")
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(root, "include"),
                    os.path.join(here, "host_harness", "abi_check.c"), "-L", libdir, "-lb200g16",
                    "-Wl,-rpath," + libdir, "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    import torch
    if not torch.cuda.is_available():
        assert out.stdout.startswith("nogpu:") and "no CPU fallback" in out.stdout
