"""GPU parity: Fr NTT / coset NTT / computeH (through the C-ABI) vs the python oracle that
restates gnark-crypto's fft.Domain and gnark's computeH (reached from mt.go:448,496).
Bit-exact on every output limb."""
import numpy as np
import pytest

from gnark_whir_b200 import lib
from oracle import bn254 as bn
from oracle import ntt as ont
from oracle.bn254 import R

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("logn", [0, 1, 2, 3, 5, 8, 9, 11, 12])
@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("coset", [False, True])
@pytest.mark.parametrize("dec", [ont.DIF, ont.DIT])
def test_ntt_matches_oracle(ctx, rng, logn, inverse, coset, dec):
    n = 1 << logn
    a = [rng.randrange(R) for _ in range(n)]
    d = ont.Domain(n)
    exp = list(a)
    (d.fft_inverse if inverse else d.fft)(exp, dec, coset=coset)
    got = ctx.ntt(bn.fr_to_mont_array(a), inverse=inverse, coset=coset,
                  decimation=lib.DIF if dec == ont.DIF else lib.DIT)
    assert np.array_equal(got, bn.fr_to_mont_array(exp))


@pytest.mark.parametrize("logn", [16, 17, 20])
def test_ntt_large_roundtrip_and_linearity(ctx, logn):
    """Sizes the python oracle cannot reach: size-independent properties (3 passes at 2^17+)."""
    n = 1 << logn
    rs = np.random.Generator(np.random.PCG64(logn))
    a = rs.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    f = ctx.ntt(a, decimation=lib.DIF)
    back = ctx.ntt(f, inverse=True, decimation=lib.DIT)
    assert np.array_equal(back, a)
    fc = ctx.ntt(a, coset=True, decimation=lib.DIF)
    backc = ctx.ntt(fc, inverse=True, coset=True, decimation=lib.DIT)
    assert np.array_equal(backc, a)
    # a delta at position 1 transforms to the powers of w: X[k] = w^k (definition, any size)
    delta = np.zeros((n, 4), dtype=np.uint64)
    delta[1] = bn.fr_to_mont_array([1])[0]
    fd = ctx.ntt(delta, decimation=lib.DIF)
    d = ont.Domain(n)
    for k in (0, 1, 2, n // 2 + 3, n - 1):
        assert bn.fr_from_mont_array(fd[ont.bitrev(k, logn)])[0] == pow(d.gen, k, R)
    # linearity: NTT(a + delta) = NTT(a) + NTT(delta), checked on a slice
    a2 = a.copy()
    a2[1] = bn.fr_to_mont_array([(bn.fr_from_mont_array(a[1])[0] + 1) % R])[0]
    f2 = ctx.ntt(a2, decimation=lib.DIF)
    lhs = bn.fr_from_mont_array(f2[:64])
    rhs = [(x + y) % R for x, y in zip(bn.fr_from_mont_array(f[:64]), bn.fr_from_mont_array(fd[:64]))]
    assert lhs == rhs


@pytest.mark.parametrize("n_constraints,logn", [(1, 0), (3, 2), (50, 6), (64, 6), (1000, 10), (3000, 12)])
def test_compute_h_matches_oracle(ctx, rng, n_constraints, logn):
    a = [rng.randrange(R) for _ in range(n_constraints)]
    b = [rng.randrange(R) for _ in range(n_constraints)]
    c = [x * y % R for x, y in zip(a, b)]
    exp = ont.compute_h(a, b, c, ont.Domain(1 << logn))
    got = ctx.compute_h(bn.fr_to_mont_array(a), bn.fr_to_mont_array(b), bn.fr_to_mont_array(c), logn)
    assert np.array_equal(got, bn.fr_to_mont_array(exp))


def test_compute_h_large_identity(ctx, rng):
    """2^18 constraints: check A(x)B(x) - C(x) = H(x)(x^N - 1) at a random x using device
    NTTs only for interpolation of h (bit-reversed -> natural) and python Horner for evaluation."""
    logn = 18
    n = 1 << logn
    rs = np.random.Generator(np.random.PCG64(7))

    def rnd():
        v = rs.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
        v[:, 3] &= np.uint64((1 << 60) - 1)
        return v
    a, b = rnd(), rnd()
    ai, bi = bn.fr_from_mont_array(a), bn.fr_from_mont_array(b)
    c = bn.fr_to_mont_array([x * y % R for x, y in zip(ai, bi)])
    h = ctx.compute_h(a, b, c, logn)
    x = rng.randrange(R)

    pw = [1] * n                        # x^i
    for i in range(1, n):
        pw[i] = pw[i - 1] * x % R
    rev = [ont.bitrev(i, logn) for i in range(n)]

    def eval_br(coeffs_br):             # sum c_{rev(i)} x^{rev(i)} for bit-reversed coefficients
        return sum(cv * pw[rev[i]] for i, cv in enumerate(coeffs_br)) % R

    def eval_poly_from_evals(ev):       # ev = evaluations on the domain, natural order
        return eval_br(bn.fr_from_mont_array(ctx.ntt(ev, inverse=True, decimation=lib.DIF)))
    ea = eval_poly_from_evals(a)
    eb = eval_poly_from_evals(b)
    ec = eval_poly_from_evals(c)
    hv = bn.fr_from_mont_array(h)
    eh = eval_br(hv)
    assert hv[n - 1] == 0                # top coefficient (its own bit reversal) is zero
    assert (ea * eb - ec) % R == eh * (pow(x, n, R) - 1) % R
