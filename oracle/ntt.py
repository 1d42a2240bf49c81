"""Fr NTT with gnark-crypto's fft.Domain conventions, and groth16's computeH.

Oracle = test infrastructure (see oracle/__init__.py).  Restates
  gnark-crypto v0.14.1-0.20241217131346-b998989abdbe  ecc/bn254/fr/fft/{domain,fft}.go
  gnark v0.11.0  backend/groth16/bn254/prove.go  computeH
(not on disk; reached from the reference at mt.go:448 (Setup builds the Domain) and
mt.go:496 (Prove calls computeH)).  Conventions restated:
  * Cardinality N = next power of two >= m; generator w = w28^(2^(28-log2 N)),
    w28 = 2^28-th root of unity below; coset shift g = 5 (FrMultiplicativeGen)
  * FFT(a, DIF): natural-order input -> bit-reversed output (Gentleman-Sande)
  * FFT(a, DIT): bit-reversed input -> natural-order output (Cooley-Tukey)
  * OnCoset(): forward multiplies a[i] by g^i first (index bit-reversed if DIT);
    inverse multiplies by g^-i / N last (index bit-reversed if DIF)
  * FFTInverse uses w^-1 and scales by 1/N
  * computeH: iNTT(DIF) x3, coset NTT(DIT) x3, (a*b-c)/(g^N-1), coset iNTT(DIF) -> h in
    bit-reversed order, length N
"""
from __future__ import annotations

from .bn254 import R

ROOT_2_28 = 19103219067921713944291392827692070036145651957329286315305642004821462161904
COSET_GEN = 5
DIF, DIT = 0, 1


def bitrev(i, logn):
    r = 0
    for _ in range(logn):
        r = (r << 1) | (i & 1)
        i >>= 1
    return r


class Domain:
    def __init__(self, m, shift=COSET_GEN):
        n = 1
        logn = 0
        while n < m:
            n <<= 1
            logn += 1
        self.n, self.logn = n, logn
        self.gen = pow(ROOT_2_28, 1 << (28 - logn), R)
        self.gen_inv = pow(self.gen, -1, R)
        self.card_inv = pow(n, -1, R)
        self.shift = shift % R
        self.shift_inv = pow(self.shift, -1, R)

    # -- kernels
    def _dif(self, a, w):
        n = self.n
        m = n >> 1
        wm = w
        while m >= 1:
            tw = [1] * m
            for j in range(1, m):
                tw[j] = tw[j - 1] * wm % R
            for s in range(0, n, 2 * m):
                for j in range(m):
                    x, y = a[s + j], a[s + j + m]
                    a[s + j] = (x + y) % R
                    a[s + j + m] = (x - y) * tw[j] % R
            wm = wm * wm % R
            m >>= 1

    def _dit(self, a, w):
        n = self.n
        m = 1
        while m < n:
            wm = pow(w, n // (2 * m), R)
            tw = [1] * m
            for j in range(1, m):
                tw[j] = tw[j - 1] * wm % R
            for s in range(0, n, 2 * m):
                for j in range(m):
                    x, y = a[s + j], a[s + j + m] * tw[j] % R
                    a[s + j] = (x + y) % R
                    a[s + j + m] = (x - y) % R
            m <<= 1

    # -- public (mirror fft.Domain.FFT / FFTInverse; in place on a python list of ints)
    def fft(self, a, decimation, coset=False):
        assert len(a) == self.n
        if coset:
            g = 1
            pw = []
            for _ in range(self.n):
                pw.append(g)
                g = g * self.shift % R
            for i in range(self.n):
                k = bitrev(i, self.logn) if decimation == DIT else i
                a[i] = a[i] * pw[k] % R
        (self._dif if decimation == DIF else self._dit)(a, self.gen)

    def fft_inverse(self, a, decimation, coset=False):
        assert len(a) == self.n
        (self._dif if decimation == DIF else self._dit)(a, self.gen_inv)
        if not coset:
            for i in range(self.n):
                a[i] = a[i] * self.card_inv % R
            return
        g = self.card_inv
        pw = []
        for _ in range(self.n):
            pw.append(g)
            g = g * self.shift_inv % R
        for i in range(self.n):
            k = bitrev(i, self.logn) if decimation == DIF else i
            a[i] = a[i] * pw[k] % R


def dft_naive(a, w):
    """Definition: X[k] = sum_j a[j] w^(jk)  (O(N^2)); pins the fast transforms."""
    n = len(a)
    return [sum(a[j] * pow(w, j * k, R) for j in range(n)) % R for k in range(n)]


def compute_h(a, b, c, domain):
    """gnark v0.11.0 prove.go computeH; returns h (len N, bit-reversed order)."""
    n = domain.n
    a = list(a) + [0] * (n - len(a))
    b = list(b) + [0] * (n - len(b))
    c = list(c) + [0] * (n - len(c))
    for v in (a, b, c):
        domain.fft_inverse(v, DIF)
    for v in (a, b, c):
        domain.fft(v, DIT, coset=True)
    den = pow((pow(domain.shift, n, R) - 1) % R, -1, R)
    for i in range(n):
        a[i] = (a[i] * b[i] - c[i]) * den % R
    domain.fft_inverse(a, DIF, coset=True)
    return a
