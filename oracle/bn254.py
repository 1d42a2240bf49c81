"""BN254 (alt_bn128) field / curve / pairing arithmetic on Python big ints.

Oracle = test infrastructure (see oracle/__init__.py).  Restates what the reference
reaches through gnark-crypto v0.14.1-0.20241217131346-b998989abdbe (go.mod:7):
  ecc/bn254/fp, ecc/bn254/fr          -> Fp / Fr (here: python int mod P / mod R)
  ecc/bn254/g1.go, g2.go              -> affine / Jacobian group law
  ecc/bn254/multiexp.go MultiExp      -> msm_* (result only; internals not observable)
  ecc/bn254/pairing.go                -> pairing / pairing_check (optimal ate)
The reference's own view of the Fr modulus and the 4xu64 little-endian limb layout is
at typeConverters/typeConverters.go:26-44 and utilities/utilities.go:102; the prove
call sites that reach this arithmetic are mt.go:447-497.

Memory layout helpers (to_mont_limbs & co.) produce gnark-crypto's in-memory form:
an element is [4]uint64 little-endian limbs of x*2^256 mod m (Montgomery form),
G1Affine = {X,Y}, G2Affine = {X.A0,X.A1,Y.A0,Y.A1}, infinity = all-zero.
"""
from __future__ import annotations

import numpy as np

P = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47  # base field
R = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001  # scalar field
MONT_R = 1 << 256
B1 = 3
U_BN = 4965661367192848881           # BN parameter
ATE_LOOP = 6 * U_BN + 2               # 29793968203157093288

G1_GEN = (1, 2)
G2_GEN = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)

# ----------------------------------------------------------------------------- Fp2
# Fp2 = Fp[u]/(u^2+1); element = (a0, a1)

def f2_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def f2_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def f2_neg(a):
    return ((-a[0]) % P, (-a[1]) % P)


def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def f2_sqr(a):
    return ((a[0] + a[1]) * (a[0] - a[1]) % P, 2 * a[0] * a[1] % P)


def f2_muls(a, s):
    return (a[0] * s % P, a[1] * s % P)


def f2_conj(a):
    return (a[0], (-a[1]) % P)


def f2_inv(a):
    n = pow(a[0] * a[0] + a[1] * a[1], -1, P)
    return (a[0] * n % P, (-a[1]) * n % P)


def f2_pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = f2_mul(r, a)
        a = f2_sqr(a)
        e >>= 1
    return r


F2_ZERO = (0, 0)
F2_ONE = (1, 0)
XI = (9, 1)                                   # non-residue 9+u
B2 = f2_mul((3, 0), f2_inv(XI))               # twist coefficient 3/(9+u)

# ----------------------------------------------------------------------------- generic curve ops
# A "field" is a small namespace of functions so G1 (ints) and G2 (Fp2 tuples) share code.


class _F1:
    zero = 0
    one = 1
    add = staticmethod(lambda a, b: (a + b) % P)
    sub = staticmethod(lambda a, b: (a - b) % P)
    mul = staticmethod(lambda a, b: a * b % P)
    sqr = staticmethod(lambda a: a * a % P)
    neg = staticmethod(lambda a: (-a) % P)
    inv = staticmethod(lambda a: pow(a, -1, P))
    muls = staticmethod(lambda a, s: a * s % P)
    is_zero = staticmethod(lambda a: a == 0)
    b = B1


class _F2:
    zero = F2_ZERO
    one = F2_ONE
    add = staticmethod(f2_add)
    sub = staticmethod(f2_sub)
    mul = staticmethod(f2_mul)
    sqr = staticmethod(f2_sqr)
    neg = staticmethod(f2_neg)
    inv = staticmethod(f2_inv)
    muls = staticmethod(f2_muls)
    is_zero = staticmethod(lambda a: a == F2_ZERO)
    b = B2


def _on_curve(F, pt):
    if pt is None:
        return True
    x, y = pt
    return F.sqr(y) == F.add(F.mul(F.sqr(x), x), F.b)


def _jac_dbl(F, p):
    X, Y, Z = p
    if F.is_zero(Z):
        return p
    A = F.sqr(X)
    Bv = F.sqr(Y)
    C = F.sqr(Bv)
    D = F.muls(F.sub(F.sub(F.sqr(F.add(X, Bv)), A), C), 2)
    E = F.muls(A, 3)
    Fv = F.sqr(E)
    X3 = F.sub(Fv, F.muls(D, 2))
    Y3 = F.sub(F.mul(E, F.sub(D, X3)), F.muls(C, 8))
    Z3 = F.muls(F.mul(Y, Z), 2)
    return (X3, Y3, Z3)


def _jac_add(F, p, q):
    X1, Y1, Z1 = p
    X2, Y2, Z2 = q
    if F.is_zero(Z1):
        return q
    if F.is_zero(Z2):
        return p
    Z1Z1 = F.sqr(Z1)
    Z2Z2 = F.sqr(Z2)
    U1 = F.mul(X1, Z2Z2)
    U2 = F.mul(X2, Z1Z1)
    S1 = F.mul(F.mul(Y1, Z2), Z2Z2)
    S2 = F.mul(F.mul(Y2, Z1), Z1Z1)
    if U1 == U2:
        if S1 == S2:
            return _jac_dbl(F, p)
        return (F.one, F.one, F.zero)
    H = F.sub(U2, U1)
    Rr = F.sub(S2, S1)
    HH = F.sqr(H)
    HHH = F.mul(H, HH)
    V = F.mul(U1, HH)
    X3 = F.sub(F.sub(F.sqr(Rr), HHH), F.muls(V, 2))
    Y3 = F.sub(F.mul(Rr, F.sub(V, X3)), F.mul(S1, HHH))
    Z3 = F.mul(F.mul(Z1, Z2), H)
    return (X3, Y3, Z3)


def _to_jac(F, pt):
    if pt is None:
        return (F.one, F.one, F.zero)
    return (pt[0], pt[1], F.one)


def _to_aff(F, p):
    X, Y, Z = p
    if F.is_zero(Z):
        return None
    zi = F.inv(Z)
    zi2 = F.sqr(zi)
    return (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))


def _neg(F, pt):
    if pt is None:
        return None
    return (pt[0], F.neg(pt[1]))


def _mul(F, pt, k, mod=R):
    k %= mod
    acc = (F.one, F.one, F.zero)
    if pt is None or k == 0:
        return None
    base = _to_jac(F, pt)
    for bit in bin(k)[2:]:
        acc = _jac_dbl(F, acc)
        if bit == "1":
            acc = _jac_add(F, acc, base)
    return _to_aff(F, acc)


def _add(F, a, b):
    return _to_aff(F, _jac_add(F, _to_jac(F, a), _to_jac(F, b)))


def _msm_naive(F, points, scalars):
    acc = (F.one, F.one, F.zero)
    for pt, s in zip(points, scalars):
        q = _mul(F, pt, s)
        acc = _jac_add(F, acc, _to_jac(F, q))
    return _to_aff(F, acc)


def _msm_bucket(F, points, scalars, c=8):
    """Plain (unsigned-digit) bucket method — independent of the CUDA/C code's signed-digit
    Pippenger; used only to make python MSMs over a few thousand points bearable."""
    scalars = [s % R for s in scalars]
    nwin = (254 + c - 1) // c
    total = (F.one, F.one, F.zero)
    for w in reversed(range(nwin)):
        for _ in range(c):
            total = _jac_dbl(F, total)
        buckets = [None] * (1 << c)
        for pt, s in zip(points, scalars):
            if pt is None:
                continue
            d = (s >> (w * c)) & ((1 << c) - 1)
            if d:
                jp = _to_jac(F, pt)
                buckets[d] = jp if buckets[d] is None else _jac_add(F, buckets[d], jp)
        run = (F.one, F.one, F.zero)
        acc = (F.one, F.one, F.zero)
        for d in range((1 << c) - 1, 0, -1):
            if buckets[d] is not None:
                run = _jac_add(F, run, buckets[d])
            acc = _jac_add(F, acc, run)
        total = _jac_add(F, total, acc)
    return _to_aff(F, total)


# ----------------------------------------------------------------------------- G1 / G2 front ends
def g1_on_curve(pt): return _on_curve(_F1, pt)
def g2_on_curve(pt): return _on_curve(_F2, pt)
def g1_add(a, b): return _add(_F1, a, b)
def g2_add(a, b): return _add(_F2, a, b)
def g1_neg(a): return _neg(_F1, a)
def g2_neg(a): return _neg(_F2, a)
def g1_mul(pt, k): return _mul(_F1, pt, k)
def g2_mul(pt, k): return _mul(_F2, pt, k)
def g1_msm_naive(points, scalars): return _msm_naive(_F1, points, scalars)
def g2_msm_naive(points, scalars): return _msm_naive(_F2, points, scalars)
def g1_msm(points, scalars, c=8): return _msm_bucket(_F1, points, scalars, c)
def g2_msm(points, scalars, c=8): return _msm_bucket(_F2, points, scalars, c)


def g1_sum(points):
    acc = (1, 1, 0)
    for pt in points:
        acc = _jac_add(_F1, acc, _to_jac(_F1, pt))
    return _to_aff(_F1, acc)


def g2_sum(points):
    acc = (F2_ONE, F2_ONE, F2_ZERO)
    for pt in points:
        acc = _jac_add(_F2, acc, _to_jac(_F2, pt))
    return _to_aff(_F2, acc)


def g1_batch_mul_gen(scalars, c=8):
    """[k_i]G1 for many k_i (fixed-base windowed table); returns affine list."""
    return _batch_mul_gen(_F1, G1_GEN, scalars, c)


def g2_batch_mul_gen(scalars, c=8):
    return _batch_mul_gen(_F2, G2_GEN, scalars, c)


def _batch_mul_gen(F, gen, scalars, c):
    nwin = (254 + c - 1) // c
    table = []                                   # table[w][d] = d * 2^(w c) * G  (jacobian)
    base = _to_jac(F, gen)
    for w in range(nwin):
        row = [(F.one, F.one, F.zero)]
        for d in range(1, 1 << c):
            row.append(_jac_add(F, row[-1], base))
        table.append(row)
        for _ in range(c):
            base = _jac_dbl(F, base)
    out = []
    for s in scalars:
        s %= R
        acc = (F.one, F.one, F.zero)
        for w in range(nwin):
            d = (s >> (w * c)) & ((1 << c) - 1)
            if d:
                acc = _jac_add(F, acc, table[w][d])
        out.append(acc)
    # batch normalise (Montgomery trick)
    zs = [p[2] for p in out]
    prefix = []
    run = F.one
    for z in zs:
        prefix.append(run)
        if not F.is_zero(z):
            run = F.mul(run, z)
    inv = F.inv(run)
    res = [None] * len(out)
    for i in range(len(out) - 1, -1, -1):
        z = zs[i]
        if F.is_zero(z):
            continue
        zi = F.mul(inv, prefix[i])
        inv = F.mul(inv, z)
        zi2 = F.sqr(zi)
        res[i] = (F.mul(out[i][0], zi2), F.mul(out[i][1], F.mul(zi2, zi)))
    return res


# ----------------------------------------------------------------------------- pairing (optimal ate)
# Fp12 = Fp2[w]/(w^6 - XI); element = list of 6 Fp2 coefficients.
F12_ONE = [F2_ONE] + [F2_ZERO] * 5


def f12_mul(a, b):
    t = [[0, 0] for _ in range(11)]
    for i in range(6):
        ai0, ai1 = a[i]
        if ai0 == 0 and ai1 == 0:
            continue
        for j in range(6):
            bj0, bj1 = b[j]
            t[i + j][0] += ai0 * bj0 - ai1 * bj1
            t[i + j][1] += ai0 * bj1 + ai1 * bj0
    out = []
    for k in range(6):
        c0, c1 = t[k]
        if k < 5:
            h0, h1 = t[k + 6]
            c0 += 9 * h0 - h1              # (h0 + h1 u)(9 + u)
            c1 += 9 * h1 + h0
        out.append((c0 % P, c1 % P))
    return out


def f12_pow(a, e):
    r = F12_ONE
    while e:
        if e & 1:
            r = f12_mul(r, a)
        a = f12_mul(a, a)
        e >>= 1
    return r


_GAMMA_X = f2_pow(XI, (P - 1) // 3)      # w^(2(p-1))
_GAMMA_Y = f2_pow(XI, (P - 1) // 2)      # w^(3(p-1))


def _twist_frob(q):
    return (f2_mul(f2_conj(q[0]), _GAMMA_X), f2_mul(f2_conj(q[1]), _GAMMA_Y))


def _line(T, Q, Pt):
    """Line through twist points T,Q (tangent if equal) evaluated at P in G1, as sparse Fp12,
    and T+Q.  Untwist (x,y)->(x w^2, y w^3):  l = yP - lam*xP*w + (lam*xT - yT)*w^3."""
    xp, yp = Pt
    if T[0] == Q[0] and T[1] == Q[1]:
        lam = f2_mul(f2_muls(f2_sqr(T[0]), 3), f2_inv(f2_muls(T[1], 2)))
    else:
        lam = f2_mul(f2_sub(Q[1], T[1]), f2_inv(f2_sub(Q[0], T[0])))
    x3 = f2_sub(f2_sub(f2_sqr(lam), T[0]), Q[0])
    y3 = f2_sub(f2_mul(lam, f2_sub(T[0], x3)), T[1])
    l = [(yp % P, 0), f2_neg(f2_muls(lam, xp)), F2_ZERO,
         f2_sub(f2_mul(lam, T[0]), T[1]), F2_ZERO, F2_ZERO]
    return l, (x3, y3)


def miller_loop(Pt, Q):
    if Pt is None or Q is None:
        return F12_ONE
    f = F12_ONE
    T = Q
    for bit in bin(ATE_LOOP)[3:]:
        l, T2 = _line(T, T, Pt)
        f = f12_mul(f12_mul(f, f), l)
        T = T2
        if bit == "1":
            l, T = _line(T, Q, Pt)
            f = f12_mul(f, l)
    Q1 = _twist_frob(Q)
    Q2 = g2_neg(_twist_frob(Q1))
    l, T = _line(T, Q1, Pt)
    f = f12_mul(f, l)
    l, _ = _line(T, Q2, Pt)
    f = f12_mul(f, l)
    return f


_FINAL_EXP = (P ** 12 - 1) // R


def final_exp(f):
    return f12_pow(f, _FINAL_EXP)


def pairing(Pt, Q):
    return final_exp(miller_loop(Pt, Q))


def pairing_check(pairs):
    """prod e(P_i, Q_i) == 1"""
    f = F12_ONE
    for Pt, Q in pairs:
        f = f12_mul(f, miller_loop(Pt, Q))
    return final_exp(f) == F12_ONE


# ----------------------------------------------------------------------------- gnark-crypto memory layout
_MASK64 = (1 << 64) - 1


def int_to_limbs(x):
    return [(x >> (64 * i)) & _MASK64 for i in range(4)]


def limbs_to_int(l):
    return int(l[0]) | int(l[1]) << 64 | int(l[2]) << 128 | int(l[3]) << 192


def fr_to_mont_array(vals):
    """list of ints mod R -> (n,4) uint64 Montgomery limbs"""
    return np.array([int_to_limbs(v % R * MONT_R % R) for v in vals], dtype=np.uint64).reshape(-1, 4)


def fr_from_mont_array(arr):
    rinv = pow(MONT_R, -1, R)
    return [limbs_to_int(row) * rinv % R for row in np.asarray(arr, dtype=np.uint64).reshape(-1, 4)]


def fp_to_mont_limbs(x):
    return int_to_limbs(x % P * MONT_R % P)


def fp_from_mont_limbs(l):
    return limbs_to_int(l) * pow(MONT_R, -1, P) % P


def g1_to_array(points):
    """affine list -> (n,8) uint64 : X limbs, Y limbs (Montgomery); None -> zeros"""
    out = np.zeros((len(points), 8), dtype=np.uint64)
    for i, pt in enumerate(points):
        if pt is not None:
            out[i, :4] = fp_to_mont_limbs(pt[0])
            out[i, 4:] = fp_to_mont_limbs(pt[1])
    return out


def g1_from_array(arr):
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 8)
    out = []
    for row in arr:
        if not row.any():
            out.append(None)
        else:
            out.append((fp_from_mont_limbs(row[:4]), fp_from_mont_limbs(row[4:])))
    return out


def g2_to_array(points):
    out = np.zeros((len(points), 16), dtype=np.uint64)
    for i, pt in enumerate(points):
        if pt is not None:
            (x0, x1), (y0, y1) = pt
            out[i, 0:4] = fp_to_mont_limbs(x0)
            out[i, 4:8] = fp_to_mont_limbs(x1)
            out[i, 8:12] = fp_to_mont_limbs(y0)
            out[i, 12:16] = fp_to_mont_limbs(y1)
    return out


def g2_from_array(arr):
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 16)
    out = []
    for row in arr:
        if not row.any():
            out.append(None)
        else:
            out.append(((fp_from_mont_limbs(row[0:4]), fp_from_mont_limbs(row[4:8])),
                        (fp_from_mont_limbs(row[8:12]), fp_from_mont_limbs(row[12:16]))))
    return out
