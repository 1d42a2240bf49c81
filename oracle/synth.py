"""Synthetic Groth16 proving keys with KNOWN discrete logs + the closed-form answers of a prove over them.
TEST INFRASTRUCTURE (used by tests/ and by bench.py's pre-timing parity gates only; never by the product).

A real pk's points are tau-dependent multiples of the generators; for a full-size parity check any points with
known discrete logs do: with A_i = ka_i * G etc. every MultiExp of the prover (gnark
backend/groth16/bn254/prove.go, reached from /root/reference/mt.go:496) has the closed form [sum w_i k_i] G,
which the C restatement computes as one Fr dot product and one scalar multiplication - independent of any
bucket method - and h comes from the C restatement of computeH.  Ar, Bs, Krs then follow in plain integer
arithmetic on the discrete logs.

Witness shape: 40% zero/one, 30% bytes, 30% uniform field elements.
"""
from __future__ import annotations

import numpy as np

from . import bn254 as bn
from . import cport
from .bn254 import R

G1 = bn.g1_to_array([bn.G1_GEN])[0]
G2 = bn.g2_to_array([bn.G2_GEN])[0]


def rand_fr(rs, n):
    a = rs.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)        # < 2^252 < r: every row is a valid Montgomery residue
    return a


def whir_mix(rs, n):
    """(n, 4) Montgomery scalars: 40% in {0, 1}, 30% bytes, 30% uniform."""
    out = rand_fr(rs, n)
    u = rs.random(n)
    small = rs.integers(0, 256, size=n)
    tbl = bn.fr_to_mont_array(list(range(256)))
    m01, mb = u < 0.4, (u >= 0.4) & (u < 0.7)
    out[m01] = tbl[small[m01] & 1]
    out[mb] = tbl[small[mb]]
    return out


def _int(x):
    return bn.fr_from_mont_array(np.asarray(x, dtype=np.uint64).reshape(1, 4))[0]


class KnownDlogKey:
    """pk of 2^log2n constraints and as many wires over points with known discrete logs, resident on `ctx`'s GPU.
    Wire 0 is public (excluded from K); no point at infinity in A / B."""

    def __init__(self, ctx, log2n, seed, precompute=True, lo_hi=None):
        """lo_hi: optional dict name -> (lo, hi) to build only a point-range shard of each vector
        (names a, b, k, z); discrete logs are still drawn for the whole vectors so shards of the same seed agree."""
        self.ctx, self.L, self.N = ctx, log2n, 1 << log2n
        N = self.N
        rs = np.random.Generator(np.random.PCG64(seed))
        self.k = {"a": rand_fr(rs, N), "b": rand_fr(rs, N), "k": rand_fr(rs, N - 1), "z": rand_fr(rs, N - 1),
                  "b2": None}
        self.k["b2"] = rand_fr(rs, N)
        self.small = rand_fr(rs, 5)             # dlogs of alpha, beta, delta (G1), beta2, delta2 (G2)
        self.spans = {n: (0, len(self.k[n])) for n in ("a", "b", "k", "z")}
        if lo_hi:
            self.spans.update(lo_hi)
        sp = self.spans
        self.vec = {n: ctx.fixed_base_mul(G1, self.k[n][sp[n][0]:sp[n][1]], group=1, resident=True) for n in ("a", "b", "k", "z")}
        self.vec["b2"] = ctx.fixed_base_mul(G2, self.k["b2"][sp["b"][0]:sp["b"][1]], group=2, resident=True)
        if precompute:
            for v in self.vec.values():
                v.precompute(0)
        g1s = ctx.fixed_base_mul(G1, self.small[:3], group=1)
        g2s = ctx.fixed_base_mul(G2, self.small[3:], group=2)
        self.points = dict(alpha=g1s[0], beta=g1s[1], delta=g1s[2], beta2=g2s[0], delta2=g2s[1])
        self.k_skip = np.zeros(N, dtype=np.uint8)
        self.k_skip[0] = 1
        self.zeros = np.zeros(N, dtype=np.uint8)
        partial = lo_hi is not None
        self.handle = ctx.pk_upload(log2n, N, self.vec["a"], self.vec["b"], self.vec["k"], self.vec["z"], self.vec["b2"],
                                    g1s[0], g1s[1], g1s[2], g2s[0], g2s[1], self.zeros, self.zeros, self.k_skip,
                                    partial=partial, offsets=(sp["a"][0], sp["b"][0], sp["k"][0], sp["z"][0]))

    def free(self):
        self.ctx.pk_free(self.handle)
        for v in self.vec.values():
            v.free()

    # ---- closed forms
    def expected_msms(self, wires, h, nthreads=0):
        """The five complete MultiExp results (affine arrays) + their discrete logs."""
        N = self.N
        d = {"a": cport.fr_dot(self.k["a"], wires, nthreads), "b1": cport.fr_dot(self.k["b"], wires, nthreads),
             "k": cport.fr_dot(self.k["k"], wires[1:], nthreads), "z": cport.fr_dot(self.k["z"], h[:N - 1], nthreads),
             "b2": cport.fr_dot(self.k["b2"], wires, nthreads)}
        pts = {"msm_a": cport.g1_gen_mul(d["a"]), "msm_b1": cport.g1_gen_mul(d["b1"]), "msm_k": cport.g1_gen_mul(d["k"]),
               "msm_z": cport.g1_gen_mul(d["z"]), "msm_b2": cport.g2_gen_mul(d["b2"])}
        return pts, {k: _int(v) for k, v in d.items()}

    def expected(self, wires, a, b, c, r, s, nthreads=0):
        """h (C restatement of computeH), the five MultiExp results and Ar / Bs / Krs in closed form."""
        h = cport.compute_h(a, b, c, self.L, nthreads)
        pts, d = self.expected_msms(wires, h, nthreads)
        al, be, de, be2, de2 = [_int(x) for x in self.small]
        ri, si = _int(r), _int(s)
        ar = (d["a"] + al + ri * de) % R
        bs1 = (d["b1"] + be + si * de) % R
        krs = (d["k"] + d["z"] - ri * si * de + si * ar + ri * bs1) % R
        bs = (d["b2"] + be2 + si * de2) % R
        pts["ar"] = cport.g1_gen_mul(bn.fr_to_mont_array([ar]))
        pts["bs1"] = cport.g1_gen_mul(bn.fr_to_mont_array([bs1]))
        pts["krs"] = cport.g1_gen_mul(bn.fr_to_mont_array([krs]))
        pts["bs"] = cport.g2_gen_mul(bn.fr_to_mont_array([bs]))
        return pts, h


def check_proof(got, exp, names=("msm_a", "msm_b1", "msm_k", "msm_z", "msm_b2", "ar", "bs", "krs", "bs1")):
    """Names whose GPU value differs from the closed form (empty list = parity)."""
    return [n for n in names if not np.array_equal(np.asarray(got[n], dtype=np.uint64).reshape(-1), exp[n].reshape(-1))]
