"""Groth16 over BN254 with gnark v0.11.0's conventions (Setup / Prove / Verify).

Oracle = test infrastructure (see oracle/__init__.py).  Restates
gnark v0.11.0 backend/groth16/bn254/{setup,prove,verify}.go, reached from the reference at
mt.go:448 (groth16.Setup), mt.go:496 (groth16.Prove), mt.go:497 (groth16.Verify); plus
gnark-crypto ecc/bn254/fr/pedersen (BSB22 commitment) and fr/hash_to_field (RFC 9380
expand_message_xmd/SHA-256, L=48).  Not on disk -> restated from the published algorithm;
PARITY UNPINNED by the reference (no tests there).  Pins used instead: independent pairing
check (verify()), and closed-form proof elements from the toxic waste (closed_form_proof()).

Differences from gnark that are deliberate, because gnark hides the randomness:
  * setup() takes the toxic waste explicitly; prove() takes r, s explicitly.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field

from . import bn254 as bn
from .bn254 import R
from .ntt import Domain, bitrev, compute_h


@dataclass
class R1CS:
    """wires: [0, nb_public) public (wire 0 is the constant 1), the rest private.
    constraints: list of (L, R, O), each a list of (wire, coeff)."""
    nb_wires: int
    nb_public: int
    constraints: list
    # optional single BSB22 commitment
    private_committed: list = field(default_factory=list)   # wire ids committed to (private)
    public_committed: list = field(default_factory=list)    # public wire ids hashed into the challenge
    commitment_wire: int = -1                               # wire that receives the challenge

    @property
    def nb_constraints(self):
        return len(self.constraints)


def lc_eval(lc, w):
    return sum(c * w[i] for i, c in lc) % R


def solve_abc(r1cs, w):
    a = [lc_eval(L, w) for L, _, _ in r1cs.constraints]
    b = [lc_eval(Rr, w) for _, Rr, _ in r1cs.constraints]
    c = [lc_eval(O, w) for _, _, O in r1cs.constraints]
    return a, b, c


def is_satisfied(r1cs, w):
    a, b, c = solve_abc(r1cs, w)
    return all(x * y % R == z for x, y, z in zip(a, b, c))


# ------------------------------------------------------------------ hash_to_field (RFC 9380 5.3.1)
def expand_message_xmd(msg, dst, length):
    ell = (length + 31) // 32
    dst_prime = dst + bytes([len(dst)])
    z_pad = bytes(64)
    l_i_b = length.to_bytes(2, "big")
    b0 = hashlib.sha256(z_pad + msg + l_i_b + b"\x00" + dst_prime).digest()
    b1 = hashlib.sha256(b0 + b"\x01" + dst_prime).digest()
    out = b1
    bi = b1
    for i in range(2, ell + 1):
        bi = hashlib.sha256(bytes(x ^ y for x, y in zip(b0, bi)) + bytes([i]) + dst_prime).digest()
        out += bi
    return out[:length]


def fr_hash(msg, dst, count=1):
    L = 48
    u = expand_message_xmd(msg, dst, count * L)
    return [int.from_bytes(u[i * L:(i + 1) * L], "big") % R for i in range(count)]


def g1_marshal(pt):
    if pt is None:
        return bytes(64)            # G1Affine.Marshal() = RawBytes(): uncompressed infinity is all zero on bn254
    return pt[0].to_bytes(32, "big") + pt[1].to_bytes(32, "big")


def commitment_challenge(commitment, public_committed_values):
    msg = g1_marshal(commitment) + b"".join(int(v % R).to_bytes(32, "big") for v in public_committed_values)
    return fr_hash(msg, b"bsb22-commitment", 1)[0]


# ------------------------------------------------------------------ keys
@dataclass
class ProvingKey:
    domain: Domain
    alpha1: tuple
    beta1: tuple
    delta1: tuple
    A: list            # G1, wires with A_i(tau)=0 removed
    B: list            # G1, wires with B_i(tau)=0 removed
    Z: list            # G1, bit-reversed, length N-1
    K: list            # G1, private non-committed, non-commitment wires in wire order
    beta2: tuple
    delta2: tuple
    B2: list           # G2, same filter as B
    infinity_a: list
    infinity_b: list
    k_wires: list      # wire ids behind K (derived; gnark recomputes this in prove)
    ped_basis: list = field(default_factory=list)
    ped_basis_exp_sigma: list = field(default_factory=list)


@dataclass
class VerifyingKey:
    alpha1: tuple
    beta2: tuple
    gamma2: tuple
    delta2: tuple
    K: list            # public wires, then the commitment wire (if any)
    ped_g: tuple = None
    ped_g_sigma_neg: tuple = None
    public_committed: list = field(default_factory=list)
    has_commitment: bool = False


@dataclass
class Proof:
    Ar: tuple
    Bs: tuple
    Krs: tuple
    commitments: list = field(default_factory=list)
    commitment_pok: tuple = None


@dataclass
class ToxicWaste:
    tau: int
    alpha: int
    beta: int
    gamma: int
    delta: int
    sigma: int = 7     # pedersen


def wire_polys_at_tau(r1cs, domain, tau):
    """A_i(tau), B_i(tau), C_i(tau) per wire (setup.go setupABC)."""
    n = domain.n
    A = [0] * r1cs.nb_wires
    B = [0] * r1cs.nb_wires
    C = [0] * r1cs.nb_wires
    zn = (pow(tau, n, R) - 1) * domain.card_inv % R
    wj = 1
    for L, Rr, O in r1cs.constraints:
        lag = zn * wj % R * pow((tau - wj) % R, -1, R) % R
        for i, c in L:
            A[i] = (A[i] + c * lag) % R
        for i, c in Rr:
            B[i] = (B[i] + c * lag) % R
        for i, c in O:
            C[i] = (C[i] + c * lag) % R
        wj = wj * domain.gen % R
    return A, B, C


def setup(r1cs, tw):
    domain = Domain(r1cs.nb_constraints)
    n = domain.n
    A, B, C = wire_polys_at_tau(r1cs, domain, tw.tau)
    gi = pow(tw.gamma, -1, R)
    di = pow(tw.delta, -1, R)
    committed = set(r1cs.private_committed)
    k_priv, k_wires, k_pub, k_ped = [], [], [], []
    for i in range(r1cs.nb_wires):
        k = (tw.beta * A[i] + tw.alpha * B[i] + C[i]) % R
        if i < r1cs.nb_public:
            k_pub.append(k * gi % R)
        elif i in committed:
            pass
        elif i == r1cs.commitment_wire:
            pass
        else:
            k_priv.append(k * di % R)
            k_wires.append(i)
    if r1cs.commitment_wire >= 0:
        i = r1cs.commitment_wire
        k_pub.append((tw.beta * A[i] + tw.alpha * B[i] + C[i]) * gi % R)
        for i in r1cs.private_committed:
            k_ped.append((tw.beta * A[i] + tw.alpha * B[i] + C[i]) * gi % R)
    zdt = (pow(tw.tau, n, R) - 1) * di % R
    Zs = [zdt * pow(tw.tau, i, R) % R for i in range(n)]
    Zs = [Zs[bitrev(i, domain.logn)] for i in range(n)][:n - 1]
    inf_a = [a == 0 for a in A]
    inf_b = [b == 0 for b in B]
    g1s = bn.g1_batch_mul_gen(
        [a for a in A if a] + [b for b in B if b] + Zs + k_priv + k_pub + k_ped
        + [tw.alpha, tw.beta, tw.delta])
    na, nb_ = sum(1 for a in A if a), sum(1 for b in B if b)
    o = 0
    pkA = g1s[o:o + na]; o += na
    pkB = g1s[o:o + nb_]; o += nb_
    pkZ = g1s[o:o + n - 1]; o += n - 1
    pkK = g1s[o:o + len(k_priv)]; o += len(k_priv)
    vkK = g1s[o:o + len(k_pub)]; o += len(k_pub)
    ped = g1s[o:o + len(k_ped)]; o += len(k_ped)
    alpha1, beta1, delta1 = g1s[o:o + 3]
    g2s = bn.g2_batch_mul_gen([b for b in B if b] + [tw.beta, tw.delta, tw.gamma])
    pkB2 = g2s[:nb_]
    beta2, delta2, gamma2 = g2s[nb_:]
    pk = ProvingKey(domain, alpha1, beta1, delta1, pkA, pkB, pkZ, pkK, beta2, delta2, pkB2,
                    inf_a, inf_b, k_wires)
    vk = VerifyingKey(alpha1, beta2, gamma2, delta2, vkK)
    if r1cs.commitment_wire >= 0:
        pk.ped_basis = ped
        pk.ped_basis_exp_sigma = [bn.g1_mul(p, tw.sigma) for p in ped]
        vk.ped_g = bn.G2_GEN
        vk.ped_g_sigma_neg = bn.g2_neg(bn.g2_mul(bn.G2_GEN, tw.sigma))
        vk.public_committed = list(r1cs.public_committed)
        vk.has_commitment = True
    return pk, vk


def commit_and_fill(r1cs, pk, w):
    """What gnark's solver does when it reaches the commitment hint: Pedersen-commit to the
    committed wires and write the challenge into the commitment wire.  Returns the commitment."""
    vals = [w[i] for i in r1cs.private_committed]
    com = bn.g1_msm(pk.ped_basis, vals)
    w[r1cs.commitment_wire] = commitment_challenge(com, [w[i] for i in r1cs.public_committed])
    return com


def prove(r1cs, pk, w, r, s, commitment=None, msm=bn.g1_msm, msm2=bn.g2_msm):
    a, b, c = solve_abc(r1cs, w)
    h = compute_h(a, b, c, pk.domain)
    n = pk.domain.n
    wa = [w[i] for i in range(r1cs.nb_wires) if not pk.infinity_a[i]]
    wb = [w[i] for i in range(r1cs.nb_wires) if not pk.infinity_b[i]]
    wk = [w[i] for i in pk.k_wires]
    ar = bn.g1_sum([msm(pk.A, wa), pk.alpha1, bn.g1_mul(pk.delta1, r)])
    bs1 = bn.g1_sum([msm(pk.B, wb), pk.beta1, bn.g1_mul(pk.delta1, s)])
    krs2 = msm(pk.Z, h[:n - 1])
    krs = msm(pk.K, wk)
    kr = (-r * s) % R
    krs = bn.g1_sum([krs, krs2, bn.g1_mul(pk.delta1, kr), bn.g1_mul(ar, s), bn.g1_mul(bs1, r)])
    bs2 = bn.g2_sum([msm2(pk.B2, wb), pk.beta2, bn.g2_mul(pk.delta2, s)])
    proof = Proof(ar, bs2, krs)
    if r1cs.commitment_wire >= 0:
        vals = [w[i] for i in r1cs.private_committed]
        proof.commitments = [commitment if commitment is not None else msm(pk.ped_basis, vals)]
        proof.commitment_pok = msm(pk.ped_basis_exp_sigma, vals)
    return proof, dict(a=a, b=b, c=c, h=h, bs1=bs1, krs2=krs2)


def verify(proof, vk, public_w):
    """public_w: values of the public wires INCLUDING wire 0 (=1)."""
    pub = list(public_w)
    extra = []
    if vk.has_commitment:
        com = proof.commitments[0]
        pub.append(commitment_challenge(com, [public_w[i] for i in vk.public_committed]))
        extra = [com]
        if not bn.pairing_check([(com, vk.ped_g_sigma_neg), (proof.commitment_pok, vk.ped_g)]):
            return False
    if len(pub) != len(vk.K):
        return False
    ksum = bn.g1_sum([bn.g1_msm(vk.K, pub)] + extra)
    return bn.pairing_check([
        (proof.Ar, proof.Bs),
        (bn.g1_neg(ksum), vk.gamma2),
        (bn.g1_neg(proof.Krs), vk.delta2),
        (bn.g1_neg(vk.alpha1), vk.beta2),
    ])


def closed_form_proof(r1cs, tw, w, r, s, h):
    """Ar, Bs, Krs as single scalar multiples of the generators, from the toxic waste — shares no
    code with any MSM.  h = bit-reversed H coefficients (as compute_h returns)."""
    domain = Domain(r1cs.nb_constraints)
    n = domain.n
    A, B, C = wire_polys_at_tau(r1cs, domain, tw.tau)
    sa = sum(x * y for x, y in zip(A, w)) % R
    sb = sum(x * y for x, y in zip(B, w)) % R
    ar = (tw.alpha + sa + r * tw.delta) % R
    bs = (tw.beta + sb + s * tw.delta) % R
    di = pow(tw.delta, -1, R)
    committed = set(r1cs.private_committed)
    sk = 0
    for i in range(r1cs.nb_public, r1cs.nb_wires):
        if i in committed or i == r1cs.commitment_wire:
            continue
        sk += (tw.beta * A[i] + tw.alpha * B[i] + C[i]) * w[i]
    hc = [h[bitrev(i, domain.logn)] for i in range(n)]
    ht = sum(c * pow(tw.tau, i, R) for i, c in enumerate(hc[:n - 1])) % R
    zt = (pow(tw.tau, n, R) - 1) % R
    krs = (sk * di + ht * zt * di + s * ar + r * bs - r * s * tw.delta) % R
    return bn.g1_mul(bn.G1_GEN, ar), bn.g2_mul(bn.G2_GEN, bs), bn.g1_mul(bn.G1_GEN, krs)


# ------------------------------------------------------------------ synthetic circuits
def synthetic_r1cs(nb_constraints, nb_public, rng, with_commitment=False, small_frac=0.7):
    """A satisfiable random R1CS shaped like a byte/bit heavy verifier circuit: each constraint is
    (lc) * (lc) = fresh wire.  Returns (r1cs, witness-without-commitment-challenge)."""
    w = [1] + [rng.randrange(256) for _ in range(nb_public - 1)]
    n_seed = 4
    for _ in range(n_seed):
        w.append(rng.randrange(R))
    commitment_wire = -1
    if with_commitment:
        commitment_wire = len(w)
        w.append(0)                       # filled by commit_and_fill
    cons = []

    def small_or_big():
        return rng.randrange(2) if rng.random() < small_frac else rng.randrange(R)

    deferred = []
    for j in range(nb_constraints):
        avail = len(w)
        pick = lambda: rng.randrange(avail) if (commitment_wire < 0) else \
            rng.choice([i for i in (rng.randrange(avail), rng.randrange(avail), 0) if i != commitment_wire])
        L = [(pick(), small_or_big() or 1) for _ in range(rng.randrange(1, 4))]
        Rr = [(pick(), small_or_big() or 1) for _ in range(rng.randrange(1, 3))]
        if rng.random() < 0.3:
            # boolean-ish constraint: product lands on a small value sometimes
            Rr = [(0, rng.randrange(2))]
        if with_commitment and j == nb_constraints - 1:
            L = [(commitment_wire, 1)]
            deferred.append((j, L, Rr))
        v = lc_eval(L, w) * lc_eval(Rr, w) % R
        w.append(v)
        cons.append((L, Rr, [(len(w) - 1, 1)]))
    r1cs = R1CS(len(w), nb_public, cons)
    if with_commitment:
        r1cs.commitment_wire = commitment_wire
        priv = [i for i in range(nb_public, len(w) - 1) if i != commitment_wire]
        r1cs.private_committed = sorted(rng.sample(priv, max(1, len(priv) // 4)))
        r1cs.public_committed = [1] if nb_public > 1 else []
    return r1cs, w


def finalize_witness(r1cs, pk, w):
    """Fill the commitment challenge and re-solve the wires that depend on it."""
    com = None
    if r1cs.commitment_wire >= 0:
        com = commit_and_fill(r1cs, pk, w)
        L, Rr, O = r1cs.constraints[-1]
        w[O[0][0]] = lc_eval(L, w) * lc_eval(Rr, w) % R
    return com
