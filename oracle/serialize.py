"""TEST ORACLE (not product code): gnark-crypto / gnark wire formats for BN254, restated in python big ints.

Follows (recalled, sources absent from /root/reference — go.mod:6-7 pins the modules):
  gnark-crypto ecc/bn254/marshal.go   G1Affine.Bytes / RawBytes / SetBytes, G2Affine likewise, Encoder slices
  gnark backend/groth16/bn254/marshal.go   Proof.WriteTo / WriteRawTo / ReadFrom
reached from the reference when a proof leaves the process (mt.go:496-497 keep it in memory; ProveKit-side
tooling serialises it).

Point encodings (big-endian field elements, 2 flag bits in the most significant bits of the first byte):
  0b00 uncompressed (X || Y; the point at infinity is X = Y = 0, i.e. an all-zero record: bn254 has only two
       spare bits, so there is no "uncompressed infinity" flag)
  0b01 COMPRESSED point at infinity (rest zero, compressed length: a decoder consumes 32 / 64 bytes)
  0b10 compressed, Y is the lexicographically smallest root        0b11 compressed, Y the largest
G1: X is 32 bytes.  G2: X = X.A1 || X.A0 (64 bytes), Y likewise; "largest" compares A1 first, then A0.
Slices: uint32 big-endian length, then the elements.
Proof: Ar (G1) | Bs (G2) | Krs (G1) | len(Commitments) u32 | Commitments... (G1) | CommitmentPok (G1).
VerifyingKey: G1.Alpha | G1.Beta | G2.Beta | G2.Gamma | G1.Delta | G2.Delta | u32 len(K) | K... |
  PublicAndCommitmentCommitted [][]uint64 (u32 outer len; per group u32 len + big-endian u64s) |
  u32 #commitment keys | per key: pedersen G | GSigmaNeg (G2).
Parity is UNPINNED by the reference (no vectors on disk); pins here are algebraic (round trips, on-curve,
sign rule) — see tests/test_oracle_cpu.py.
"""
from . import bn254 as bn
from .bn254 import P

M_UNCOMPRESSED, M_INFINITY, M_SMALLEST, M_LARGEST = 0x00, 0x40, 0x80, 0xC0
HALF = (P - 1) // 2


def fp_lex_largest(y):
    return y > HALF


def fp2_lex_largest(y):
    return y[1] > HALF if y[1] != 0 else y[0] > HALF


def fp_sqrt(a):
    r = pow(a, (P + 1) // 4, P)
    return r if r * r % P == a % P else None


def fp2_sqrt(a):
    """complex method, p = 3 mod 4; None if a is not a square"""
    a0, a1 = a[0] % P, a[1] % P
    if a1 == 0:
        r = fp_sqrt(a0)
        if r is not None:
            return (r, 0)
        r = fp_sqrt(-a0 % P)
        return None if r is None else (0, r)
    s = fp_sqrt((a0 * a0 + a1 * a1) % P)
    if s is None:
        return None
    inv2 = pow(2, -1, P)
    t = (a0 + s) * inv2 % P
    x0 = fp_sqrt(t)
    if x0 is None:
        t = (a0 - s) * inv2 % P
        x0 = fp_sqrt(t)
        if x0 is None:
            return None
    x1 = a1 * pow(2 * x0, -1, P) % P
    return (x0, x1) if bn.f2_sqr((x0, x1)) == (a0, a1) else None


# ---------------------------------------------------------------- G1
def g1_bytes(pt):
    """G1Affine.Bytes(): 32 bytes compressed"""
    if pt is None:
        return bytes([M_INFINITY]) + bytes(31)
    b = bytearray(pt[0].to_bytes(32, "big"))
    b[0] |= M_LARGEST if fp_lex_largest(pt[1]) else M_SMALLEST
    return bytes(b)


def g1_raw_bytes(pt):
    """G1Affine.RawBytes() / Marshal(): 64 bytes uncompressed"""
    if pt is None:
        return bytes(64)
    return pt[0].to_bytes(32, "big") + pt[1].to_bytes(32, "big")


def g1_set_bytes(buf):
    """G1Affine.SetBytes: -> (point, bytes consumed); raises ValueError on an invalid encoding"""
    flag = buf[0] & 0xC0
    if flag == M_INFINITY:
        n = 32  # gnark-crypto consumes the compressed size for a compressed-infinity marker
        if any(buf[1:32]) or buf[0] & 0x3F:
            raise ValueError("invalid infinity encoding")
        return None, n
    if flag == M_UNCOMPRESSED:
        if len(buf) < 64:
            raise ValueError("short buffer")
        x, y = int.from_bytes(buf[:32], "big"), int.from_bytes(buf[32:64], "big")
        if x == 0 and y == 0:
            return None, 64
        if x >= P or y >= P or not bn.g1_on_curve((x, y)):
            raise ValueError("invalid point")
        return (x, y), 64
    x = int.from_bytes(bytes([buf[0] & 0x3F]) + bytes(buf[1:32]), "big")
    if x >= P:
        raise ValueError("x not reduced")
    y = fp_sqrt((x * x * x + 3) % P)
    if y is None:
        raise ValueError("x is not on the curve")
    if fp_lex_largest(y) != (flag == M_LARGEST):
        y = P - y
    return (x, y), 32


# ---------------------------------------------------------------- G2
def g2_bytes(pt):
    if pt is None:
        return bytes([M_INFINITY]) + bytes(63)
    (x0, x1), y = pt
    b = bytearray(x1.to_bytes(32, "big") + x0.to_bytes(32, "big"))
    b[0] |= M_LARGEST if fp2_lex_largest(y) else M_SMALLEST
    return bytes(b)


def g2_raw_bytes(pt):
    if pt is None:
        return bytes(128)
    (x0, x1), (y0, y1) = pt
    return b"".join(v.to_bytes(32, "big") for v in (x1, x0, y1, y0))


def g2_set_bytes(buf, subgroup_check=True):
    flag = buf[0] & 0xC0
    if flag == M_INFINITY:
        if any(buf[1:64]) or buf[0] & 0x3F:
            raise ValueError("invalid infinity encoding")
        return None, 64
    if flag == M_UNCOMPRESSED:
        if len(buf) < 128:
            raise ValueError("short buffer")
        x1, x0, y1, y0 = (int.from_bytes(buf[32 * i:32 * i + 32], "big") for i in range(4))
        if x0 == x1 == y0 == y1 == 0:
            return None, 128
        pt = ((x0, x1), (y0, y1))
        if max(x0, x1, y0, y1) >= P or not bn.g2_on_curve(pt):
            raise ValueError("invalid point")
        n = 128
    else:
        x1 = int.from_bytes(bytes([buf[0] & 0x3F]) + bytes(buf[1:32]), "big")
        x0 = int.from_bytes(buf[32:64], "big")
        if x0 >= P or x1 >= P:
            raise ValueError("x not reduced")
        x = (x0, x1)
        y = fp2_sqrt(bn.f2_add(bn.f2_mul(bn.f2_sqr(x), x), bn.B2))
        if y is None:
            raise ValueError("x is not on the curve")
        if fp2_lex_largest(y) != (flag == M_LARGEST):
            y = bn.f2_neg(y)
        pt, n = (x, y), 64
    if subgroup_check and bn._mul(bn._F2, pt, bn.R, mod=1 << 300) is not None:
        raise ValueError("point not in the r-torsion subgroup")
    return pt, n


# ---------------------------------------------------------------- groth16 Proof
def proof_write(ar, bs, krs, commitments, pok, raw=False):
    e1, e2 = (g1_raw_bytes, g2_raw_bytes) if raw else (g1_bytes, g2_bytes)
    out = e1(ar) + e2(bs) + e1(krs) + len(commitments).to_bytes(4, "big")
    for c in commitments:
        out += e1(c)
    return out + e1(pok)


def proof_read(buf):
    o = 0
    ar, n = g1_set_bytes(buf[o:]); o += n
    bs, n = g2_set_bytes(buf[o:]); o += n
    krs, n = g1_set_bytes(buf[o:]); o += n
    k = int.from_bytes(buf[o:o + 4], "big"); o += 4
    coms = []
    for _ in range(k):
        c, n = g1_set_bytes(buf[o:]); o += n
        coms.append(c)
    pok, n = g1_set_bytes(buf[o:]); o += n
    return ar, bs, krs, coms, pok, o


# ---------------------------------------------------------------- groth16 VerifyingKey
def vk_write(alpha1, beta1, beta2, gamma2, delta1, delta2, K, committed_groups, pedersen_keys, raw=False):
    """committed_groups: list of lists of wire ids; pedersen_keys: list of (G, GSigmaNeg) G2 pairs."""
    e1, e2 = (g1_raw_bytes, g2_raw_bytes) if raw else (g1_bytes, g2_bytes)
    out = e1(alpha1) + e1(beta1) + e2(beta2) + e2(gamma2) + e1(delta1) + e2(delta2)
    out += len(K).to_bytes(4, "big") + b"".join(e1(k) for k in K)
    out += len(committed_groups).to_bytes(4, "big")
    for g in committed_groups:
        out += len(g).to_bytes(4, "big") + b"".join(int(x).to_bytes(8, "big") for x in g)
    out += len(pedersen_keys).to_bytes(4, "big")
    for g, gs in pedersen_keys:
        out += e2(g) + e2(gs)
    return out


# ---------------------------------------------------------------- fft.Domain and groth16 ProvingKey
def domain_write(domain):
    """(*fft.Domain).WriteTo as recalled: Cardinality u64 | CardinalityInv | Generator | GeneratorInv |
    FrMultiplicativeGen | FrMultiplicativeGenInv (32-byte big-endian canonical) | withPrecompute (1 byte = 1)."""
    vals = [domain.card_inv, domain.gen, domain.gen_inv, domain.shift, domain.shift_inv]
    return domain.n.to_bytes(8, "big") + b"".join(int(v).to_bytes(32, "big") for v in vals) + b"\x01"


def pk_write(pk, raw=False):
    """(*ProvingKey).WriteTo / WriteRawTo as recalled (see gnark_whir_b200/groth16.py pk_write_to) for an
    oracle.groth16.ProvingKey."""
    e1, e2 = (g1_raw_bytes, g2_raw_bytes) if raw else (g1_bytes, g2_bytes)

    def sl(points, enc):
        return len(points).to_bytes(4, "big") + b"".join(enc(q) for q in points)
    out = domain_write(pk.domain) + e1(pk.alpha1) + e1(pk.beta1) + e1(pk.delta1)
    out += sl(pk.A, e1) + sl(pk.B, e1) + sl(pk.Z, e1) + sl(pk.K, e1) + e2(pk.beta2) + e2(pk.delta2) + sl(pk.B2, e2)
    n = len(pk.infinity_a)
    out += n.to_bytes(8, "big") + sum(pk.infinity_a).to_bytes(8, "big") + sum(pk.infinity_b).to_bytes(8, "big")
    out += bytes(1 if x else 0 for x in pk.infinity_a) + bytes(1 if x else 0 for x in pk.infinity_b)
    has = 1 if pk.ped_basis else 0
    out += has.to_bytes(4, "big")
    if has:
        out += sl(pk.ped_basis, e1) + sl(pk.ped_basis_exp_sigma, e1)
    return out
