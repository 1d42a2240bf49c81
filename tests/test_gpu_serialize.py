"""GPU parity: gnark-crypto point encodings and the groth16 Proof wire format through the C-ABI
(b200g16_g1/g2_encode, b200g16_g1/g2_decode; gnark_whir_b200.groth16.proof_write_to / proof_read_from)
against the python restatement oracle/serialize.py.  Byte-exact both ways."""
import random

import numpy as np
import pytest

from gnark_whir_b200 import groth16 as g16
from oracle import bn254 as bn
from oracle import serialize as ser
from oracle.bn254 import R

pytestmark = pytest.mark.gpu


def _points(rng, n):
    g1 = [bn.g1_mul(bn.G1_GEN, rng.randrange(1, R)) for _ in range(n)]
    g2 = [bn.g2_mul(bn.G2_GEN, rng.randrange(1, R)) for _ in range(n)]
    g1 += [bn.g1_neg(p) for p in g1[:n // 2]] + [None, bn.G1_GEN]
    g2 += [bn.g2_neg(p) for p in g2[:n // 2]] + [None, bn.G2_GEN]
    return g1, g2


@pytest.mark.parametrize("raw", [False, True])
def test_encode_decode_match_oracle(ctx, raw):
    rng = random.Random(41 + raw)
    g1, g2 = _points(rng, 24)
    e1 = ctx.encode_points(bn.g1_to_array(g1), group=1, raw=raw)
    e2 = ctx.encode_points(bn.g2_to_array(g2), group=2, raw=raw)
    o1 = [ser.g1_raw_bytes(p) if raw else ser.g1_bytes(p) for p in g1]
    o2 = [ser.g2_raw_bytes(p) if raw else ser.g2_bytes(p) for p in g2]
    assert [bytes(r) for r in e1] == o1
    assert [bytes(r) for r in e2] == o2
    d1, ok1 = ctx.decode_points(b"".join(o1), group=1, raw=raw)
    d2, ok2 = ctx.decode_points(b"".join(o2), group=2, raw=raw)
    assert ok1.all() and ok2.all()
    assert np.array_equal(d1, bn.g1_to_array(g1)) and np.array_equal(d2, bn.g2_to_array(g2))


def test_decode_rejects_invalid_records(ctx):
    rng = random.Random(43)
    good = ser.g1_bytes(bn.g1_mul(bn.G1_GEN, 7))
    bad_x = next(bytes([0x80]) + x.to_bytes(31, "big") for x in range(2, 200)
                 if ser.fp_sqrt((x ** 3 + 3) % bn.P) is None)                    # x^3 + 3 not a square
    too_big = bytes([0xBF]) + bytes([0xFF] * 31)                                  # x >= p
    bad_inf = bytes([0x40]) + bytes(30) + bytes([1])                              # infinity flag with payload
    uncompressed_in_compressed = bytes([0x00]) + bytes(31)
    pts, ok = ctx.decode_points(good + bad_x + too_big + bad_inf + uncompressed_in_compressed, group=1)
    assert list(ok) == [True, False, False, False, False]
    assert not pts[1:].any()
    # raw record that is not on the curve
    p = bn.g1_mul(bn.G1_GEN, 9)
    off = p[0].to_bytes(32, "big") + ((p[1] + 1) % bn.P).to_bytes(32, "big")
    _, ok = ctx.decode_points(ser.g1_raw_bytes(p) + off, group=1, raw=True)
    assert list(ok) == [True, False]
    # raw infinity is the all-zero record (bn254 has no "uncompressed infinity" flag); the compressed-infinity flag
    # is invalid inside a raw batch (gnark's decoder would consume only 32 bytes for it)
    assert ser.g1_raw_bytes(None) == bytes(64) and ser.g2_raw_bytes(None) == bytes(128)
    pts, ok = ctx.decode_points(bytes(64) + bytes([0x40]) + bytes(63), group=1, raw=True)
    assert list(ok) == [True, False] and not pts.any()
    pts, ok = ctx.decode_points(bytes(128) + bytes([0x40]) + bytes(127), group=2, raw=True)
    assert list(ok) == [True, False] and not pts.any()
    # G2: a twist point outside the r-torsion subgroup passes without the subgroup check only
    x = (5, 1)
    while True:
        y = ser.fp2_sqrt(bn.f2_add(bn.f2_mul(bn.f2_sqr(x), x), bn.B2))
        if y is not None:
            break
        x = (x[0] + 1, x[1])
    enc = ser.g2_bytes((x, y))
    _, ok = ctx.decode_points(enc + ser.g2_bytes(bn.G2_GEN), group=2, subgroup_check=True)
    assert list(ok) == [False, True]
    pts, ok = ctx.decode_points(enc, group=2, subgroup_check=False)
    assert ok[0] and np.array_equal(pts[0], bn.g2_to_array([(x, y)])[0])


@pytest.mark.parametrize("with_commitment", [False, True])
@pytest.mark.parametrize("raw", [False, True])
def test_proof_wire_format_round_trip(ctx, raw, with_commitment):
    rng = random.Random(47)
    ar, krs, com, pok = (bn.g1_mul(bn.G1_GEN, rng.randrange(1, R)) for _ in range(4))
    bs = bn.g2_mul(bn.G2_GEN, rng.randrange(1, R))
    coms = [com] if with_commitment else []
    pk = pok if with_commitment else None
    proof = g16.Proof(bn.g1_to_array([ar])[0], bn.g1_to_array([krs])[0], bn.g2_to_array([bs])[0],
                      [bn.g1_to_array([c])[0] for c in coms], bn.g1_to_array([pk])[0])
    data = g16.proof_write_to(ctx, proof, raw=raw)
    assert data == ser.proof_write(ar, bs, krs, coms, pk, raw=raw)               # byte-exact with the oracle
    back = g16.proof_read_from(ctx, data)               # per-point flag bits: no raw / compressed hint needed
    if not with_commitment:                              # CommitmentPok = infinity: 0x40 record or 64 zero bytes
        assert data.endswith(bytes(64) if raw else bytes([0x40]) + bytes(31))
    assert np.array_equal(back.Ar, proof.Ar) and np.array_equal(back.Bs, proof.Bs) and np.array_equal(back.Krs, proof.Krs)
    assert len(back.Commitments) == len(coms) and np.array_equal(back.CommitmentPok, proof.CommitmentPok)
    with pytest.raises(ValueError):
        g16.proof_read_from(ctx, data[:-1], raw=raw)
    tampered = bytearray(data)
    tampered[5] ^= 0x55
    try:                                                                          # either rejected or a different point
        other = g16.proof_read_from(ctx, bytes(tampered), raw=raw)
        assert not np.array_equal(other.Ar, proof.Ar)
    except ValueError:
        pass


def test_decode_large_batch_matches_encode(ctx):
    """2^16 compressed G1 points: decode(encode(P)) == P (the key-loading path)."""
    rs = np.random.Generator(np.random.PCG64(5))
    ks = rs.integers(0, 1 << 62, size=(1 << 16, 4), dtype=np.uint64)
    ks[:, 3] &= np.uint64((1 << 60) - 1)
    pts = ctx.fixed_base_mul(bn.g1_to_array([bn.G1_GEN])[0], ks, group=1)
    enc = ctx.encode_points(pts, group=1)
    dec, ok = ctx.decode_points(enc, group=1)
    assert ok.all() and np.array_equal(dec, pts)


@pytest.mark.parametrize("with_commitment", [False, True])
@pytest.mark.parametrize("raw", [False, True])
def test_verifying_key_wire_format(ctx, raw, with_commitment):
    """Setup's vk through vk_write_to: byte-exact with the oracle's layout, and vk_read_from gives back a key
    that verifies a proof made with the original."""
    from oracle import groth16 as og
    rng = random.Random(53)
    r1cs, w = og.synthetic_r1cs(20, 3, rng, with_commitment=with_commitment)
    tw = og.ToxicWaste(*[rng.randrange(1, R) for _ in range(5)], sigma=rng.randrange(1, R))
    pk, vk = g16.Setup(ctx, r1cs, g16.ToxicWaste(tw.tau, tw.alpha, tw.beta, tw.gamma, tw.delta, tw.sigma))
    try:
        _, ovk = og.setup(r1cs, tw)
        data = g16.vk_write_to(ctx, vk, raw=raw)
        exp = ser.vk_write(ovk.alpha1, bn.g1_mul(bn.G1_GEN, tw.beta), ovk.beta2, ovk.gamma2,
                           bn.g1_mul(bn.G1_GEN, tw.delta), ovk.delta2, ovk.K,
                           [list(ovk.public_committed)] if with_commitment else [],
                           [(ovk.ped_g, ovk.ped_g_sigma_neg)] if with_commitment else [], raw=raw)
        assert data == exp
        back = g16.vk_read_from(ctx, data, raw=raw)
        for f in ("G1_Alpha", "G1_K", "G2_Beta", "G2_Gamma", "G2_Delta", "G1_Beta", "G1_Delta"):
            assert np.array_equal(getattr(back, f), getattr(vk, f)), f
        assert back.has_commitment == with_commitment

        def resolve(wit):
            L, Rr, O = r1cs.constraints[-1]
            wit[O[0][0]] = og.lc_eval(L, wit) * og.lc_eval(Rr, wit) % R
        proof = g16.Prove(ctx, r1cs, pk, w, resolve=resolve if with_commitment else None)
        g16.Verify(ctx, proof, back, proof.debug["witness"][1:r1cs.nb_public])
        with pytest.raises(ValueError):
            g16.vk_read_from(ctx, data[:-3], raw=raw)
    finally:
        pk.free()


@pytest.mark.parametrize("with_commitment", [False, True])
@pytest.mark.parametrize("raw", [False, True])
def test_proving_key_wire_format(ctx, raw, with_commitment):
    """Setup's pk through pk_write_to: byte-exact with the oracle's layout of (*ProvingKey).WriteTo / WriteRawTo, and
    pk_read_from gives back a key that proves identically (batched GPU decode: one square root per compressed point)."""
    from oracle import groth16 as og
    rng = random.Random(59 + raw)
    r1cs, w = og.synthetic_r1cs(40, 3, rng, with_commitment=with_commitment)
    tw = og.ToxicWaste(*[rng.randrange(1, R) for _ in range(5)], sigma=rng.randrange(1, R))
    pk, vk = g16.Setup(ctx, r1cs, g16.ToxicWaste(tw.tau, tw.alpha, tw.beta, tw.gamma, tw.delta, tw.sigma))
    opk, _ = og.setup(r1cs, tw)
    try:
        data = g16.pk_write_to(ctx, pk, raw=raw)
        assert data == ser.pk_write(opk, raw=raw)
        back = g16.pk_read_from(ctx, data, k_skip=pk.k_skip)
        for f in ("G1_Alpha", "G1_Beta", "G1_Delta", "G1_A", "G1_B", "G1_Z", "G1_K", "G2_Beta", "G2_Delta", "G2_B", "InfinityA", "InfinityB"):
            assert np.array_equal(np.asarray(getattr(back, f)), np.asarray(getattr(pk, f))), f
        assert back.log2_domain == pk.log2_domain and len(back.CommitmentKeys) == len(pk.CommitmentKeys)
        if with_commitment:
            assert np.array_equal(back.CommitmentKeys[0].Basis, pk.CommitmentKeys[0].Basis)
            assert np.array_equal(back.CommitmentKeys[0].BasisExpSigma, pk.CommitmentKeys[0].BasisExpSigma)
        for cut in (7, 8 + 5 * 32, len(data) - 1):            # truncated streams are errors, never a partial key
            with pytest.raises(ValueError):
                g16.pk_read_from(ctx, data[:cut])
        bad = bytearray(data)
        bad[3] ^= 1                                            # cardinality no longer a power of two
        with pytest.raises(ValueError):
            g16.pk_read_from(ctx, bytes(bad))
        if not with_commitment:                                # the re-read key proves bit-identically
            r, s_ = rng.randrange(R), rng.randrange(R)
            p0 = g16.Prove(ctx, r1cs, pk, list(w), r=r, s=s_)
            p1 = g16.Prove(ctx, r1cs, back, list(w), r=r, s=s_)
            assert np.array_equal(p0.Ar, p1.Ar) and np.array_equal(p0.Bs, p1.Bs) and np.array_equal(p0.Krs, p1.Krs)
    finally:
        pk.free()
        back.free() if "back" in dir() else None
