"""GPU parity at BASELINE.json's full sizes, through size-independent properties (the oracle cannot
finish these sizes in seconds, except where the C port's closed forms can):

  configs[1]  G1 MSM 2^24: result == [sum s_i k_i] G (dot product by the C oracle, one scalar mul),
              and linearity MSM(s) + MSM(t) == MSM(s + t) with s + t formed on the host
  configs[2]  Fr NTT 2^24: forward/inverse round trips in both conventions, linearity; 2^26 (the largest size of the sweep):
              definition check on a sparse input + round trips (full-output comparisons
              of NTT 2^22 / 2^24 and computeH 2^22 against the C port live in test_gpu_bench_sizes.py)
  configs[3]  Keccak: 2^20 Merkle paths, a random sample compared with the C oracle + all roots of paths
              opened from one real tree equal that tree's root
"""
import numpy as np
import pytest
import torch

from gnark_whir_b200 import lib
from oracle import bn254 as bn
from oracle import cport
from oracle.bn254 import R

pytestmark = pytest.mark.gpu


def _rand_fr(rs, n):
    a = rs.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    return a


def test_msm_g1_2p24_closed_form_and_linearity(ctx):
    rs = np.random.Generator(np.random.PCG64(2024))
    n = 1 << 24
    ks = _rand_fr(rs, n)
    bases = ctx.fixed_base_mul(bn.g1_to_array([bn.G1_GEN])[0], ks, group=1, resident=True)
    s, t = _rand_fr(rs, n), _rand_fr(rs, n)
    ds, dt = torch.from_numpy(s.view(np.int64)).cuda(), torch.from_numpy(t.view(np.int64)).cuda()
    ms = ctx.msm(bases, ds.data_ptr(), n=n)
    assert np.array_equal(ms, cport.g1_gen_mul(cport.fr_dot(ks, s)))
    bases.precompute(0)                                   # the same vector with a window table
    assert np.array_equal(ctx.msm(bases, ds.data_ptr(), n=n), ms)
    mt = ctx.msm(bases, dt.data_ptr(), n=n)
    # s + t as plain 256-bit integers (both < 2^252, so no reduction): Montgomery form is additive
    carry = np.zeros(n, dtype=np.uint64)
    st = np.empty_like(s)
    for j in range(4):
        a = s[:, j].astype(object) + t[:, j].astype(object) + carry.astype(object)
        st[:, j] = (a & ((1 << 64) - 1)).astype(np.uint64)
        carry = (a >> 64).astype(np.uint64)
    dst = torch.from_numpy(st.view(np.int64)).cuda()
    assert np.array_equal(ctx.msm(bases, dst.data_ptr(), n=n), lib.g1_add(ms, mt))
    bases.free()


def test_ntt_2p24_round_trips_and_linearity(ctx):
    L, n = 24, 1 << 24
    g = torch.Generator(device="cuda").manual_seed(7)

    def rnd():
        a = torch.randint(0, 1 << 62, (n, 4), dtype=torch.int64, device="cuda", generator=g)
        a[:, 3] &= (1 << 60) - 1
        return a
    a = rnd()
    ref = a.clone()
    torch.cuda.synchronize()      # the library runs on its own streams: torch's fill and clone must have finished
    ctx.ntt_dev(a.data_ptr(), L, decimation=lib.DIF)
    assert not torch.equal(a, ref)
    ctx.ntt_dev(a.data_ptr(), L, inverse=True, decimation=lib.DIT)
    assert torch.equal(a, ref)
    ctx.ntt_dev(a.data_ptr(), L, coset=True, decimation=lib.DIT)     # input taken as bit-reversed
    ctx.ntt_dev(a.data_ptr(), L, inverse=True, coset=True, decimation=lib.DIF)
    assert torch.equal(a, ref)
    # linearity: NTT(x) + NTT(y) == NTT(x + y); x + y formed as plain integers (both < 2^252, no reduction)
    x, y = ref, rnd()
    xs, ys = x.cpu().numpy().view(np.uint64), y.cpu().numpy().view(np.uint64)
    carry = np.zeros(n, dtype=np.uint64)
    out = np.empty_like(xs)
    for j in range(4):
        lo = xs[:, j] + ys[:, j]
        c1 = (lo < xs[:, j]).astype(np.uint64)
        lo2 = lo + carry
        c2 = (lo2 < lo).astype(np.uint64)
        out[:, j] = lo2
        carry = c1 + c2
    s = torch.from_numpy(out.view(np.int64)).cuda()
    torch.cuda.synchronize()
    for v in (x, y, s):
        ctx.ntt_dev(v.data_ptr(), L, decimation=lib.DIF)
    # field addition of the two spectra on the host for a strided sample
    idx = np.arange(0, n, 4099)
    X = bn.fr_from_mont_array(x[idx].cpu().numpy().view(np.uint64))
    Y = bn.fr_from_mont_array(y[idx].cpu().numpy().view(np.uint64))
    S = bn.fr_from_mont_array(s[idx].cpu().numpy().view(np.uint64))
    assert all((p + q) % R == r for p, q, r in zip(X, Y, S))


def test_ntt_2p26_largest_size_definition_check_and_round_trips(ctx):
    """BASELINE.json configs[2] sweeps up to 2^26.  At that size: (1) a sparse input (a handful of nonzero coefficients)
    makes EVERY output a short closed-form sum, X[k] = sum_t v_t w^(j_t k), which python big ints evaluate exactly for a
    sample of k — a check against the transform's definition with gnark's DIF output order (bit-reversed); (2) dense
    round trips in both conventions."""
    from oracle import ntt as ontt
    L, n = 26, 1 << 26
    dom = ontt.Domain(n)
    rs = np.random.Generator(np.random.PCG64(26))
    pos = [0, 1, n // 2 + 3, n - 1] + [int(x) for x in rs.integers(0, n, size=4)]
    vals = [int(x) for x in rs.integers(1, 1 << 62, size=len(pos))]
    a = torch.zeros((n, 4), dtype=torch.int64, device="cuda")
    a[torch.tensor(pos, device="cuda")] = torch.from_numpy(bn.fr_to_mont_array(vals).view(np.int64)).cuda()
    torch.cuda.synchronize()
    ctx.ntt_dev(a.data_ptr(), L, decimation=lib.DIF)          # natural in, bit-reversed out
    ks = [0, 1, 2, n // 2, n - 1] + [int(x) for x in rs.integers(0, n, size=59)]
    got = bn.fr_from_mont_array(a[torch.tensor([ontt.bitrev(k, L) for k in ks], device="cuda")].cpu().numpy().view(np.uint64))
    for k, g in zip(ks, got):
        assert g == sum(v * pow(dom.gen, (j * k) % n, R) for j, v in zip(pos, vals)) % R
    del a
    g = torch.Generator(device="cuda").manual_seed(26)
    a = torch.randint(0, 1 << 62, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    a[:, 3] &= (1 << 60) - 1
    ref = a.clone()
    torch.cuda.synchronize()      # the library runs on its own streams: torch's fill and clone must have finished
    ctx.ntt_dev(a.data_ptr(), L, decimation=lib.DIF)
    assert not torch.equal(a, ref)
    ctx.ntt_dev(a.data_ptr(), L, inverse=True, decimation=lib.DIT)
    assert torch.equal(a, ref)
    ctx.ntt_dev(a.data_ptr(), L, coset=True, decimation=lib.DIT)
    ctx.ntt_dev(a.data_ptr(), L, inverse=True, coset=True, decimation=lib.DIF)
    assert torch.equal(a, ref)


def test_keccak_merkle_2p20_paths(ctx):
    from oracle import keccak as ok
    rs = np.random.Generator(np.random.PCG64(5))
    q, height, leaf_len = 1 << 20, 20, 512
    leaves = rs.integers(0, 256, size=(q, leaf_len), dtype=np.uint8)
    sib = rs.integers(0, 256, size=(q, 32), dtype=np.uint8)
    auth = rs.integers(0, 256, size=(q, height - 1, 32), dtype=np.uint8)
    idx = rs.integers(0, 1 << height, size=q, dtype=np.uint64)
    roots, _ = ctx.keccak_merkle_paths(leaves, sib, auth, idx)
    pick = rs.integers(0, q, size=256)
    exp = cport.merkle_paths(leaves[pick], sib[pick], auth[pick], idx[pick])
    assert np.array_equal(roots[pick], exp)
    # a real tree (2^10 leaves): every opened path recomputes the tree's root
    h2 = 10
    tl = [bytes(rs.integers(0, 256, size=64, dtype=np.uint8)) for _ in range(1 << h2)]
    lv = ok.build_merkle_tree(tl)
    opened = [ok.merkle_open(lv, i) for i in range(1 << h2)]
    roots, okf = ctx.keccak_merkle_paths(
        np.stack([np.frombuffer(x, np.uint8) for x in tl]),
        np.stack([np.frombuffer(o[0], np.uint8) for o in opened]),
        np.stack([np.frombuffer(b"".join(o[1]), np.uint8).reshape(h2 - 1, 32) for o in opened]),
        np.arange(1 << h2, dtype=np.uint64), expected_root=lv[-1][0])
    assert okf.all()


def test_pageable_host_buffers_take_the_staged_copy_and_match(ctx):
    """Large buffers in ordinary (pageable) host memory — what a Go slice is — are staged by the library's own host
    threads (csrc/host_copy.cuh).  An odd-sized MSM (the last chunk is partial, the upload pieces are uneven) and an
    NTT through the host-pointer entry points must equal the same calls on device-resident inputs."""
    rs = np.random.Generator(np.random.PCG64(77))
    n = (1 << 21) + 4321                                   # 64 MiB + a bit of scalars: two geometric pieces, 8 stripes
    ks, sc = _rand_fr(rs, n), _rand_fr(rs, n)
    bases = ctx.fixed_base_mul(bn.g1_to_array([bn.G1_GEN])[0], ks, group=1, resident=True)
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    want = ctx.msm(bases, d_sc.data_ptr(), n=n)
    assert np.array_equal(ctx.msm(bases, sc), want)         # numpy array = pageable
    assert np.array_equal(want, cport.g1_gen_mul(cport.fr_dot(ks, sc)))
    bases.free()
    logn = 21
    a = _rand_fr(rs, 1 << logn)
    d_a = torch.from_numpy(a.view(np.int64)).cuda()
    ctx.ntt_dev(d_a.data_ptr(), logn, decimation=lib.DIF)
    got = ctx.ntt(a, decimation=lib.DIF)                    # host-pointer entry point: 64 MiB up, 64 MiB back
    assert np.array_equal(got, d_a.cpu().numpy().view(np.uint64))
    # Keccak-f batch through host memory: 40 MB of states up and back, an odd count (partial last chunk)
    nst = 200_003
    st = rs.integers(0, 1 << 63, size=(nst, 25), dtype=np.uint64)
    d_st = torch.from_numpy(st.view(np.int64)).cuda()
    ctx.keccak_f_batch_dev(d_st.data_ptr(), nst)
    assert np.array_equal(ctx.keccak_f_batch(st), d_st.cpu().numpy().view(np.uint64))
