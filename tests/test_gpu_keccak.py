"""GPU parity: batched Keccak-f / duplex sponge / Merkle-path recompute vs the oracle that
restates keccakSponge/keccakSponge.go and mtUtilities.go:109-141.  Byte-exact."""
import os

import numpy as np
import pytest

from oracle import keccak as ok

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000])
def test_keccak_f_batch(ctx, n):
    rs = np.random.Generator(np.random.PCG64(n))
    st = rs.integers(0, 1 << 63, size=(n, 25), dtype=np.uint64) * np.uint64(2) + rs.integers(0, 2, size=(n, 25), dtype=np.uint64)
    if n >= 1:
        st[0, :] = 0                       # FIPS-202 zero-state known answer
    got = ctx.keccak_f_batch(st)
    assert int(got[0, 0]) == 0xF1258F7940E1DDE7
    for i in range(min(n, 40)):
        assert [int(v) for v in got[i]] == ok.keccak_f([int(v) for v in st[i]])


@pytest.mark.parametrize("in_len,out_len", [(0, 32), (1, 1), (7, 32), (8, 8), (64, 32), (135, 32), (136, 32),
                                            (137, 32), (512, 32), (300, 200), (272, 136), (100, 137)])
def test_sponge_batch(ctx, in_len, out_len):
    n = 37
    data = np.frombuffer(os.urandom(n * max(in_len, 1)), dtype=np.uint8)[: n * in_len].reshape(n, in_len)
    got = ctx.keccak_sponge_batch(data, out_len)
    for i in range(n):
        s = ok.Sponge()
        s.absorb(bytes(data[i]))
        assert bytes(got[i]) == s.squeeze(out_len), (i, in_len, out_len)


@pytest.mark.parametrize("height,leaf_len", [(1, 64), (4, 512), (10, 512), (5, 32), (3, 136), (3, 144)])
def test_merkle_paths(ctx, height, leaf_len):
    nleaves = 1 << height
    leaves = [os.urandom(leaf_len) for _ in range(nleaves)]
    levels = ok.build_merkle_tree(leaves)
    root = levels[-1][0]
    idxs = list(range(nleaves)) if nleaves <= 64 else [0, 1, 2, nleaves - 1, nleaves // 2, 77, 500, 1023]
    L, S, A = [], [], []
    for i in idxs:
        sib, ap = ok.merkle_open(levels, i)
        assert ok.merkle_root_from_path(leaves[i], sib, ap, i) == root
        L.append(np.frombuffer(leaves[i], dtype=np.uint8))
        S.append(np.frombuffer(sib, dtype=np.uint8))
        A.append(np.frombuffer(b"".join(ap), dtype=np.uint8).reshape(height - 1, 32))
    roots, okf = ctx.keccak_merkle_paths(np.stack(L), np.stack(S), np.stack(A).reshape(len(idxs), height - 1, 32),
                                         np.array(idxs, dtype=np.uint64), expected_root=root)
    assert all(bytes(r) == root for r in roots)
    assert okf.all()
    # a corrupted sibling / wrong index must not verify
    S2 = np.stack(S).copy()
    S2[0, 0] ^= 1
    roots2, ok2 = ctx.keccak_merkle_paths(np.stack(L), S2, np.stack(A).reshape(len(idxs), height - 1, 32),
                                          np.array(idxs, dtype=np.uint64), expected_root=root)
    assert not ok2[0] and ok2[1:].all()


@pytest.mark.parametrize("cfg", [10, 11, 20, 21, 40, 41])
def test_merkle_paths_warp_kernel_variants(ctx, monkeypatch, cfg):
    """The latency kernel (one warp per path) in every shape it can be launched in: 1 / 2 / 4 warps per CTA, lanes
    exchanged by shuffles (x0) or through shared memory (x1) — all byte-exact against the oracle, including a leaf that
    ends in a partial block and one that is a whole number of blocks."""
    monkeypatch.setenv("B200G16_MERKLE_WARP", str(cfg))
    for height, leaf_len in ((6, 512), (3, 136), (2, 24), (5, 272)):
        nleaves = 1 << height
        leaves = [os.urandom(leaf_len) for _ in range(nleaves)]
        levels = ok.build_merkle_tree(leaves)
        root = levels[-1][0]
        L, S, A = [], [], []
        for i in range(nleaves):
            sib, ap = ok.merkle_open(levels, i)
            L.append(np.frombuffer(leaves[i], dtype=np.uint8))
            S.append(np.frombuffer(sib, dtype=np.uint8))
            A.append(np.frombuffer(b"".join(ap), dtype=np.uint8).reshape(height - 1, 32))
        roots, okf = ctx.keccak_merkle_paths(np.stack(L), np.stack(S), np.stack(A).reshape(nleaves, height - 1, 32),
                                             np.arange(nleaves, dtype=np.uint64), expected_root=root)
        assert all(bytes(r) == root for r in roots) and okf.all()


def test_merkle_paths_prefix_decoded(ctx):
    """Paths arriving in the reference's wire form (prefix-compressed, root-first; mt.go:267-281)."""
    height, leaf_len = 6, 64
    leaves = [os.urandom(leaf_len) for _ in range(1 << height)]
    levels = ok.build_merkle_tree(leaves)
    idxs = [5, 7, 20, 21, 63]
    full = []                                   # root-first full paths (levels 1..h-1)
    for i in idxs:
        _, ap = ok.merkle_open(levels, i)
        full.append(list(reversed(ap)))
    suffixes, plens = [full[0]], [0]
    for j in range(1, len(idxs)):
        p = 0
        while p < len(full[j]) and full[j][p] == full[j - 1][p]:
            p += 1
        suffixes.append(full[j][p:])
        plens.append(p)
    decoded = ok.prefix_decode_paths(suffixes, plens)
    A = np.stack([np.frombuffer(b"".join(p), dtype=np.uint8).reshape(height - 1, 32) for p in decoded])
    S = np.stack([np.frombuffer(ok.merkle_open(levels, i)[0], dtype=np.uint8) for i in idxs])
    Lv = np.stack([np.frombuffer(leaves[i], dtype=np.uint8) for i in idxs])
    roots, okf = ctx.keccak_merkle_paths(Lv, S, A, np.array(idxs, dtype=np.uint64), expected_root=levels[-1][0])
    assert okf.all()


def test_external_known_answer_keccak256(ctx):
    """Ethereum's Keccak-256 of "" and "abc" (public vectors) through k_keccak_f_batch: one padded
    rate block XORed into the zero state, one permutation, first 32 bytes."""
    exp = {b"": "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470",
           b"abc": "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"}
    states = np.zeros((len(exp), 200), dtype=np.uint8)
    for k, msg in enumerate(exp):
        states[k, :len(msg)] = np.frombuffer(msg, dtype=np.uint8)
        states[k, len(msg)] ^= 0x01
        states[k, 135] ^= 0x80
    out = ctx.keccak_f_batch(states.view(np.uint64)).view(np.uint8).reshape(len(exp), 200)
    for k, msg in enumerate(exp):
        assert bytes(out[k, :32]).hex() == exp[msg]
