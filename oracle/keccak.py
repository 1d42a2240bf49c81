"""Keccak-f[1600], the reference's overwrite-mode duplex sponge, and Merkle-path recompute.

Oracle = test infrastructure (see oracle/__init__.py).
  keccak_f            <- gnark v0.11.0 std/permutation/keccakf.Permute, called at
                         keccakSponge/keccakSponge.go:48,69 (standard FIPS-202 permutation,
                         25 little-endian u64 lanes, lane index = x + 5y)
  Sponge              <- keccakSponge/keccakSponge.go:9-75 (Digest / NewKeccak /
                         NewKeccakWithTag / Absorb / Squeeze): rate 136 B, OVERWRITE absorb,
                         no padding, no domain byte
  merkle_root_from_path <- mtUtilities.go:109-141 VerifyMerkleTreeProofs (direction rule:
                         index bit `level` set => current node is the RIGHT child; level 0
                         pairs with LeafSiblingHash, level k>=1 with AuthPaths[k-1]), with
                         the 2-to-1 hash instantiated by the Keccak duplex as BASELINE.json
                         config 4 asks (the reference snapshot instantiates it with
                         Skyscraper)
  prefix_decode_paths <- mt.go:267-281 + utilities/utilities.go:58-78
"""
from __future__ import annotations

RATE = 136
_M = (1 << 64) - 1

RC = []
ROT = [[0] * 5 for _ in range(5)]


def _init():
    # round constants via the LFSR of FIPS-202 3.2.5; rotation offsets via 3.2.2
    def rc_bit(t):
        t %= 255
        if t == 0:
            return 1
        r = 1
        for _ in range(t):
            r <<= 1
            if r & 0x100:
                r ^= 0x171
        return r & 1
    for i in range(24):
        c = 0
        for j in range(7):
            if rc_bit(j + 7 * i):
                c |= 1 << ((1 << j) - 1)
        RC.append(c)
    x, y = 1, 0
    for t in range(24):
        ROT[x][y] = ((t + 1) * (t + 2) // 2) % 64
        x, y = y, (2 * x + 3 * y) % 5


_init()


def _rol(v, n):
    n %= 64
    return ((v << n) | (v >> (64 - n))) & _M if n else v


def keccak_f(state):
    """state: list of 25 ints (lane x+5y). Returns new list."""
    a = list(state)
    for rnd in range(24):
        c = [a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [a[i] ^ d[i % 5] for i in range(25)]
        b = [0] * 25
        for x in range(5):
            for y in range(5):
                b[y + 5 * ((2 * x + 3 * y) % 5)] = _rol(a[x + 5 * y], ROT[x][y])
        a = [b[i] ^ ((~b[(i % 5 + 1) % 5 + 5 * (i // 5)]) & _M & b[(i % 5 + 2) % 5 + 5 * (i // 5)])
             for i in range(25)]
        a[0] ^= RC[rnd]
    return a


def state_to_bytes(st):
    return b"".join(int(v).to_bytes(8, "little") for v in st)


def bytes_to_state(b):
    return [int.from_bytes(b[8 * i:8 * i + 8], "little") for i in range(25)]


def sha3_like(msg, rate, suffix, outlen):
    """FIPS-202 sponge (XOR absorb + pad10*1) on top of keccak_f — exists only to pin keccak_f
    against hashlib."""
    st = bytearray(200)
    msg = bytearray(msg) + bytes([suffix])
    while len(msg) % rate:
        msg.append(0)
    msg[-1] |= 0x80
    for off in range(0, len(msg), rate):
        for i in range(rate):
            st[i] ^= msg[off + i]
        st = bytearray(state_to_bytes(keccak_f(bytes_to_state(st))))
    out = b""
    while len(out) < outlen:
        out += bytes(st[:rate])
        if len(out) < outlen:
            st = bytearray(state_to_bytes(keccak_f(bytes_to_state(st))))
    return out[:outlen]


class Sponge:
    """keccakSponge.Digest (keccakSponge.go:9-75)."""

    def __init__(self, tag=b""):
        self.state = bytearray(200)
        self.absorb_pos = 0
        self.squeeze_pos = RATE
        self.n_permutes = 0
        for i, t in enumerate(tag):               # NewKeccakWithTag, :31-38
            self.state[RATE + i] = t

    def _permute(self):
        self.state = bytearray(state_to_bytes(keccak_f(bytes_to_state(self.state))))
        self.n_permutes += 1

    def absorb(self, data):                       # :40-56
        for byte in data:
            if self.absorb_pos == RATE:
                self._permute()
                self.absorb_pos = 0
            self.state[self.absorb_pos] = byte
            self.absorb_pos += 1
        self.squeeze_pos = RATE

    def squeeze(self, n):                         # :64-75
        out = bytearray()
        for _ in range(n):
            if self.squeeze_pos == RATE:
                self.squeeze_pos = 0
                self.absorb_pos = 0
                self._permute()
            out.append(self.state[self.squeeze_pos])
            self.squeeze_pos += 1
        return bytes(out)


def sponge_hash(data, outlen=32):
    s = Sponge()
    s.absorb(data)
    return s.squeeze(outlen)


def merkle_leaf_hash(leaf_bytes):
    return sponge_hash(leaf_bytes, 32)


def merkle_node_hash(left, right):
    return sponge_hash(left + right, 32)


def merkle_root_from_path(leaf_bytes, leaf_sibling, auth_path, index):
    """auth_path[k-1] is the sibling at level k (k = 1..height-1), leaf_sibling at level 0;
    tree height = len(auth_path)+1 (mtUtilities.go:113)."""
    cur = merkle_leaf_hash(leaf_bytes)
    sib = leaf_sibling
    height = len(auth_path) + 1
    for level in range(height):
        if level > 0:
            sib = auth_path[level - 1]
        if (index >> level) & 1:
            cur = merkle_node_hash(sib, cur)
        else:
            cur = merkle_node_hash(cur, sib)
    return cur


def build_merkle_tree(leaves_bytes):
    """Full tree for test generation: returns list of levels, levels[0] = leaf hashes."""
    level = [merkle_leaf_hash(l) for l in leaves_bytes]
    levels = [level]
    while len(level) > 1:
        level = [merkle_node_hash(level[i], level[i + 1]) for i in range(0, len(level), 2)]
        levels.append(level)
    return levels


def merkle_open(levels, index):
    """-> (leaf_sibling, auth_path) in the layout merkle_root_from_path expects."""
    sibs = []
    for lv in levels[:-1]:
        sibs.append(lv[index ^ 1])
        index >>= 1
    return sibs[0], sibs[1:]


def prefix_decode_paths(suffixes, prefix_lens):
    """mt.go:267-281: suffixes[j] is root-first; path j = prev[:prefix_lens[j]] + suffixes[j];
    the circuit consumes each path reversed (leaf-side first)."""
    out = []
    prev = list(suffixes[0])
    out.append(list(reversed(prev)))
    for j in range(1, len(suffixes)):
        if prefix_lens[j] == 0:
            prev = list(suffixes[j])
        else:
            prev = prev[:prefix_lens[j]] + list(suffixes[j])
        out.append(list(reversed(prev)))
    return out
