"""One short pass over every kernel family, for `ncu --set full` (profiles/r02*_ncu_full_summary.txt):
G1 MSM 2^22 with a window table (digits, scatter, accumulate, tree reduction), G2 MSM 2^20, Fr NTT 2^24 and
computeH 2^22, Keccak-f batch 2^22, Merkle paths (256 = warp-per-path kernel, 2^18 = thread-per-path kernel)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from gnark_whir_b200 import groth16 as g16  # noqa: E402
from gnark_whir_b200 import lib  # noqa: E402

rs = np.random.Generator(np.random.PCG64(5))


def rand_fr(n):
    a = rs.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    return a


def dev(a):
    return torch.from_numpy(a.view(np.int64)).cuda()


ctx = lib.Context(0)
n = 1 << 22
b1 = ctx.fixed_base_mul(g16.g1_point(g16.G1_GEN), rand_fr(n), group=1, resident=True)
b1.precompute(0)
sc = dev(rand_fr(n))
for _ in range(2):
    ctx.msm(b1, sc.data_ptr(), n=n)
b1.free()
m = 1 << 20
b2 = ctx.fixed_base_mul(g16.g2_point(g16.G2_GEN), rand_fr(m), group=2, resident=True)
b2.precompute(0)
for _ in range(2):
    ctx.msm(b2, sc.data_ptr(), n=m)
b2.free()
a, b, c = dev(rand_fr(1 << 24)), dev(rand_fr(1 << 22)), dev(rand_fr(1 << 22))
for _ in range(2):
    ctx.ntt_dev(a.data_ptr(), 24, decimation=lib.DIF)
ctx.compute_h_dev(a.data_ptr(), b.data_ptr(), c.data_ptr(), 22)
ctx.compute_h_dev(a.data_ptr(), b.data_ptr(), c.data_ptr(), 22)
st = torch.randint(0, 1 << 62, (1 << 22, 25), dtype=torch.int64, device="cuda")
ctx.keccak_f_batch_dev(st.data_ptr(), 1 << 22)
ctx.keccak_f_batch_dev(st.data_ptr(), 1 << 22)
for q in (256, 1 << 18):
    leaves = rs.integers(0, 256, size=(q, 512), dtype=np.uint8)
    sib = rs.integers(0, 256, size=(q, 32), dtype=np.uint8)
    auth = rs.integers(0, 256, size=(q, 19, 32), dtype=np.uint8)
    idx = rs.integers(0, 1 << 20, size=q, dtype=np.uint64)
    for _ in range(2):
        ctx.keccak_merkle_paths(leaves, sib, auth, idx)
ctx.close()
print("profile_families ok")
