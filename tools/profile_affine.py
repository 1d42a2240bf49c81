"""One G1 MSM with the batched-affine accumulation forced (for ncu): python tools/profile_affine.py [logn] [levels]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from gnark_whir_b200 import groth16 as g16  # noqa: E402
from gnark_whir_b200 import lib  # noqa: E402

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 22
levels = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rs = np.random.Generator(np.random.PCG64(7))


def rand_fr(n):
    a = rs.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    return a


ctx = lib.Context(0)
n = 1 << logn
bases = ctx.fixed_base_mul(g16.g1_point(g16.G1_GEN), rand_fr(n), group=1, resident=True)
bases.precompute(0)
sc = torch.from_numpy(rand_fr(n).view(np.int64)).cuda()
ctx.set_msm_batch_affine(2, levels, 0)
for _ in range(2):
    ctx.msm(bases, sc.data_ptr(), n=n)
    print([round(x, 3) for x in ctx.last_timings()])
ctx.close()
