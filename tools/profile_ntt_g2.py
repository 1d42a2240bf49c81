"""Fr NTT 2^24 (3 passes) and one G2 MSM 2^20 with a window table, for `ncu --import-source on` captures."""
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


what = sys.argv[1] if len(sys.argv) > 1 else "ntt"
ctx = lib.Context(0)
if what == "ntt":
    a = torch.from_numpy(rand_fr(1 << 24).view(np.int64)).cuda()
    for _ in range(3):
        ctx.ntt_dev(a.data_ptr(), 24, decimation=lib.DIF)
else:
    m = 1 << 20
    b2 = ctx.fixed_base_mul(g16.g2_point(g16.G2_GEN), rand_fr(m), group=2, resident=True)
    b2.precompute(0)
    sc = torch.from_numpy(rand_fr(m).view(np.int64)).cuda()
    for _ in range(2):
        ctx.msm(b2, sc.data_ptr(), n=m)
        print([round(x, 3) for x in ctx.last_timings()])
ctx.close()
print("ok")
