"""NTT / computeH timing at one size (A/B of kernel variants selected by environment variables)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from gnark_whir_b200 import lib  # noqa: E402

rs = np.random.Generator(np.random.PCG64(3))


def rand_fr(n):
    a = rs.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    return a


def t_ms(fn, reps=5):
    fn()
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


ctx = lib.Context(0)
for logn in (20, 22, 24):
    n = 1 << logn
    a, b, c = (torch.from_numpy(rand_fr(n).view(np.int64)).cuda() for _ in range(3))
    f = t_ms(lambda: ctx.ntt_dev(a.data_ptr(), logn, decimation=lib.DIF))
    i = t_ms(lambda: ctx.ntt_dev(a.data_ptr(), logn, inverse=True, decimation=lib.DIT))
    h = t_ms(lambda: ctx.compute_h_dev(a.data_ptr(), b.data_ptr(), c.data_ptr(), logn))
    print(f"log2n={logn} fwd_dif={f:.4f} ms inv_dit={i:.4f} ms compute_h={h:.4f} ms")
ctx.close()
