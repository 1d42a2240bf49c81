import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from gnark_whir_b200 import lib, groth16 as g16
ctx = lib.Context(0)
rs = np.random.Generator(np.random.PCG64(5))
def rand_fr(n):
    a = rs.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64); a[:, 3] &= np.uint64((1 << 60) - 1); return a
n = 1 << 18
bases = ctx.fixed_base_mul(g16.g1_point(g16.G1_GEN), rand_fr(n), group=1, resident=True)
sc = rand_fr(n)
pin = torch.from_numpy(sc.view(np.int64)).pin_memory()
pin_np = pin.numpy().view(np.uint64)
dev = pin.cuda()
def t(fn, reps=20):
    fn(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps * 1e3
for c in (0, 13, 15, 17):
    ctx.set_msm_window(c)
    print("c", c, "plan", ctx.msm_plan(bases, n), "pageable %.3f  pinned %.3f  device-ptr %.3f ms" % (
        t(lambda: ctx.msm(bases, sc)), t(lambda: ctx.msm(bases, pin_np)), t(lambda: ctx.msm(bases, dev.data_ptr(), n=n))),
        "device phases", [round(x, 3) for x in ctx.last_timings()])
ctx.set_msm_window(0)
ctx.close()
