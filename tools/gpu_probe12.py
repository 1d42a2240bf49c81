"""Developer probe: latency of small host-scalar MSMs (the BSB22 commitment inside the solver), run under gpurun."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from gnark_whir_b200 import lib, groth16 as g16
ctx = lib.Context(0)
rs = np.random.Generator(np.random.PCG64(5))
def rand_fr(n):
    a = rs.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64); a[:, 3] &= np.uint64((1 << 60) - 1); return a
for logn in (6, 8, 10, 12, 14, 16, 18):
    n = 1 << logn
    bases = ctx.fixed_base_mul(g16.g1_point(g16.G1_GEN), rand_fr(n), group=1, resident=True)
    sc = rand_fr(n)
    for tab in (False, True):
        if tab:
            bases.precompute(0)
        ctx.msm(bases, sc)
        t0 = time.perf_counter()
        for _ in range(10):
            ctx.msm(bases, sc)
        wall = (time.perf_counter() - t0) / 10 * 1e3
        ph = ctx.last_timings()
        print(f"n=2^{logn} table={tab} plan={ctx.msm_plan(bases, n)} wall {wall:.3f} ms  device {sum(ph):.3f} ms phases {[round(x,3) for x in ph]}")
    bases.free()
ctx.close()
