"""Developer probe: MSM window sweep at large n (run under gpurun)."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
from gnark_whir_b200 import lib, groth16 as g16
ctx = lib.Context(0)
rs = np.random.Generator(np.random.PCG64(5))
def rand_fr(n):
    a = rs.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64); a[:, 3] &= np.uint64((1 << 60) - 1); return a
import torch
for logn in [int(x) for x in sys.argv[1:]] or [22, 24]:
    n = 1 << logn
    bases = ctx.fixed_base_mul(g16.g1_point(g16.G1_GEN), rand_fr(n), group=1, resident=True)
    sc = torch.from_numpy(rand_fr(n).view(np.int64)).cuda()
    ref = None
    for c in (0, 16, 17, 18, 19, 20):
        ctx.set_msm_window(c)
        best = None
        for _ in range(3):
            out = ctx.msm(bases, sc.data_ptr(), n=n)
            ph = ctx.last_timings()
            if best is None or sum(ph) < sum(best): best = ph
        if ref is None: ref = out
        assert np.array_equal(out, ref), c
        print(json.dumps({"logn": logn, "c": c, "dev_ms": round(sum(best), 3), "phases": [round(x, 3) for x in best],
                          "Mpts_s": round(n / sum(best) / 1e3, 1)}), flush=True)
    ctx.set_msm_window(0)
    bases.free()
ctx.close()
