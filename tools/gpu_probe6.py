"""Developer probe: G2 MSM timing (run under gpurun)."""
import json, sys
import numpy as np, torch
sys.path.insert(0, ".")
from gnark_whir_b200 import lib, groth16 as g16
ctx = lib.Context(0)
rs = np.random.Generator(np.random.PCG64(5))
def rand_fr(n):
    a = rs.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64); a[:, 3] &= np.uint64((1 << 60) - 1); return a
for logn in [int(x) for x in sys.argv[1:]] or [20, 22]:
    n = 1 << logn
    bases = ctx.fixed_base_mul(g16.g2_point(g16.G2_GEN), rand_fr(n), group=2, resident=True)
    sc = torch.from_numpy(rand_fr(n).view(np.int64)).cuda()
    best = None
    for _ in range(3):
        ctx.msm(bases, sc.data_ptr(), n=n); ph = ctx.last_timings()
        if best is None or sum(ph) < sum(best): best = ph
    print(json.dumps({"G2 logn": logn, "dev_ms": round(sum(best), 3), "phases": [round(x, 3) for x in best], "Mpts_s": round(n / sum(best) / 1e3, 1)}), flush=True)
    bases.free()
ctx.close()
