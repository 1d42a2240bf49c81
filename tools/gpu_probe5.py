"""Developer probe: MSM phase timings for the WHIR-shaped scalar mix vs uniform (run under gpurun)."""
import json, sys
import numpy as np, torch
sys.path.insert(0, ".")
from gnark_whir_b200 import lib, groth16 as g16
ctx = lib.Context(0)
rs = np.random.Generator(np.random.PCG64(5))
def rand_fr(n):
    a = rs.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64); a[:, 3] &= np.uint64((1 << 60) - 1); return a
tbl = torch.from_numpy(g16.fr_array(list(range(256))).view(np.int64)).cuda()
for logn in [int(x) for x in sys.argv[1:]] or [22, 24]:
    n = 1 << logn
    bases = ctx.fixed_base_mul(g16.g1_point(g16.G1_GEN), rand_fr(n), group=1, resident=True)
    uni = torch.from_numpy(rand_fr(n).view(np.int64)).cuda()
    u = torch.rand(n, device="cuda"); sv = torch.randint(0, 256, (n,), device="cuda")
    mix = uni.clone(); m01 = u < 0.4; mb = (u >= 0.4) & (u < 0.7)
    mix[m01] = tbl[sv[m01] & 1]; mix[mb] = tbl[sv[mb]]
    ones = tbl[torch.ones(n, dtype=torch.int64, device="cuda")].contiguous()
    for name, sc in (("uniform", uni), ("whir_mix", mix), ("all_ones", ones)):
        best = None
        for _ in range(3):
            ctx.msm(bases, sc.data_ptr(), n=n); ph = ctx.last_timings()
            if best is None or sum(ph) < sum(best): best = ph
        print(json.dumps({"logn": logn, "scalars": name, "dev_ms": round(sum(best), 3),
                          "phases[digits,sort,accumulate,merge,reduce]": [round(x, 3) for x in best]}), flush=True)
    bases.free()
ctx.close()
