"""Developer probe: window-table MSM, table width sweep with per-phase timings (run under gpurun)."""
import json, sys
import numpy as np
sys.path.insert(0, ".")
from gnark_whir_b200 import lib, groth16 as g16
import torch
ctx = lib.Context(0)
rs = np.random.Generator(np.random.PCG64(5))
def rand_fr(n):
    a = rs.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64); a[:, 3] &= np.uint64((1 << 60) - 1); return a
group = 1
args = [a for a in sys.argv[1:]]
if args and args[0] == "g2":
    group = 2; args = args[1:]
for spec in args or ["24:18,20,21,22"]:
    logn, cs = spec.split(":")
    logn = int(logn); n = 1 << logn
    ks, scn = rand_fr(n), rand_fr(n)
    sc = torch.from_numpy(scn.view(np.int64)).cuda()
    gen = g16.g1_point(g16.G1_GEN) if group == 1 else g16.g2_point(g16.G2_GEN)
    ref = None
    for c in [int(x) for x in cs.split(",")]:
        bases = ctx.fixed_base_mul(gen, ks, group=group, resident=True)
        if c >= 0:
            bases.precompute(c)
        cw = ctx.msm_plan(bases, n)
        best = None
        for _ in range(3):
            out = ctx.msm(bases, sc.data_ptr(), n=n)
            ph = ctx.last_timings()
            if best is None or sum(ph) < sum(best): best = ph
        bases.free()
        if ref is None: ref = out
        assert np.array_equal(out, ref), c
        print(json.dumps({"group": group, "logn": logn, "table_c": c, "plan": cw, "dev_ms": round(sum(best), 3),
                          "phases[digits,sort,acc,merge,reduce]": [round(x, 3) for x in best],
                          "Mpts_s": round(n / sum(best) / 1e3, 1)}), flush=True)
ctx.close()
