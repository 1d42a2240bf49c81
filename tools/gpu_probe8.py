"""Developer probe: one pass of each kernel family for an `ncu --set full` capture (run under gpurun)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import torch
from gnark_whir_b200 import lib, groth16 as g16
ctx = lib.Context(0)
rs = np.random.Generator(np.random.PCG64(5))
def rand_fr(n):
    a = rs.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64); a[:, 3] &= np.uint64((1 << 60) - 1); return a
dev = torch.device("cuda", 0)
what = sys.argv[1:] or ["keccak", "ntt", "msm"]
if "keccak" in what:
    nk = 1 << 22
    st = torch.randint(0, 1 << 62, (nk, 25), dtype=torch.int64, device=dev)
    for _ in range(2):
        ctx.keccak_f_batch_dev(st.data_ptr(), nk)
if "ntt" in what:
    L = 24
    a = torch.from_numpy(rand_fr(1 << L).view(np.int64)).to(dev)
    for _ in range(2):
        ctx.ntt_dev(a.data_ptr(), L, decimation=lib.DIF)
if "msm" in what:
    logn = 22
    n = 1 << logn
    bases = ctx.fixed_base_mul(g16.g1_point(g16.G1_GEN), rand_fr(n), group=1, resident=True)
    bases.precompute(0)
    sc = torch.from_numpy(rand_fr(n).view(np.int64)).to(dev)
    for _ in range(2):
        ctx.msm(bases, sc.data_ptr(), n=n)
    bases.free()
ctx.close()
