"""Developer probe: NTT / computeH timings (run under gpurun)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import torch
from gnark_whir_b200 import lib
ctx = lib.Context(0)
def rnd(n):
    a = torch.randint(0, 1 << 62, (n, 4), dtype=torch.int64, device="cuda"); a[:, 3] &= (1 << 60) - 1; return a
def t_ms(fn, reps=5):
    fn(); best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
for L in (16, 18, 20, 22, 23, 24, 25, 26):
    a, b, c = rnd(1 << L), rnd(1 << L), rnd(1 << L)
    ctx.ntt_dev(a.data_ptr(), L, coset=True, decimation=lib.DIF)
    f = t_ms(lambda: ctx.ntt_dev(a.data_ptr(), L, decimation=lib.DIF))
    i = t_ms(lambda: ctx.ntt_dev(a.data_ptr(), L, inverse=True, coset=True, decimation=lib.DIT))
    h = t_ms(lambda: ctx.compute_h_dev(a.data_ptr(), b.data_ptr(), c.data_ptr(), L))
    print(f"L={L} ntt_dif {f:.4f} ms  intt_coset_dit {i:.4f} ms  compute_h {h:.4f} ms  ({(1<<L)/2*L*136/f/1e9:.3f} TMAD/s)")
ctx.close()
