"""Developer probe: Keccak-f batch timing (variant chosen by B200G16_UNUSED), checked against the C oracle."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
import torch
from gnark_whir_b200 import lib
from oracle import cport
ctx = lib.Context(0)
nk = 1 << 22
st = torch.randint(-(1 << 62), 1 << 62, (nk, 25), dtype=torch.int64, device="cuda")
sample = st[:64].cpu().numpy().view(np.uint64).copy()
ctx.keccak_f_batch_dev(st.data_ptr(), nk)
got = st[:64].cpu().numpy().view(np.uint64)
exp = cport.keccak_f_batch(sample)
ok = bool(np.array_equal(got, exp))
best = 1e9
for _ in range(5):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.keccak_f_batch_dev(st.data_ptr(), nk); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(f"variant={os.environ.get('B200G16_UNUSED','0')} ok={ok} ms={best:.4f} Gperm/s={nk/best/1e6:.3f}")
