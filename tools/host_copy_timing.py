"""Host-pointer entry points on pageable buffers (csrc/host_copy.cuh): wall-clock of b200g16_ntt at 2^24
(512 MiB up, 512 MiB back) and of b200g16_keccak_f_batch on 2^22 states (800 MiB each way)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from gnark_whir_b200 import lib  # noqa: E402

ctx = lib.Context(0)
rs = np.random.Generator(np.random.PCG64(1))
a = rs.integers(0, 1 << 62, size=(1 << 24, 4), dtype=np.uint64)
a[:, 3] &= np.uint64((1 << 60) - 1)
L = lib.load()
buf = a.copy()
for _ in range(3):
    t = time.time()
    lib._check(L.b200g16_ntt(ctx.h, lib._ptr(buf), 24, 0, 0, lib.DIF))
    print("b200g16_ntt 2^24 from / to pageable memory:", round((time.time() - t) * 1e3, 1), "ms (transform itself 3.5 ms)")
st = rs.integers(0, 1 << 63, size=(1 << 22, 25), dtype=np.uint64)
for _ in range(3):
    t = time.time()
    lib._check(L.b200g16_keccak_f_batch(ctx.h, lib._ptr(st), st.shape[0]))
    print("b200g16_keccak_f_batch 2^22 states from / to pageable memory:", round((time.time() - t) * 1e3, 1), "ms (kernel 1.4 ms)")
ctx.close()
