"""Developer probe (run under gpurun): NTT / computeH / Keccak device timings. Prints JSON lines."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from gnark_whir_b200 import lib  # noqa: E402


def rand_fr_dev(n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = torch.randint(0, 1 << 62, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    a[:, 3] &= (1 << 60) - 1
    return a


def timed(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best


def main():
    logs = [int(x) for x in sys.argv[1:]] or [18, 20, 22, 24]
    ctx = lib.Context(0)
    for logn in logs:
        n = 1 << logn
        a = rand_fr_dev(n, 1)
        # warm tables
        ctx.ntt_dev(a.data_ptr(), logn, inverse=False, coset=True, decimation=lib.DIF)
        for (inv, coset, dec, name) in [(False, False, lib.DIF, "fwd_dif"), (True, False, lib.DIF, "inv_dif"),
                                        (False, True, lib.DIT, "fwd_coset_dit"), (True, True, lib.DIF, "inv_coset_dif")]:
            ms = timed(lambda: ctx.ntt_dev(a.data_ptr(), logn, inverse=inv, coset=coset, decimation=dec))
            print(json.dumps({"probe": "ntt", "logn": logn, "kind": name, "ms": round(ms, 4),
                              "GBs_64N": round(64 * n / ms / 1e6, 1),
                              "Gmodmul_s": round((n / 2) * logn / ms / 1e6, 2)}), flush=True)
        b, c = rand_fr_dev(n, 2), rand_fr_dev(n, 3)
        ms = timed(lambda: ctx.compute_h_dev(a.data_ptr(), b.data_ptr(), c.data_ptr(), logn))
        print(json.dumps({"probe": "compute_h", "logn": logn, "ms": round(ms, 4), "GBs_576N": round(576 * n / ms / 1e6, 1),
                          "phases_ms[3xinv,3xcoset,pointwise,inv_coset]": [round(x, 4) for x in ctx.last_timings()]}),
              flush=True)
        del a, b, c
    for logn in (20, 24):
        n = 1 << logn
        st = torch.randint(0, 1 << 62, (n, 25), dtype=torch.int64, device="cuda")
        ms = timed(lambda: ctx.keccak_f_batch_dev(st.data_ptr(), n))
        print(json.dumps({"probe": "keccak_f_batch", "logn": logn, "ms": round(ms, 4),
                          "Gperm_s": round(n / ms / 1e6, 3), "GBs_400": round(400 * n / ms / 1e6, 1)}), flush=True)
        del st
    # Merkle paths: 2^20 independent paths, height 20, 512-byte leaves (config 4 throughput point)
    n, height, leaf_len = 1 << 20, 20, 512
    leaves = torch.randint(0, 255, (n, leaf_len), dtype=torch.uint8, device="cuda")
    sib = torch.randint(0, 255, (n, 32), dtype=torch.uint8, device="cuda")
    auth = torch.randint(0, 255, (n, height - 1, 32), dtype=torch.uint8, device="cuda")
    idx = torch.randint(0, 1 << 20, (n,), dtype=torch.int64, device="cuda")
    roots = torch.zeros((n, 32), dtype=torch.uint8, device="cuda")
    from ctypes import c_void_p
    L = lib.load()

    def run():
        lib._check(L.b200g16_keccak_merkle_paths_dev(ctx.h, c_void_p(leaves.data_ptr()), leaf_len, c_void_p(sib.data_ptr()),
                                                     c_void_p(auth.data_ptr()), c_void_p(idx.data_ptr()), height, n,
                                                     None, c_void_p(roots.data_ptr()), None))
    ms = timed(run)
    print(json.dumps({"probe": "merkle_paths", "n": n, "height": height, "leaf_len": leaf_len, "ms": round(ms, 4),
                      "Mpaths_s": round(n / ms / 1e3, 2), "Gperm_s": round(24 * n / ms / 1e6, 3)}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
