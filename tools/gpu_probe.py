"""Developer probe (run under gpurun): integer-pipe peak and MSM phase timings.
Not part of the product or of the test-suite; prints JSON lines to stdout."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from gnark_whir_b200 import lib  # noqa: E402


def rand_fr(rs, n):
    a = rs.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + rs.integers(0, 2, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    return a


def main():
    logs = [int(x) for x in sys.argv[1:]] or [16, 18, 20, 22]
    ctx = lib.Context(0)
    rs = np.random.Generator(np.random.PCG64(20261018))
    for bps in (1, 2, 4, 8):
        for chains in (1, 2, 4):
            rate, ms = ctx.modmul_probe(bps, chains, 2000)
            print(json.dumps({"probe": "modmul", "blocks_per_sm": bps, "chains": chains, "ms": round(ms, 3),
                              "Gmodmul_s": round(rate / 1e9, 2), "TMAD_s": round(rate * 136 / 1e12, 2)}), flush=True)
    gen = np.zeros(8, dtype=np.uint64)
    from oracle import bn254 as bn
    gen = bn.g1_to_array([bn.G1_GEN])[0]
    for logn in logs:
        n = 1 << logn
        t0 = time.time()
        bases = ctx.fixed_base_mul(gen, rand_fr(rs, n), group=1, resident=True)
        t_gen = time.time() - t0
        sc = rand_fr(rs, n)
        for c in ([0] if logn < 20 else [0, 14, 15, 16]):
            ctx.set_msm_window(c)
            best = None
            for _ in range(3):
                t0 = time.time()
                ctx.msm(bases, sc)
                wall = time.time() - t0
                ph = ctx.last_timings()
                if best is None or sum(ph) < sum(best[1]):
                    best = (wall, ph)
            print(json.dumps({"probe": "msm_g1", "logn": logn, "c": c, "gen_s": round(t_gen, 3),
                              "wall_ms": round(best[0] * 1e3, 3), "dev_ms": round(sum(best[1]), 3),
                              "phases_ms[digits,sort,accumulate,merge,reduce]": [round(x, 3) for x in best[1]],
                              "Mpts_s_dev": round(n / sum(best[1]) / 1e3, 2)}), flush=True)
        ctx.set_msm_window(0)
        bases.free()
    ctx.close()


if __name__ == "__main__":
    main()
