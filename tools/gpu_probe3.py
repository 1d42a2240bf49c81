"""Developer probe: instruction-level integer pipe ceilings (run under gpurun)."""
import json, sys
sys.path.insert(0, ".")
from gnark_whir_b200 import lib
ctx = lib.Context(0)
names = {0: "IMAD", 1: "IMAD.WIDE(+IADD3 pair)", 2: "IMAD.WIDE.X chain", 3: "IMAD.HI", 4: "IADD3.X chain", 5: "WIDE.X + IADD3 interleaved"}
for mode in range(6):
    for bps in (2, 4, 8):
        rate, ms = ctx.pipe_probe(mode, bps, 4000)
        print(json.dumps({"mode": mode, "name": names[mode], "blocks_per_sm": bps, "ms": round(ms, 3),
                          "Tops_s": round(rate / 1e12, 3), "per_clk_per_sm@1.965GHz": round(rate / 148 / 1.965e9, 2)}), flush=True)
for bps, ch in ((4, 1), (8, 4), (12, 2), (16, 1)):
    rate, ms = ctx.modmul_probe(bps, ch, 2000)
    print(json.dumps({"modmul": True, "bps": bps, "chains": ch, "Gmodmul_s": round(rate / 1e9, 2)}), flush=True)
ctx.close()
