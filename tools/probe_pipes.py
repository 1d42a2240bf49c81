"""Instruction-pipe and multiplier probes (run under gpurun): one JSON line per measurement -> stdout.

  integer pipes   IMAD, IMAD.WIDE, IMAD.WIDE.X carry chains, IMAD.HI, IADD3.X, wide + IADD3 interleaved
  FP64 pipe       DFMA, DFMA + three-input 64-bit adds, DFMA + IMAD.WIDE.X 1:1 (co-issue), DADD
  multipliers     field.cuh's 8x32-bit integer Montgomery product vs fp52.cuh's 5x52-bit DFMA product, alone and
                  side by side in one thread (b200g16_fp52_probe)
"""
import json
import sys

sys.path.insert(0, ".")
from gnark_whir_b200 import lib  # noqa: E402

NAMES = {0: "IMAD", 1: "IMAD.WIDE(+IADD3 pair)", 2: "IMAD.WIDE.X chain", 3: "IMAD.HI", 4: "IADD3.X chain",
         5: "WIDE.X + IADD3 interleaved (wide rate)", 6: "DFMA.RZ", 7: "DFMA + 3-input IADD64 interleaved (DFMA rate)",
         8: "DFMA + IMAD.WIDE.X 1:1 (rate of each)", 9: "DADD"}
VARIANTS = {0: "fp52 x1", 1: "fp52 x2", 2: "fp52 x4", 3: "int x1 + fp52 x1", 4: "int x1 + fp52 x2", 5: "int x2"}


def main():
    ctx = lib.Context(0)
    clk = 1.965e9 * 148
    for mode in range(10):
        for bps in (4, 8):
            rate, ms = ctx.pipe_probe(mode, bps, 2000)
            print(json.dumps({"mode": mode, "name": NAMES[mode], "blocks_per_sm": bps, "ms": round(ms, 3),
                              "Tops_s": round(rate / 1e12, 3), "per_clk_per_sm@1.965GHz": round(rate / clk, 2)}), flush=True)
    for bps, chains in ((8, 4), (12, 2)):
        rate, _ = ctx.modmul_probe(bps, chains, 2000)
        print(json.dumps({"modmul": "integer (field.cuh)", "bps": bps, "chains": chains, "Gmodmul_s": round(rate / 1e9, 2)}), flush=True)
    for v in range(6):
        for bps in (2, 4, 6, 8):
            rate, ms = ctx.fp52_probe(v, bps, 1000)
            print(json.dumps({"fp52_probe": VARIANTS[v], "blocks_per_sm": bps, "ms": round(ms, 3),
                              "Gmodmul_s": round(rate / 1e9, 2)}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
