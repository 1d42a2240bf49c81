"""Per-instruction stall samples of one kernel from an `ncu --set full --import-source on` report, condensed:
   python tools/ncu_hotspots.py report.ncu-rep [top] > profiles/x.txt
Prints the stall-reason totals, a coarse code map (samples per 64 instructions) and the `top` instructions by samples."""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    print("#", rows[0][1][:150])
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]

    def num(r, c):
        try:
            return int(r[ix[c]] or 0)
        except (ValueError, KeyError):
            return 0
    cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = {c: sum(num(r, c) for r in data) for c in cols}
    s = sum(tot.values()) or 1
    print("stall samples:", "  ".join(f"{c[6:]}={100 * v / s:.1f}%" for c, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v * 200 > s))
    base = int(data[0][ix["Address"]], 16)
    print("code map (offset: median executions, samples, long_sb / wait / math / barrier / short_sb / mio, #IMAD.WIDE, #LDS+STS, #LDG+STG):")
    for k in range(0, len(data), 64):
        blk = data[k:k + 64]
        ex = sorted(num(r, "Instructions Executed") for r in blk)[len(blk) // 2]
        src = [r[ix["Source"]] for r in blk]
        print(f"  {int(blk[0][ix['Address']], 16) - base:#7x}: {ex:9d} {sum(num(r, '# Samples') for r in blk):7d}  "
              f"{sum(num(r, 'stall_long_sb') for r in blk):6d}/{sum(num(r, 'stall_wait') for r in blk):6d}/"
              f"{sum(num(r, 'stall_math') for r in blk):6d}/{sum(num(r, 'stall_barrier') for r in blk):6d}/"
              f"{sum(num(r, 'stall_short_sb') for r in blk):6d}/{sum(num(r, 'stall_mio') for r in blk):6d}  "
              f"{sum('IMAD.WIDE' in x for x in src):3d} {sum(('LDS' in x) or ('STS' in x) for x in src):3d} "
              f"{sum(('LDG' in x) or ('STG' in x) or ('LD.E' in x) or ('ST.E' in x) for x in src):3d}")
    print(f"top {top} instructions by samples (offset, samples, long_sb, executions, SASS):")
    for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:top]:
        print(f"  {int(r[ix['Address']], 16) - base:#7x} {num(r, '# Samples'):7d} {num(r, 'stall_long_sb'):7d} "
              f"{num(r, 'Instructions Executed'):9d}  {r[ix['Source']].strip()[:100]}")


if __name__ == "__main__":
    main()
