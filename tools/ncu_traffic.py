"""DRAM traffic of one captured kernel launch -> the JSON artefact bench.py reads (profiles/*ncu_traffic*.json):
   python tools/ncu_traffic.py report.ncu-rep 'k_accumulate<Fp>' <log2n> <window_table 0|1> > profiles/rNN_ncu_traffic_....json"""
import csv
import json
import subprocess
import sys


def main():
    rep, kernel, logn, table = sys.argv[1], sys.argv[2], int(sys.argv[3]), bool(int(sys.argv[4]))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u, r = rows[0], rows[1], rows[2]
    d = {h[i]: (r[i], u[i]) for i in range(len(h))}

    def to_bytes(key):
        v, unit = d[key]
        return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
    rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    print(json.dumps({"kernel": kernel, "log2n": logn, "window_table": table, "dram_bytes_per_launch": rd + wr,
                      "dram_bytes_read": rd, "dram_bytes_write": wr, "gpu_time_ms": float(d["gpu__time_duration.sum"][0]),
                      "grid": int(d["launch__grid_size"][0]), "commit": commit, "report": rep,
                      "how": "ncu --set full --clock-control none, one launch; dram__bytes_read.sum + dram__bytes_write.sum"}, indent=1))


if __name__ == "__main__":
    main()
