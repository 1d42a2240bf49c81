"""Compact per-kernel table from an `ncu --set full` report:  python tools/ncu_summary.py report.ncu-rep > profiles/x.txt"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe_alu%"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "pipe_fma%"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "pipe_fmaheavy%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("launch__registers_per_thread", "regs"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_scoreboard"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_scoreboard"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio_throttle"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_selected"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg_throttle"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_no_instruction"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall_branch_resolving"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall_dispatch"),
    ("smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio", "stall_imc_miss"),
    ("smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio", "stall_tex_throttle"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active_threads_per_inst"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit%"), ("lts__t_sector_hit_rate.pct", "l2_hit%"),
]


def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {sys.argv[1]}: ncu --set full --clock-control none, one line per captured launch")
    for r in rows[2:]:
        parts = [r[idx["Kernel Name"]].split("(")[0]]
        for k, short in KEYS:
            if k in idx:
                v = r[idx[k]]
                try:
                    v = f"{float(v.replace(',', '')):.4g}"
                except ValueError:
                    pass
                parts.append(f"{short}={v}{units[idx[k]] if short in ('time', 'dram_rd', 'dram_wr') else ''}")
        print("  ".join(parts))


if __name__ == "__main__":
    main()
