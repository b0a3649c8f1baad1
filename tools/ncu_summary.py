"""Summarise an .ncu-rep (raw page) into the handful of counters DESIGN.md and
profiles/ quote.  usage: python tools/ncu_summary.py <file.ncu-rep> [pattern ...]"""
import csv
import io
import subprocess
import sys

DEFAULT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__occupancy_limit", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "smsp__cycles_active.avg", "pipe_alu", "pipe_fma", "pipe_lsu", "pipe_xu", "pipe_uniform",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__average_warp", "smsp__average_warps_issue_stalled",
    "issue_stalled", "smsp__thread_inst_executed_per_inst_executed", "smsp__inst_executed.avg.per_cycle_active",
    "sm__inst_executed.avg.per_cycle_elapsed", "l1tex__data_bank_conflicts", "clock_rate", "sm__maximum_warps",
]


def main():
    rep = sys.argv[1]
    pats = sys.argv[2:] or DEFAULT
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    for r in data:
        print("==", r[ki], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    for i, h in enumerate(hdr):
        if any(p in h for p in pats):
            vals = [r[i] for r in data]
            print(f"{h} [{units[i]}] = {vals}")


if __name__ == "__main__":
    main()
