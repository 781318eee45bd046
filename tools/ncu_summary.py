"""Filter `ncu -i X.ncu-rep --page raw --csv` down to the columns quoted in DESIGN.md / read by bench.py
(profiles/r01*_ncu_summary.csv).  usage: python tools/ncu_summary.py raw.csv > summary.csv"""
import csv
import sys

COLS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 20]
h = rows[0]
idx = [h.index(c) for c in COLS if c in h]
w = csv.writer(sys.stdout)
for r in rows:
    w.writerow([r[i] for i in idx])
