#!/usr/bin/env python
"""Selected columns of `ncu -i X.ncu-rep --page raw --csv` (stdin), one line per captured launch."""
import csv
import sys

WANT = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "us"),
        ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("smsp__inst_executed.sum", "inst"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "mem%"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("l1tex__t_sector_hit_rate.pct", "l1hit"), ("lts__t_sector_hit_rate.pct", "l2hit"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st_noinst"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "st_branch"),
        ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "st_disp"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_notsel"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64c%")]
rows = list(csv.reader(sys.stdin))
h = rows[0]
cols = [(h.index(k), n) for k, n in WANT if k in h]
w = csv.writer(sys.stdout)
w.writerow([n for _, n in cols])
for r in rows[2:]:
    out = []
    for i, n in cols:
        v = r[i]
        try:
            v = f"{float(v):.2f}" if n not in ("kernel", "grid") else v
        except ValueError:
            pass
        out.append(v[:34] if n == "kernel" else v)
    w.writerow(out)
