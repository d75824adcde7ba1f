"""Condense `ncu -i X.ncu-rep --page raw --csv` into one line per kernel launch (the summaries committed under profiles/)."""
import csv, sys
COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("smsp__inst_executed.sum", "winst"), ("launch__registers_per_thread", "regs"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "st_branch"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st_noinst"),
        ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "st_disp"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst")]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
idx = [(hdr.index(c), n) for c, n in COLS if c in hdr]
w = csv.writer(sys.stdout)
w.writerow(["kernel"] + ["%s[%s]" % (n, units[i]) for i, n in idx])
for r in rows[2:]:
    name = r[ki].split("(")[0].replace("void ", "")
    w.writerow([name] + [r[i] for i, _ in idx])
