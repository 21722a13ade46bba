"""Key metrics of every kernel in an .ncu-rep (from `ncu --set full`), as markdown.
Usage: ncu_keymetrics.py report.ncu-rep"""
import csv, subprocess, sys, io
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(h, r)); u = dict(zip(h, units))
    print(f"#### `{d['Kernel Name'][:110]}`  grid {d.get('Grid Size')} block {d.get('Block Size')}\n")
    print("| metric | value | unit |\n|---|---:|---|")
    for k in KEYS:
        if k in d and d[k] != "":
            print(f"| {k} | {d[k]} | {u[k]} |")
    print()
