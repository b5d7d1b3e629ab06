"""print the key counters of `ncu --page raw --csv` exports: python scripts/ncu_summ.py gpurun_out/prof_x.csv ..."""
import csv, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum"]
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2]
    print("==", path, vals[names.index("Kernel Name")][:60])
    for k in KEYS:
        if k in names:
            i = names.index(k)
            print("  %-82s %14s %s" % (k, vals[i], units[i]))
