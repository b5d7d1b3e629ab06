"""profiles/<tag>_kernels.md from the `ncu --set full` captures of scripts/ncu_kernel.sh:
python scripts/summarise_kernels.py <tag> k_partition k_unit_lists ..."""
import csv, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, kernels = sys.argv[1], sys.argv[2:]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_atom.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
with open(os.path.join(ROOT, "profiles", "%s_kernels.md" % tag), "w") as f:
    f.write("# ncu --set full, the passes after the insert kernel (%s)\n\n" % tag)
    f.write("`ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 2 -c 1` over `scripts/ncu_target.py "
            "c5 1250000 5 3` (device-resident builds of the C5 shard; third build: tables sized from the previous one).  "
            "Cold-cache, serialised; never a bench value.\n")
    for kern in kernels:
        path = os.path.join(ROOT, "gpurun_out", "prof_%s_%s.csv" % (tag, kern))
        if not os.path.exists(path):
            continue
        rows = list(csv.reader(open(path)))
        hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
        names, units, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2]
        f.write("\n## `%s`\n\n| metric | value | unit |\n|---|---|---|\n" % kern)
        for k in KEYS:
            if k in names:
                i = names.index(k)
                f.write("| %s | %s | %s |\n" % (k, vals[i], units[i]))
print("wrote profiles/%s_kernels.md" % tag)
