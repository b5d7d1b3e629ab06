"""Turn the ncu outputs of scripts/gpu_profile.sh into the tracked summaries under profiles/.

    python scripts/summarise_profile.py <tag>

  profiles/<tag>_launches.md      per-kernel device time of ONE build (the last complete one in the list),
                                  cold-cache and serialised: the SHARES are what compares with bench.py
  profiles/<tag>_insert_kernel.md key counters of the full capture of k_insert_windows
  profiles/insert_kernel_traffic.json   dram bytes per launch (bench.py's roofline.traffic reads it)
"""
import csv, json, os, re, subprocess, sys

tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)


def short(name):
    name = re.sub(r"^void ", "", name)
    m = re.match(r"(amira::\w+)", name)
    if m:
        return m.group(1)
    m = re.match(r"(cub::\w+)", name)
    return m.group(1) if m else name[:60]


lp = os.path.join(ROOT, "gpurun_out", "launches_%s.csv" % tag)
if os.path.exists(lp):
    rows = [r for r in csv.reader(open(lp)) if len(r) > 10 and r[0].isdigit()]
    names = [short(r[4]) for r in rows]
    ns = [float(r[-1].replace(",", "")) for r in rows]
    starts = [i for i, n in enumerate(names) if n == "amira::k_read_windows"]
    if len(starts) >= 2:
        a, b = starts[-2], starts[-1]          # the last complete build
    else:
        a, b = (starts[0] if starts else 0), len(rows)
    agg, order = {}, []
    for n, t in zip(names[a:b], ns[a:b]):
        if n not in agg:
            agg[n] = [0, 0.0]
            order.append(n)
        agg[n][0] += 1
        agg[n][1] += t
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(out, "%s_launches.md" % tag), "w") as f:
        f.write("# ncu launch list, one build (%s)\n\n" % tag)
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py --steps 2 --warmup 3`; "
                "launches %d..%d of %d = the last complete build. Per-launch times under ncu are cold-cache and "
                "serialised; compare shares.\n\n" % (a, b - 1, len(rows)))
        f.write("| kernel | launches | device time (us) | share |\n|---|---|---|---|\n")
        for n in sorted(order, key=lambda n: -agg[n][1]):
            f.write("| `%s` | %d | %.1f | %.1f%% |\n" % (n, agg[n][0], agg[n][1] / 1e3, 100 * agg[n][1] / tot))
        f.write("| total | %d | %.1f | 100%% |\n" % (b - a, tot / 1e3))
    print("wrote launches summary; kernels in one build:", b - a, "total us", tot / 1e3)

rp = os.path.join(ROOT, "gpurun_out", "prof_%s.ncu-rep" % tag)
if os.path.exists(rp):
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "smsp__inst_executed.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
            "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors.sum",
            "l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum", "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum",
            "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__issue_active.avg.pct_of_peak_sustained_active"]
    idx = {h: i for i, h in enumerate(hdr)}
    got = {}
    with open(os.path.join(out, "%s_insert_kernel.md" % tag), "w") as f:
        f.write("# ncu --set full, k_insert_windows (%s)\n\n" % tag)
        f.write("`ncu --set full --clock-control none --import-source on -k regex:k_insert_windows -s 3 -c 1` over "
                "`bench.py --steps 2 --warmup 3` (C5 shard: 1.25M reads x 30 calls, k=5).\n\n| metric | value | unit |\n|---|---|---|\n")
        for w in want:
            if w in idx:
                got[w] = vals[idx[w]]
                f.write("| %s | %s | %s |\n" % (w, vals[idx[w]], units[idx[w]]))

    def tobytes(name):
        v, u = float(got[name].replace(",", "")), units[idx[name]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
    if "dram__bytes_read.sum" in got:
        tr = {"dram_bytes_per_launch": int(tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")),
              "dram_read": int(tobytes("dram__bytes_read.sum")), "dram_write": int(tobytes("dram__bytes_write.sum")),
              "source": "profiles/%s_insert_kernel.md" % tag}
        json.dump(tr, open(os.path.join(out, "insert_kernel_traffic.json"), "w"))
        print(tr)
