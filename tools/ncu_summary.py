"""Per-kernel summary of an `ncu --set full` report: `ncu -i X.ncu-rep --page raw --csv | python
tools/ncu_summary.py > profiles/<name>.csv` (one row per profiled launch)."""
import csv, sys
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
rows = list(csv.reader(sys.stdin))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
names, units = rows[hdr], rows[hdr + 1]
ix = {n: i for i, n in enumerate(names)}
cols = [k for k in KEEP if k in ix]
w = csv.writer(sys.stdout)
w.writerow(["kernel"] + cols)
w.writerow(["unit"] + [units[ix[k]] for k in cols])
for r in rows[hdr + 2:]:
    if len(r) == len(names):
        w.writerow([r[ix["Kernel Name"]][:60]] + [r[ix[k]] for k in cols])
