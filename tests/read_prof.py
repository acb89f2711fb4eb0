"""Summarise a gpurun_out ncu report + launch list (used to write profiles/*.md)."""
import csv, json, subprocess, sys, collections
rep = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/prof_pk.ncu-rep"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
r = list(csv.reader(raw.splitlines()))
hdr, units = r[0], r[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]
for row in r[2:]:
    print("----")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print("%-80s %s %s" % (k, row[i][:70], units[i]))
if len(sys.argv) > 2:
    lines = [l for l in open(sys.argv[2]) if not l.startswith("==")]
    rr = csv.reader(lines); h = next(rr)
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for row in rr:
        if len(row) > vi:
            nm = row[ki].split("(")[0]
            agg.setdefault(nm, []).append(float(row[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    for nm, v in agg.items():
        print("%-50s n=%3d  avg %10.1f us   share %5.1f%%" % (nm, len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
