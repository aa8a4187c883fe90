#!/usr/bin/env python3
"""Append a one-line-per-kernel summary of an ncu report to a markdown file.
usage: ncu_summary.py <report.ncu-rep> "<title>" [out.md]"""
import csv, subprocess, sys
rep, title = sys.argv[1], sys.argv[2]
out = sys.argv[3] if len(sys.argv) > 3 else None
cols = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(raw))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
lines = ["", "## %s  (%s)" % (rep.split("/")[-1], title), "",
         "units: " + ", ".join("%s=%s" % (c, units[hdr.index(c)]) for c in cols if c in hdr and units[hdr.index(c)]), ""]
for r in rows[2:]:
    lines.append("- kernel=%s; " % r[ki][:28] + "; ".join("%s=%s" % (c.split(".")[0].replace("__", ".", 1).split(".", 1)[1] if False else c.split(".")[0], r[hdr.index(c)]) for c in cols if c in hdr))
text = "\n".join(lines) + "\n"
if out:
    open(out, "a").write(text)
print(text)
