#!/usr/bin/env python3
"""Warp-stall sampling of one kernel in an ncu report, summed over all its SASS instructions (which stall reasons cost the most
samples) and the same per source file.   usage: ncu_stalls.py <report.ncu-rep> <kernel regex>"""
import csv, subprocess, sys
rep, kern = sys.argv[1:3]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
cols = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
tot = {hdr[i]: 0.0 for i in cols}
n = 0
for r in rows[h + 1:]:
    if len(r) <= cols[-1] or r[0] == "Address":
        break
    for i in cols:
        if r[i] not in ("", "-"):
            tot[hdr[i]] += float(r[i])
    n += 1
s = sum(tot.values())
print("kernel %s: %d SASS instructions, %d stall samples" % (kern, n, s))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v:
        print("  %-24s %5.1f %%" % (k, 100 * v / s))
