#!/usr/bin/env python3
"""Per source region of one kernel in an ncu report: warp-instructions, SASS size, warp-time samples and the main stall reasons.
usage: ncu_regions.py <report.ncu-rep> <kernel regex> "file:name:lo-hi,file:name:lo-hi,..."   (lines outside the listed ranges are
summed per file)"""
import csv, subprocess, sys, collections
rep, kern, spec = sys.argv[1:4]
regions = []
for p in spec.split(","):
    f, name, rng = p.split(":"); lo, hi = map(int, rng.split("-")); regions.append((f, name, lo, hi))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "File Path": cur = {"file": r[1].split("/")[-1], "rows": []}; blocks.append(cur)
    elif cur is not None: cur["rows"].append(r)
seen = set()
acc = collections.defaultdict(lambda: collections.Counter())
for b in blocks:
    hdr = next((r for r in b["rows"] if r and r[0] == "Line No"), None)
    if not hdr: continue
    ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
    st = {c: i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c}
    line = None
    for r in b["rows"]:
        if len(r) <= cs or r[0] == "Line No": continue
        if r[0]: line = int(r[0]); continue
        if r[2] in ("...", "-", "") or r[2] in seen: continue
        seen.add(r[2])
        key = next((n for f, n, lo, hi in regions if f == b["file"] and lo <= line <= hi), "(" + b["file"] + ")")
        f = lambda x: float(x) if x not in ("", "-") else 0.0
        a = acc[key]
        a["sass"] += 1; a["inst"] += f(r[ci]); a["smp"] += f(r[cs])
        for c, i in st.items(): a[c] += f(r[i])
tot = sum(a["smp"] for a in acc.values()); ti = sum(a["inst"] for a in acc.values())
print("%-22s %6s %9s %6s %6s | %5s %5s %5s %5s %5s %5s" % ("region", "sass", "winst(k)", "inst%", "time%", "barr", "noins", "wait", "lngsb", "shtsb", "sel"))
order = [n for _, n, _, _ in regions] + sorted(k for k in acc if k.startswith("("))
for k in order:
    a = acc.get(k)
    if not a: continue
    s = max(a["smp"], 1)
    print("%-22s %6d %9.0f %6.1f %6.1f | %5.0f %5.0f %5.0f %5.0f %5.0f %5.0f" % (k, a["sass"], a["inst"] / 1e3, 100 * a["inst"] / ti, 100 * a["smp"] / tot,
          100 * a["stall_barrier"] / s, 100 * a["stall_no_inst"] / s, 100 * a["stall_wait"] / s, 100 * a["stall_long_sb"] / s, 100 * a["stall_short_sb"] / s, 100 * a["stall_selected"] / s))
