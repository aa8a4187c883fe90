#!/usr/bin/env python3
"""Per-source-line and per-phase warp-instruction table of one kernel in an ncu report (needs -lineinfo and
--import-source on).  Uses ncu's own CUDA<->SASS correlation (`--print-source cuda,sass`): every SASS instruction is
attributed to the innermost source line that ncu lists it under in the file given by --file (inlined helpers from
other headers are charged to their call site in that file when ncu nests them there).

usage: ncu_phases.py <report.ncu-rep> <kernel regex> --file heaac_sbr.cuh [--phases "name:lo-hi,name:lo-hi,..."] [--top N]
"""
import argparse, csv, subprocess, sys

ap = argparse.ArgumentParser()
ap.add_argument("rep"); ap.add_argument("kernel")
ap.add_argument("--file", required=True)
ap.add_argument("--phases", default="")
ap.add_argument("--top", type=int, default=30)
ap.add_argument("--launch", type=int, default=0, help="which matching launch of the report")
a = ap.parse_args()
out = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", "regex:" + a.kernel], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
# split into per-(launch, file) blocks
blocks, cur = [], None
for r in rows:
    if r and r[0] == "File Path":
        cur = {"file": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
mine = [b for b in blocks if b["file"].endswith(a.file)]
if not mine:
    sys.exit("no block for file %s (have: %s)" % (a.file, sorted(set(b["file"] for b in blocks))))
b = mine[min(a.launch, len(mine) - 1)]
hdr = next(r for r in b["rows"] if r and r[0] == "Line No")
ci, cs, cw, ct = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("L1 Wavefronts Shared"), hdr.index("Thread Instructions Executed")
seen, lines, src = set(), {}, {}
line = None
for r in b["rows"]:
    if len(r) <= ci or r[0] == "Line No":
        continue
    if r[0]:
        line = int(r[0]); src[line] = r[1].strip()
        continue
    if r[2] in ("...", "-", "") or line is None:
        continue
    key = r[2]
    if key in seen:           # an instruction listed under several lines: first (outermost in this file) wins
        continue
    seen.add(key)
    f = lambda x: float(x) if x not in ("", "-") else 0.0
    l = lines.setdefault(line, [0.0, 0.0, 0.0, 0.0])
    l[0] += f(r[ci]); l[1] += f(r[cs]); l[2] += f(r[cw]); l[3] += f(r[ct])
tot = [sum(v[i] for v in lines.values()) for i in range(4)]
print("kernel %s, file %s: %.2fM warp-instructions attributed, %d samples, lanes active %.1f/32" %
      (a.kernel, a.file, tot[0] / 1e6, tot[1], tot[3] / max(tot[0], 1)))
if a.phases:
    print("%-28s %10s %6s %6s %12s" % ("phase (lines)", "winst(k)", "inst%", "smp%", "smem wf(k)"))
    used = set()
    for p in a.phases.split(","):
        name, rng = p.split(":"); lo, hi = map(int, rng.split("-"))
        sel = [k for k in lines if lo <= k <= hi]; used.update(sel)
        s = [sum(lines[k][i] for k in sel) for i in range(3)]
        print("%-28s %10.0f %6.1f %6.1f %12.0f" % ("%s (%d-%d)" % (name, lo, hi), s[0] / 1e3, 100 * s[0] / tot[0], 100 * s[1] / max(tot[1], 1), s[2] / 1e3))
    rest = [k for k in lines if k not in used]
    s = [sum(lines[k][i] for k in rest) for i in range(3)]
    print("%-28s %10.0f %6.1f %6.1f %12.0f" % ("(other lines)", s[0] / 1e3, 100 * s[0] / tot[0], 100 * s[1] / max(tot[1], 1), s[2] / 1e3))
print("%-6s %10s %6s %6s %10s  %s" % ("line", "winst(k)", "inst%", "smp%", "smem wf(k)", "source"))
for k, v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:a.top]:
    print("%-6d %10.0f %6.1f %6.1f %10.0f  %s" % (k, v[0] / 1e3, 100 * v[0] / tot[0], 100 * v[1] / max(tot[1], 1), v[2] / 1e3, src.get(k, "")[:100]))
