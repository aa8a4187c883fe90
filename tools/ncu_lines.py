#!/usr/bin/env python3
"""Aggregate an ncu report's per-SASS-instruction counters by CUDA source line.

ncu's CSV source page is per SASS instruction; nvdisasm -gi of the same cubin carries '//## File ..., line N'
markers.  Both list the kernel's instructions in address order, so they can be zipped.
usage: ncu_lines.py <report.ncu-rep> <kernel regex> <lib.so> [top N]
"""
import csv, os, re, subprocess, sys, tempfile
rep, kern, lib = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
dis = []
for f in sorted(os.listdir(tmp)):          # one cubin per translation unit: take the one that holds the kernel
    if not f.endswith(".cubin"):
        continue
    d = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, f)], capture_output=True, text=True).stdout.splitlines()
    if any(l.startswith("//----") and ".text." in l and re.search(kern, l) for l in d):
        dis = d
        break
# locate the kernel's section
start = next(i for i, l in enumerate(dis) if l.startswith("//----") and ".text." in l and re.search(kern, l))
lines, cur, in_group = [], None, False
for l in dis[start + 1:]:
    if l.startswith("//----"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        if not in_group:          # a group lists the innermost (inlined) location first
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
        in_group = True
        continue
    in_group = False
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ci, cs, cw = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("L1 Wavefronts Shared")
body = [r for r in rows[h + 1:] if len(r) > ci]
# several launches may be concatenated: take the first block with len(lines) instructions
n = len(lines)
if len(body) < n:
    print("warning: ncu has %d instructions, nvdisasm %d" % (len(body), n)); n = min(n, len(body))
agg = {}
tot_i = tot_s = 0
for k in range(n):
    r = body[k]
    inst, smp = int(float(r[ci] or 0)), int(float(r[cs] or 0))
    wf = int(float(r[cw] or 0)) if r[cw] not in ("", "-") else 0
    a = agg.setdefault(lines[k], [0, 0, 0])
    a[0] += inst; a[1] += smp; a[2] += wf
    tot_i += inst; tot_s += smp
src = {}
def text(key):
    if key is None: return ""
    f, ln = key
    for d in ("ffmpeg-heaac_b200/csrc",):
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, f)
        if os.path.exists(p):
            if p not in src: src[p] = open(p).read().splitlines()
            return src[p][ln - 1].strip()[:90]
    return ""
print("kernel %s: %d SASS instr, %.1fM warp-inst executed, %d samples" % (kern, n, tot_i / 1e6, tot_s))
print("%-26s %8s %6s %6s %9s  %s" % ("file:line", "winst(k)", "inst%", "smp%", "smem wf(k)", "source"))
for key, (i, s, w) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-26s %8.0f %6.1f %6.1f %9.0f  %s" % ("%s:%d" % key if key else "?", i / 1e3, 100.0 * i / max(tot_i, 1), 100.0 * s / max(tot_s, 1), w / 1e3, text(key)))
