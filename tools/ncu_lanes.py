#!/usr/bin/env python3
"""Per source line of one kernel in an ncu report: warp-instructions executed, average active lanes, share of stall samples.
usage: ncu_lanes.py <report.ncu-rep> <kernel regex> <lib.so> [top N]"""
import csv, subprocess, re, os, tempfile, sys
rep, kern, lib = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
dis = []
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin"):
        continue
    d = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, f)], capture_output=True, text=True).stdout.splitlines()
    if any(l.startswith("//----") and ".text." in l and re.search(kern, l) for l in d):
        dis = d
        break
start = next(i for i, l in enumerate(dis) if l.startswith("//----") and ".text." in l and re.search(kern, l))
lines, cur, ing = [], None, False
for l in dis[start + 1:]:
    if l.startswith("//----"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        if not ing:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
        ing = True
        continue
    ing = False
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ci, ct, cs = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
body = [r for r in rows[h + 1:] if len(r) > ct]
agg = {}
for k in range(min(len(lines), len(body))):
    a = agg.setdefault(lines[k], [0, 0, 0])
    a[0] += int(float(body[k][ci] or 0)); a[1] += int(float(body[k][ct] or 0)); a[2] += int(float(body[k][cs] or 0))
ti, ts = sum(a[0] for a in agg.values()), sum(a[2] for a in agg.values())
print("%s: %.1f M warp-instructions, %.1f lanes on average, %d samples" % (kern, ti / 1e6, sum(a[1] for a in agg.values()) / max(ti, 1), ts))
src = {}
def text(key):
    if not key: return ""
    f, ln = key
    p = os.path.join("ffmpeg-heaac_b200/csrc", f)
    if p not in src:
        src[p] = open(p).read().splitlines() if os.path.exists(p) else []
    return src[p][ln - 1].strip()[:90] if 0 < ln <= len(src[p]) else ""
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print("%-22s winst %7.0fk %5.1f%%  lanes %4.1f  samples %5.1f%%  %s" % ("%s:%d" % key if key else "?", a[0] / 1e3, 100 * a[0] / ti, a[1] / max(a[0], 1), 100 * a[2] / max(ts, 1), text(key)))
