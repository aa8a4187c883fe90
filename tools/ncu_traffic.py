#!/usr/bin/env python3
"""gpurun_out/traffic_<workload>_<streams>_<mode>.csv (tools/ncu_traffic.sh) -> profiles/r02_dram_bytes.json:
per (kernel, workload, streams): DRAM bytes (read + write) per unit of work, averaged over the launches of the capture's last
step (state warm: every launch both reads and writes the channel state), for both cache-control modes.
usage: ncu_traffic.py [frames per step of the captures, default 8]"""
import csv, glob, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FR = int(sys.argv[1]) if len(sys.argv) > 1 else 8
NCC = {"heaac_v1_stereo": 2, "heaac_v2_ps": 1, "lc_51_48k_tns": 6, "lc_stereo_44k": 2}
NOUT = {"heaac_v1_stereo": 2, "heaac_v2_ps": 2, "lc_51_48k_tns": 6, "lc_stereo_44k": 2}
out = {}
for path in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "traffic_*.csv"))):
    m = re.match(r"traffic_(.+)_(\d+)_(all|none)\.csv", os.path.basename(path))
    wl, S, mode = m.group(1), int(m.group(2)), m.group(3)
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    ii = hdr.index("ID")
    launches = {}
    for r in rows[1:]:
        d = launches.setdefault(int(r[ii]), {"kernel": r[ki].split("(")[0].split("<")[0].replace("void ", "").strip()})
        v = float(r[vi].replace(",", ""))
        u = r[ui].lower()
        if "byte" in u:
            v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
        d[r[mi]] = v
    ids = sorted(launches)
    per_kernel = {}
    for i in ids:
        per_kernel.setdefault(launches[i]["kernel"], []).append(launches[i])
    for k, ls in per_kernel.items():
        # launches per step of this kernel: the capture holds warm-up steps + the timed step + the instrumented pass; take
        # the last quarter of the launches (steady state)
        tail = ls[-max(1, len(ls) // 4):]
        b = sum(l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"] for l in tail) / len(tail)
        if k == "k_chain":
            units = S * NCC[wl] * (FR if wl == "heaac_v1_stereo" else 1)
        elif k == "k_ps":
            units = S
        elif k == "k_synth":
            units = S * NOUT[wl]
        elif k in ("k_core", "k_tns"):
            units = S * NCC[wl]
        else:
            continue
        e = out.setdefault("%s|%s|%d" % (k, wl, S), {})
        e["dram_bytes_per_unit" if mode == "none" else "dram_bytes_per_unit_cold"] = b / units
        e["source"] = "ncu dram__bytes_read.sum + dram__bytes_write.sum, --cache-control none (tools/ncu_traffic.sh, %d frames per step; " \
                      "'_cold' = ncu's default cache flush before every kernel)" % FR
        e["launch_us_%s" % mode] = sum(l["gpu__time_duration.sum"] for l in tail) / len(tail) / (1e3 if tail[0]["gpu__time_duration.sum"] > 1e4 else 1)
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_dram_bytes.json"), "w"), indent=1, sort_keys=True)
for k, v in sorted(out.items()):
    print(k, {a: (round(b, 1) if isinstance(b, float) else b) for a, b in v.items() if a != "source"})
