#!/usr/bin/env python3
"""Opcode mix of a kernel in an ncu report, weighted by executed warp-instructions.
usage: ncu_opmix.py <report.ncu-rep> <kernel regex>"""
import csv, re, subprocess, sys
rep, kern = sys.argv[1:3]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]; ci = hdr.index("Instructions Executed"); si = hdr.index("Source")
mix = {}; tot = 0
for r in rows[h + 1:]:
    if len(r) <= ci or r[0] == "Address": continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[si])
    if not m: continue
    op = m.group(2); n = int(float(r[ci] or 0))
    mix[op] = mix.get(op, 0) + n; tot += n
groups = {"fp32": ("FADD", "FMUL", "FFMA", "MUFU", "FSETP", "FSEL", "FMNMX", "FCHK", "F2I", "I2F", "F2F", "FRND", "I2FP", "F2FP"),
          "int/addr": ("IMAD", "IADD3", "LEA", "LOP3", "SHF", "ISETP", "SEL", "MOV", "IABS", "VIADD", "VIMNMX", "PRMT", "IMNMX", "SGXT", "BMSK", "UIADD3", "UMOV", "ULEA", "UIMAD", "USHF", "ULOP3", "UISETP", "USEL", "R2UR", "S2R", "CS2R", "POPC", "FLO", "UFLO", "UPOPC", "S2UR", "PLOP3", "UPLOP3", "P2R", "R2P", "IDP", "UPRMT", "UMOV32I", "MOV32I", "USGXT", "UBMSK", "IADD", "UIADD"),
          "smem": ("LDS", "STS", "LDSM"), "global": ("LDG", "STG", "LD", "ST", "LDC", "LDCU", "ULDC"),
          "control": ("BRA", "BSSY", "BSYNC", "EXIT", "BAR", "WARPSYNC", "CALL", "RET", "NOP", "YIELD", "BREAK", "BRX", "JMP", "DEPBAR", "MEMBAR", "ERRBAR", "NANOSLEEP", "BMOV")}
g = {k: 0 for k in groups}; other = {}
for op, n in mix.items():
    for k, ops in groups.items():
        if op.split(".")[0] in ops: g[k] += n; break
    else: other[op] = n
print("kernel %s: %.1fM warp-instructions" % (kern, tot / 1e6))
for k, n in sorted(g.items(), key=lambda kv: -kv[1]): print("  %-9s %5.1f%%" % (k, 100.0 * n / tot))
print("  other     %5.1f%%  %s" % (100.0 * sum(other.values()) / tot, sorted(other.items(), key=lambda kv: -kv[1])[:8]))
print("  top ops:", ", ".join("%s %.1f%%" % (o, 100.0 * n / tot) for o, n in sorted(mix.items(), key=lambda kv: -kv[1])[:14]))
