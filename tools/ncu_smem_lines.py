#!/usr/bin/env python
"""Shared-memory wavefronts by CUDA source line, from an ncu SASS source page.

    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:<k> > sass.csv
    python tools/ncu_smem_lines.py sass.csv <lib.so> <mangled-kernel-substring>
"""
import csv, sys
from collections import defaultdict
sys.path.insert(0, __import__("os").path.dirname(__file__))
from ncu_lines import line_table

sass_csv, so, ksub = sys.argv[1:4]
table = line_table(so, ksub)
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
h = {n: i for i, n in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) > 10 and r[0].startswith("0x")]
base = int(body[0][h["Address"]], 16)
agg = defaultdict(lambda: [0, 0, 0, 0])
tw = ti = 0
for r in body:
    off = int(r[h["Address"]], 16) - base
    key = (table.get(off, (("?", 0), ""))[0] or ("?", 0))
    op = r[h["Source"]].split()[0] if r[h["Source"]].split() else "?"
    if op.startswith("@"):
        op = r[h["Source"]].split()[1]
    w = int(float(r[h["L1 Wavefronts Shared"]] or 0))
    wi = int(float(r[h["L1 Wavefronts Shared Ideal"]] or 0))
    ie = int(float(r[h["Instructions Executed"]] or 0))
    if w == 0:
        continue
    k = (key, op.split(".")[0] + ("." + op.split(".")[-1] if op.split(".")[-1] in ("64", "128") else ""))
    agg[k][0] += w; agg[k][1] += wi; agg[k][2] += ie; agg[k][3] += 1
    tw += w; ti += wi
print(f"shared wavefronts {tw:,}  ideal {ti:,}")
print(f"{'file:line':30s} {'op':10s} {'wavefr%':>8s} {'wavefronts':>12s} {'ideal':>12s} {'instr':>11s} {'w/inst':>7s} {'#sass':>5s}")
for (key, op), (w, wi, ie, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:30]:
    print(f"{key[0] + ':' + str(key[1]):30s} {op:10s} {100*w/tw:8.2f} {w:12,} {wi:12,} {ie:11,} {w/max(ie,1):7.2f} {n:5d}")
