#!/usr/bin/env python
"""Aggregate an ncu SASS source page by CUDA source line.

    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:<k> > sass.csv
    python tools/ncu_lines.py sass.csv <libnlkalman_b200.so> <mangled-kernel-substring> [launch_index]

Joins ncu's per-instruction counters (instructions executed, stall samples) with the
line table that nvdisasm prints for the cubin embedded in the library (built with
-lineinfo), by instruction offset.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def line_table(so, kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True,
                   stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    table, cur, inside = {}, None, False
    for ln in txt.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            inside = kernel_sub in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if m:
            table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def main():
    sass_csv, so, ksub = sys.argv[1:4]
    which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    table = line_table(so, ksub)
    rows = list(csv.reader(open(sass_csv)))
    # split into kernels (each starts with a "Kernel Name" row followed by the header)
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None:
            cur["rows"].append(r)
    blocks = [b for b in blocks if ksub_match(b["name"], ksub)]
    b = blocks[which]
    h = {n: i for i, n in enumerate(b["hdr"])}
    base = int(b["rows"][0][h["Address"]], 16) if b["rows"][0][h["Address"]].startswith("0x") else int(b["rows"][0][h["Address"]])
    agg = defaultdict(lambda: [0, 0, 0])
    tot_i = tot_s = 0
    for r in b["rows"]:
        a = r[h["Address"]]
        off = (int(a, 16) if a.startswith("0x") else int(a)) - base
        key = table.get(off, (("?", 0), ""))[0] or ("?", 0)
        ie = int(float(r[h["Instructions Executed"]] or 0))
        ss = int(float(r[h["# Samples"]] or 0))
        agg[key][0] += ie
        agg[key][1] += ss
        agg[key][2] += 1
        tot_i += ie
        tot_s += ss
    print(f"kernel: {b['name']}\ntotal warp-instructions {tot_i:,}  samples {tot_s:,}")
    print(f"{'file:line':32s} {'instr%':>7s} {'samp%':>7s} {'#sass':>6s}")
    for key, (ie, ss, n) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(os.environ.get("NLINES","45"))]:
        print(f"{key[0] + ':' + str(key[1]):32s} {100 * ie / max(tot_i, 1):7.2f} {100 * ss / max(tot_s, 1):7.2f} {n:6d}")


def ksub_match(name, ksub):
    # ksub is a mangled substring like k_group_filterILi8ELi3E or k_resolveILi1ELb1E: compare
    # the base name and the template arguments, in order, with the demangled name
    m = re.match(r"(\w+?)I((?:L[ib]\d+E)+)", ksub)
    if not m:
        return ksub in name
    want = re.findall(r"L[ib](\d+)E", m.group(2))
    d = re.search(re.escape(m.group(1)) + r"<([^>]*)>", name)
    if not d:
        return False
    have = re.findall(r"\)?(\d+)", d.group(1))
    return have[:len(want)] == want


if __name__ == "__main__":
    main()
