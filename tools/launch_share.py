#!/usr/bin/env python
"""Per-kernel share of the step from an ncu launch list (--metrics gpu__time_duration.sum --csv).

    python tools/launch_share.py gpurun_out/x_launches.csv > profiles/x_launch_share.txt

The per-launch times are cold-cache and serialised under the profiler, so only the SHARE of
each kernel is comparable with the CUDA-event shares that bench.py prints."""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 14 and r[0].isdigit()]
tot = defaultdict(lambda: [0.0, 0])
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("nlk::", "").replace("void ", "")
    tot[name][0] += float(r[14]) * 1e-6
    tot[name][1] += 1
total = sum(v[0] for v in tot.values())
groups = {"group_filter": ("k_group",), "search_knn": ("k_search",), "mask_resolve": ("k_resolve", "k_active", "k_set_flag"),
          "prep (colour, warp, valid, normalize)": ("k_rgb2opp", "k_opp2rgb", "k_warp", "k_valid", "k_normalize")}
print(f"{len(rows)} launches, {total:.3f} ms in kernels (under ncu: cold caches, serialised)\n")
print(f"{'kernel':60s} {'launches':>8s} {'ms':>9s} {'share':>7s}")
for name, (ms, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"{name:60s} {n:8d} {ms:9.3f} {ms / total:7.3f}")
print()
for g, keys in groups.items():
    ms = sum(v[0] for k, v in tot.items() if any(k.startswith(x) for x in keys))
    print(f"{g:60s} {'':8s} {ms:9.3f} {ms / total:7.3f}")
