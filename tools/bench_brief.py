#!/usr/bin/env python
"""Print the headline numbers and the per-kernel table of a bench.py JSON line."""
import json
import sys

d = json.load(open(sys.argv[1]))
print(f"value {d['value']:.1f} {d['unit']}  e2e {d['e2e']['value']:.1f}  ms/step {d['ms_per_step']:.3f}  "
      f"launches {d['gpu_launches']}  clocks {d.get('clocks')}")
for k in d["kernels"]:
    print(f"  {k['kernel']:14s} {k['pass']:14s} n={k['launches']:3d} avg {k['avg_ms']:.3f} ms  share {k['share_of_step']:.3f}"
          f"  frac_fp32 {k.get('frac_fp32_peak', 0):.3f}")
