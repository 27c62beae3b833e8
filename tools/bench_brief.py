#!/usr/bin/env python
"""Print the headline numbers and the per-kernel table of a bench.py JSON line."""
import json
import sys

d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
print(f"value {d['value']:.1f} {d['unit']}  e2e {d['e2e']['value']:.1f}  ms/step {d['ms_per_step']:.3f}  "
      f"launches {d['gpu_launches']}  clocks {d.get('clocks')}")
for k in d["kernels"]:
    print(f"  {k['kernel']:14s} {k['pass']:14s} n={k['launches']:3d} avg {k['avg_ms']:.3f} ms  share {k['share_of_step']:.3f}"
          f"  frac_fp32 {k.get('frac_fp32_peak', 0):.3f}")
r = d["roofline"]
print(f"roofline {r['kernel']}: frac {r['frac']:.3f} alpha {r.get('alpha')} frac_active {r.get('frac_active')} whole step {r.get('whole_step')}")
print("e2e both outputs", d["e2e"].get("both_outputs_value"), "sync", d["e2e"]["synchronous"]["value"])
print("cpu_baseline", d.get("cpu_baseline"))
print("cli", d.get("cli"))
if "strips" in d:
    print("strips", d["strips"])
if "tvl1" in d:
    print("tvl1", d["tvl1"])
