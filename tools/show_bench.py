#!/usr/bin/env python
"""Print the headline numbers and the per-kernel roofline table of a bench.py JSON line."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["metric"], "| value", round(d["value"], 1), d["unit"], "| ms/step", round(d["ms_per_step"], 3), "| e2e", round(d["e2e"]["value"], 1),
      "| launches", d["gpu_launches"], "|", d["config"].get("launch", "")[:60])
for r in d.get("roofline_all", []):
    print(f"  {r['kernel']:9s} {r['geometry']:26s} {r['avg_us']:9.1f} us  x{r['launches_per_step']:<2d} {r['achieved']:8.1f} {r['unit']:8s} frac {r['frac']:.3f}  share {r['share_of_step']:.3f}")
for k in ("cpu_baseline", "reference_cutlass_ext", "clocks"):
    if k in d:
        print(" ", k, json.dumps(d[k])[:400])
