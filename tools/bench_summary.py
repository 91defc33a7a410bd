#!/usr/bin/env python
"""bench_summary.py -- one line per workload of a bench.py JSON line (file argument): step, throughput, roofline fractions, e2e."""
import json
import sys


def show(k, w):
    r = w.get("roofline") or {}
    ws = r.get("whole_step") or {}
    print(f"{k}: ms {w['ms_per_step']:.4f}  value {w['value']:.0f}  kernel frac {r.get('frac', 0):.3f}  whole {ws.get('frac', 0):.3f}  "
          f"kernel_ms {r.get('kernel_ms_per_launch')}  index_ms {r.get('index_kernels_ms_per_launch')}  "
          f"e2e {(w.get('e2e') or {}).get('value', 0):.0f}  host_out {(w.get('e2e_host_out') or {}).get('value', 0):.0f}")


d = json.load(open(sys.argv[1]))
show("c2", d)
for k, w in (d.get("workloads") or {}).items():
    try:
        show(k, w)
    except Exception:
        print(k, {a: b for a, b in w.items() if not isinstance(b, dict)})
print("clocks", d.get("clocks"), "frac_of_h2d", (d.get("e2e") or {}).get("frac_of_h2d"), "pixels_verified", d.get("pixels_verified"),
      "cold_plan_ms", d.get("cold_plan_ms_per_step"), "cold_same_outputs_ms", d.get("cold_plan_same_outputs_ms_per_step"))
