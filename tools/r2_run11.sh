#!/usr/bin/env bash
# round-2 GPU call 11: whole GPU test suite, the default bench line (all workloads) and the reference arm with the v13 legacy kernel
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2o}
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -5 gpurun_out/${T}_pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/${T}_bench_all.json 2> gpurun_out/${T}_bench_all.err; python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench_all.json").read().strip().splitlines()[-1])
print("c2", round(d["ms_per_step"],4), d["value"], "frac", round(d["roofline"]["frac"],3), "whole", round(d["roofline"]["whole_step"]["frac"],3), "e2e", round(d["e2e"]["value"]), "verified", d.get("pixels_verified"))
for k,v in d.get("workloads",{}).items():
    print(k, {kk:(round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ("ms_per_step","value","pixels_verified")}, (v.get("roofline") or {}).get("whole_step",{}).get("frac"), (v.get("e2e") or {}).get("value"))
PY
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; tail -c 600 gpurun_out/${T}_bench_ref.json
echo done
