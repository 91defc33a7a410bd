#!/usr/bin/env bash
# round-2 GPU call 36: cold-plan steps of the C2 bench leg (alternating output areas / same outputs), side-stream upload on and off
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2x}
for v in "" "MCRAW_PLAN_SIDE=0"; do
    env $v timeout 300 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/${T}_c2.json 2> gpurun_out/${T}_c2.err || tail -5 gpurun_out/${T}_c2.err
    echo "== $v"; python -c "
import json;d=json.load(open('gpurun_out/${T}_c2.json'));print('step',round(d['ms_per_step'],4),'cold alternating',round(d['cold_plan_ms_per_step'],4),'cold same outputs',round(d['cold_plan_same_outputs_ms_per_step'],4), d['pixels_verified'])"
done | tee gpurun_out/${T}_cold_plans.txt
echo done
