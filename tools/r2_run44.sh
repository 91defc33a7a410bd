#!/usr/bin/env bash
# round-2 GPU call 44: CTAs held back for the next batch's index kernel (k_meta_warp): 24 (default) / 32 / 40 on C2 and C3
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2g}
{
for c in 24 32 40; do MCRAW_CHAIN=$c python tools/c2_steps.py --label c2_chain$c | cut -c1-120; done
for c in 24 32 40; do
    MCRAW_CHAIN=$c timeout 300 python bench.py --workload c3 --no-cpu-baseline --steps 20 > gpurun_out/${T}_c3.json 2> gpurun_out/${T}_c3.err
    python -c "
import json;d=json.load(open('gpurun_out/${T}_c3.json'));print('c3 chain $c: step', round(d['ms_per_step'],4), 'ms, whole', round(d['roofline']['whole_step']['frac'],4))"
done
} | tee gpurun_out/${T}_chain_holdback.txt
echo done
