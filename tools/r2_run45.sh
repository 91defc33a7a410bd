#!/usr/bin/env bash
# round-2 GPU call 45: units per work item of k_units (24 is the default): 12 / 16 / 31 on the chained C2 step
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2k2}
{
python tools/c2_steps.py --label upw24
for u in 12 16 31; do MCRAW_B200_LIB=libmcraw_b200_upw$u.so python tools/c2_steps.py --label upw$u; done
python tools/c2_steps.py --label upw24_again
} | cut -c1-150 | tee gpurun_out/${T}_upw_ab.jsonl
echo done
