#!/usr/bin/env bash
# round-2 GPU call 23: where k_units' time goes -- the kernel without the block decode, and without decode + emit (staging and copy-out only)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2l}
{
python tools/c2_steps.py --label base
MCRAW_B200_LIB=libmcraw_b200_nodecode.so python tools/c2_steps.py --label nodecode
MCRAW_B200_LIB=libmcraw_b200_noemit.so python tools/c2_steps.py --label noemit
} > gpurun_out/${T}_c2_ab.jsonl 2> gpurun_out/${T}_c2_ab.err
cut -c1-200 gpurun_out/${T}_c2_ab.jsonl; tail -3 gpurun_out/${T}_c2_ab.err
echo done
