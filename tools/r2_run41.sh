#!/usr/bin/env bash
# round-2 GPU call 41: fused epilogue with k_units<EPI> compiled for 4 CTAs per SM (128 registers, ~150 bytes of spills) against 3 (168 registers)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2e}
{
python tools/c2_steps.py --label raw
python tools/c2_steps.py --levels 1 --label black_sub_3ctas
python tools/c2_steps.py --levels 2 --label norm_f16_3ctas
MCRAW_B200_LIB=libmcraw_b200_epi4.so python tools/c2_steps.py --levels 1 --label black_sub_4ctas
MCRAW_B200_LIB=libmcraw_b200_epi4.so python tools/c2_steps.py --levels 2 --label norm_f16_4ctas
} > gpurun_out/${T}_epi_ab.jsonl 2> gpurun_out/${T}_epi_ab.err
cut -c1-160 gpurun_out/${T}_epi_ab.jsonl; tail -3 gpurun_out/${T}_epi_ab.err
echo done
