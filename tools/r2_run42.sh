#!/usr/bin/env bash
# round-2 GPU call 42: half-float epilogue without I2F (PRMT into the mantissa of 2^23 + one exact subtraction) -- parity, C2 steps
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2f}
timeout 600 python -m pytest tests/test_gpu_epilogue.py -m gpu -q --timeout 300 > gpurun_out/${T}_pytest_epi.txt 2>&1; tail -3 gpurun_out/${T}_pytest_epi.txt
{
python tools/c2_steps.py --label raw
python tools/c2_steps.py --levels 1 --label black_sub
python tools/c2_steps.py --levels 2 --label norm_f16_prmt
} > gpurun_out/${T}_epi_ab.jsonl 2> gpurun_out/${T}_epi_ab.err
cut -c1-160 gpurun_out/${T}_epi_ab.jsonl; tail -3 gpurun_out/${T}_epi_ab.err
echo done
