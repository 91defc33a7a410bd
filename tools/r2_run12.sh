#!/usr/bin/env bash
# round-2 GPU call 12: epilogue constants in registers -- parity of the epilogue tests, cost on C2 (raw / black-subtracted / half), k_units profile with the epilogue
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2p}
timeout 600 python -m pytest tests/test_gpu_epilogue.py tests/test_gpu_legacy.py tests/test_golden.py -m gpu -q -x --timeout 200 > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -4 gpurun_out/${T}_pytest_gpu.txt
{
timeout 100 python tools/c2_steps.py --label raw
timeout 100 python tools/c2_steps.py --label black_sub_u16 --levels 1
timeout 100 python tools/c2_steps.py --label norm_f16 --levels 2
} > gpurun_out/${T}_c2_epi.jsonl 2> gpurun_out/${T}_c2_epi.err
cat gpurun_out/${T}_c2_epi.jsonl; tail -2 gpurun_out/${T}_c2_epi.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_units -s 4 -c 1 -f -o gpurun_out/${T}_ncu_k_units_epi \
    python tools/c2_steps.py --label ncu --levels 1 --steps 4 > gpurun_out/${T}_ncu_k_units_epi.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_units -s 4 -c 1 -f -o gpurun_out/${T}_ncu_k_units_raw \
    python tools/c2_steps.py --label ncu --steps 4 > gpurun_out/${T}_ncu_k_units_raw.log 2>&1
echo done
