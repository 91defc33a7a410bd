#!/usr/bin/env bash
# round-2 GPU call 10: legacy v9 (mul.hi sample extraction, leaders-only pair list, early ticket) -- parity, A/B, CTAs/SM, profile
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2j}
timeout 600 python -m pytest tests/test_golden.py tests/test_gpu_legacy.py tests/test_gpu_fuzz.py tests/test_gpu_epilogue.py -m gpu -q -x --timeout 200 > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -6 gpurun_out/${T}_pytest_gpu.txt
c4() { timeout 120 python bench.py --workload c4 --steps 30 --no-cpu-baseline 2>gpurun_out/${T}_c4.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4 $1', round(d['ms_per_step'],4), 'verified', d['pixels_verified'], 'frac', round(d['roofline']['whole_step']['frac'],3))"; }
for lib in ${LIBS:-libmcraw_b200.so libmcraw_b200_lgv8.so libmcraw_b200_pf4.so libmcraw_b200_pf16.so}; do MCRAW_B200_LIB=$lib c4 $lib; done
for n in 6 7; do MCRAW_LGW_CTAS_PER_SM=$n c4 "ctas/sm=$n"; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_legacy_warp -s 6 -c 1 -f -o gpurun_out/${T}_ncu_k_legacy_warp \
    python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_k_legacy_warp.log 2>&1
echo done
