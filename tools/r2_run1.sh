#!/usr/bin/env bash
# round-2 GPU call 1: parity suite, legacy fused vs split A/B, new bench line, ncu of the fused kernel
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_gpu.txt 2>&1; tail -15 gpurun_out/r2a_pytest_gpu.txt
timeout 300 python bench.py --workload c4 --steps 30 --no-cpu-baseline > gpurun_out/r2a_bench_c4_fused.json 2> gpurun_out/r2a_bench_c4_fused.err; cut -c1-1800 gpurun_out/r2a_bench_c4_fused.json; tail -3 gpurun_out/r2a_bench_c4_fused.err
MCRAW_LEGACY_SPLIT=1 timeout 300 python bench.py --workload c4 --steps 30 --no-cpu-baseline > gpurun_out/r2a_bench_c4_split.json 2> gpurun_out/r2a_bench_c4_split.err; cut -c1-600 gpurun_out/r2a_bench_c4_split.json
for c in 4 6 8; do MCRAW_LGF_CTAS_PER_SM=$c timeout 300 python bench.py --workload c4 --steps 30 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ctas/sm $c', d['ms_per_step'])"; done
timeout 900 python bench.py --steps 20 > gpurun_out/r2a_bench_all.json 2> gpurun_out/r2a_bench_all.err; cut -c1-3000 gpurun_out/r2a_bench_all.json; tail -5 gpurun_out/r2a_bench_all.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err; cut -c1-1500 gpurun_out/r2a_bench_ref.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_legacy_fused -s 6 -c 1 -f -o gpurun_out/r2a_ncu_k_legacy_fused \
    python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_ncu_k_legacy_fused.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2a_launches_c4.csv \
    python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_launches_c4.log 2>&1
tail -5 gpurun_out/r2a_launches_c4.csv
