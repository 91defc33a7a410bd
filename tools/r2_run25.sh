#!/usr/bin/env bash
# round-2 GPU call 25: evidence for the tree with k_meta_warp -- GPU suite, both bench arms, launch list, ncu of k_meta_warp and k_units
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2m}
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/${T}_bench_all.json 2> gpurun_out/${T}_bench_all.err; python tools/bench_summary.py gpurun_out/${T}_bench_all.json
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; tail -c 200 gpurun_out/${T}_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_c2.csv \
    python bench.py --workload c2 --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_launches_c2.log 2>&1
for k in k_meta_warp k_units; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 30 -c 1 -f -o gpurun_out/${T}_ncu_$k \
        python tools/c2_steps.py --steps 20 > gpurun_out/${T}_ncu_$k.log 2>&1
    python tools/ncu_summary.py gpurun_out/${T}_ncu_$k.ncu-rep > gpurun_out/${T}_ncu_$k.txt 2>&1
    head -5 gpurun_out/${T}_ncu_$k.txt
done
echo done
