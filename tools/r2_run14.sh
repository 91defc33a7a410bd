#!/usr/bin/env bash
# round-2 GPU call 14: ncu of k_meta_split on C1
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2r}
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_meta_split -s 4 -c 1 -f -o gpurun_out/${T}_ncu_k_meta_split \
    python bench.py --workload c1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_k_meta_split.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${T}_launches_c1.csv python bench.py --workload c1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_launches_c1.log 2>&1
tail -12 gpurun_out/${T}_launches_c1.csv | cut -d, -f5,12-
echo done
