#!/usr/bin/env bash
# round-2 GPU call 49: ncu --set full of k_units on the C3 workload (DRAM traffic per launch for bench.py's c3 roofline) and of k_legacy_warp (final build)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2n3}
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_units -s 8 -c 1 -f -o gpurun_out/${T}_ncu_k_units_c3 \
    python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_k_units_c3.log 2>&1
python tools/ncu_summary.py gpurun_out/${T}_ncu_k_units_c3.ncu-rep > gpurun_out/${T}_ncu_k_units_c3.txt 2>&1; head -8 gpurun_out/${T}_ncu_k_units_c3.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_legacy_warp -s 8 -c 1 -f -o gpurun_out/${T}_ncu_k_legacy_warp \
    python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_k_legacy_warp.log 2>&1
python tools/ncu_summary.py gpurun_out/${T}_ncu_k_legacy_warp.ncu-rep > gpurun_out/${T}_ncu_k_legacy_warp.txt 2>&1; head -8 gpurun_out/${T}_ncu_k_legacy_warp.txt
rm -f gpurun_out/${T}_*.ncu-rep
echo done
