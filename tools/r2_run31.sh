#!/usr/bin/env bash
# round-2 GPU call 31: C3 device-resident step, default against MCRAW_META_WARP=0 / MCRAW_CHAIN=0 (regression hunt)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2s}
for v in "" "MCRAW_META_WARP=0" "MCRAW_CHAIN=0" "MCRAW_CHAIN=48"; do
    env $v timeout 300 python bench.py --workload c3 --no-cpu-baseline --steps 12 > gpurun_out/${T}_c3.json 2> gpurun_out/${T}_c3.err
    echo "== $v"; python tools/bench_summary.py gpurun_out/${T}_c3.json | head -1
done
echo done
