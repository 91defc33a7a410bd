#!/usr/bin/env bash
# round-2 GPU call 35: new plans on a side stream + epoch-tagged done words (batches of new descriptors chain) -- parity, cold-plan step, sanitizers
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2w}
timeout 900 python -m pytest tests/test_gpu_cold_chain.py -m gpu -q --timeout 600 -x > gpurun_out/${T}_pytest_cold.txt 2>&1; tail -5 gpurun_out/${T}_pytest_cold.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.txt
for v in "" "MCRAW_PLAN_SIDE=0"; do
    env $v timeout 300 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/${T}_c2.json 2> gpurun_out/${T}_c2.err
    echo "== $v"; python tools/bench_summary.py gpurun_out/${T}_c2.json 2>/dev/null | tail -2 | cut -c1-400
done
for tool in memcheck racecheck; do
    timeout 900 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_cold_chain.py -m gpu -q --timeout 800 -k "few_frames or failed_frame" \
        > gpurun_out/${T}_sanitizer_${tool}_cold_chain.txt 2>&1
    grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/${T}_sanitizer_${tool}_cold_chain.txt | tail -3
done
echo done
