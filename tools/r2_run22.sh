#!/usr/bin/env bash
# round-2 GPU call 22: the whole GPU suite and the bench line with k_meta_warp in place; compute-sanitizer over the new kernel
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2k}
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/${T}_bench_all.json 2> gpurun_out/${T}_bench_all.err; tail -c 300 gpurun_out/${T}_bench_all.json
for tool in memcheck racecheck; do
    timeout 900 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_meta_warp.py -m gpu -q --timeout 800 \
        -k "vectors_through or rejects" > gpurun_out/${T}_sanitizer_${tool}_meta_warp.txt 2>&1
    grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/${T}_sanitizer_${tool}_meta_warp.txt | tail -3
done
echo done
