#!/usr/bin/env bash
# round-2 GPU call 18: compute-sanitizer over the round-2 kernels (k_legacy_warp with bulk stores, k_meta_split, chained launches)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2z}
sel="tests/test_golden.py tests/test_gpu_legacy.py tests/test_gpu_meta_split.py tests/test_gpu_epilogue.py"
for tool in memcheck racecheck synccheck; do
    timeout 1200 compute-sanitizer --tool $tool python -m pytest $sel -m gpu -q --timeout 1000 \
        -k "golden or vectors_batched or rejects or mixed or constant_width or one_width or split or epilogue" \
        > gpurun_out/${T}_sanitizer_$tool.txt 2>&1
    grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|error" gpurun_out/${T}_sanitizer_$tool.txt | tail -4
done
echo done
