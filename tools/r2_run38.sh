#!/usr/bin/env bash
# round-2 GPU call 38: compute-sanitizer (memcheck, racecheck, synccheck) over the parity tests of every kernel with the final protocol
# (plan ready word, epoch-tagged done words, chained launches)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2z}
sel="tests/test_golden.py tests/test_gpu_current.py tests/test_gpu_legacy.py tests/test_gpu_meta_split.py tests/test_gpu_meta_warp.py tests/test_gpu_epilogue.py tests/test_gpu_host_out.py tests/test_gpu_cold_chain.py"
for tool in memcheck racecheck synccheck; do
    timeout 1500 compute-sanitizer --tool $tool python -m pytest $sel -m gpu -q --timeout 1400 \
        -k "golden or vectors_batched or rejects or mixed or constant_width or one_width or split_matches or epilogue or vectors_through or back_to_back_few or failed_frame or constant_pitch" \
        > gpurun_out/${T}_sanitizer_$tool.txt 2>&1
    grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|error" gpurun_out/${T}_sanitizer_$tool.txt | tail -4
done
echo done
