#!/usr/bin/env bash
# round-2 GPU call 48: compute-sanitizer memcheck over the WHOLE GPU suite (no test selection; export / drop-in tests included)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2mc}
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 0 python -m pytest tests -m gpu -q --timeout 200 \
    --deselect tests/test_gpu_dropin.py::test_feed_cufile > gpurun_out/${T}_memcheck_all.txt 2>&1
grep -E "passed|failed|ERROR SUMMARY|Error|Invalid" gpurun_out/${T}_memcheck_all.txt | tail -6
echo done
