#!/usr/bin/env bash
# round-2 GPU call 39: racecheck over the k_meta_split tests after the flag read was put behind its own barrier
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2z}
timeout 1500 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_meta_split.py tests/test_gpu_current.py -m gpu -q --timeout 1400 -k "split or rejects or mixed" \
    > gpurun_out/${T}_sanitizer_racecheck_split.txt 2>&1
grep -E "passed|failed|RACECHECK SUMMARY|Error" gpurun_out/${T}_sanitizer_racecheck_split.txt | tail -4
python tools/c2_steps.py --frames 1 --steps 400 --label c1ish | cut -c1-120
echo done
