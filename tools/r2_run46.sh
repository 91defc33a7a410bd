#!/usr/bin/env bash
# round-2 GPU call 46: legacy kernels of back-to-back batches chained by programmatic launches -- parity, racecheck, C4 step with / without
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2l6}
timeout 900 python -m pytest tests/test_gpu_legacy.py tests/test_gpu_fuzz.py -m gpu -q --timeout 600 > gpurun_out/${T}_pytest.txt 2>&1; tail -4 gpurun_out/${T}_pytest.txt
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_legacy.py -m gpu -q --timeout 800 -k "back_to_back" > gpurun_out/${T}_racecheck.txt 2>&1
grep -E "passed|failed|RACECHECK SUMMARY|Error" gpurun_out/${T}_racecheck.txt | tail -3
for v in "" "MCRAW_CHAIN=0"; do
    env $v timeout 300 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/${T}_c4.json 2> gpurun_out/${T}_c4.err
    python -c "
import json;d=json.load(open('gpurun_out/${T}_c4.json'));print('c4 [$v]: step', round(d['ms_per_step'],4), 'ms, whole', round(d['roofline']['whole_step']['frac'],4), 'kernel', round(d['roofline']['kernel_ms_per_launch'],4), d['pixels_verified'])"
done | tee gpurun_out/${T}_c4_chain.txt
echo done
