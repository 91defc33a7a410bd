#!/usr/bin/env bash
# round-2 GPU call 28: plan slots found by content (8 slots) -- whole GPU suite, host-out with first chunks of 0 / 8 / 16 / 32 MB, C2 / C1 steps
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2p}
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.txt
{
MCRAW_HOSTOUT_FIRST_MB=0 python tools/e2e_only.py host-out
MCRAW_HOSTOUT_FIRST_MB=8 python tools/e2e_only.py host-out
MCRAW_HOSTOUT_FIRST_MB=16 python tools/e2e_only.py host-out
MCRAW_HOSTOUT_FIRST_MB=32 python tools/e2e_only.py host-out
python tools/e2e_only.py
} > gpurun_out/${T}_hostout_ab.txt 2>&1
cat gpurun_out/${T}_hostout_ab.txt
python tools/c2_steps.py --label c2_8slots | cut -c1-150
timeout 300 python bench.py --workload c1 --no-cpu-baseline > gpurun_out/${T}_bench_c1.json 2> gpurun_out/${T}_bench_c1.err; python tools/bench_summary.py gpurun_out/${T}_bench_c1.json
echo done
