#!/usr/bin/env bash
# round-2 GPU call 27: host-out D2H as 2-D copies -- parity tests, then the pipeline timed with different first-chunk sizes
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2o}
timeout 600 python -m pytest tests/test_gpu_host_out.py -m gpu -q --timeout 300 > gpurun_out/${T}_pytest.txt 2>&1; tail -8 gpurun_out/${T}_pytest.txt
{
MCRAW_HOSTOUT_FIRST_MB=0 python tools/e2e_only.py host-out
MCRAW_HOSTOUT_FIRST_MB=8 python tools/e2e_only.py host-out
MCRAW_HOSTOUT_FIRST_MB=32 python tools/e2e_only.py host-out
MCRAW_HOSTOUT_FIRST_MB=48 python tools/e2e_only.py host-out
} > gpurun_out/${T}_hostout_ab.txt 2>&1
cat gpurun_out/${T}_hostout_ab.txt
echo done
