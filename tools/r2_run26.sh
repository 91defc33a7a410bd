#!/usr/bin/env bash
# round-2 GPU call 26: host-out pipeline with a small first chunk (chunks double up to 96 MB) against full-size chunks
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2n}
{
MCRAW_HOSTOUT_FIRST_MB=0 python tools/e2e_only.py host-out
MCRAW_HOSTOUT_FIRST_MB=4 python tools/e2e_only.py host-out
MCRAW_HOSTOUT_FIRST_MB=8 python tools/e2e_only.py host-out
MCRAW_HOSTOUT_FIRST_MB=16 python tools/e2e_only.py host-out
MCRAW_HOSTOUT_FIRST_MB=32 python tools/e2e_only.py host-out
python tools/e2e_only.py
} > gpurun_out/${T}_hostout_ab.txt 2>&1
cat gpurun_out/${T}_hostout_ab.txt
echo done
