#!/usr/bin/env bash
# round-2 GPU call 34: the reference-shaped single-frame call (raw::Decode through the drop-in library) against the number of host copy threads
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2v}
nproc
for p in 4 6 8 12 16; do echo "MCRAW_HOST_PARTS=$p"; MCRAW_HOST_PARTS=$p python tools/dropin_latency.py 2>&1 | head -1; done | tee gpurun_out/${T}_dropin_parts.txt
python tools/dropin_latency.py 2>&1 | tail -1 | tee -a gpurun_out/${T}_dropin_parts.txt
echo done
