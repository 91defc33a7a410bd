#!/usr/bin/env bash
# round-2 GPU call 21: k_meta_warp with address-valued next pointers -- parity, chained / unchained / serial C2 steps
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2j}
timeout 900 python -m pytest tests/test_gpu_meta_warp.py tests/test_gpu_fuzz.py -m gpu -q --timeout 600 -x > gpurun_out/${T}_pytest.txt 2>&1; tail -5 gpurun_out/${T}_pytest.txt
{
python tools/c2_steps.py --label warp_chain24
MCRAW_CHAIN=0 python tools/c2_steps.py --label warp_nochain
MCRAW_CHAIN=0 MCRAW_META_WARP=0 python tools/c2_steps.py --label cta_nochain
MCRAW_NO_OVERLAP=1 python tools/c2_steps.py --label warp_serial
MCRAW_NO_OVERLAP=1 MCRAW_META_WARP=0 python tools/c2_steps.py --label cta_serial
} > gpurun_out/${T}_c2_ab.jsonl 2> gpurun_out/${T}_c2_ab.err
cut -c1-200 gpurun_out/${T}_c2_ab.jsonl
echo done
