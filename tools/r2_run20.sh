#!/usr/bin/env bash
# round-2 GPU call 20: k_meta_warp (one warp per (frame, stream)) -- parity, then the C2 step against k_meta<K1Batch> and hold-back sizes
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2i}
timeout 900 python -m pytest tests/test_gpu_meta_warp.py tests/test_gpu_current.py tests/test_gpu_fuzz.py tests/test_gpu_meta_split.py -m gpu -q --timeout 600 -x > gpurun_out/${T}_pytest.txt 2>&1; tail -15 gpurun_out/${T}_pytest.txt
{
python tools/c2_steps.py --label warp_chain24
MCRAW_META_WARP=0 python tools/c2_steps.py --label cta_chain24
MCRAW_CHAIN=8 python tools/c2_steps.py --label warp_chain8
MCRAW_CHAIN=12 python tools/c2_steps.py --label warp_chain12
MCRAW_CHAIN=16 python tools/c2_steps.py --label warp_chain16
MCRAW_CHAIN=32 python tools/c2_steps.py --label warp_chain32
MCRAW_CHAIN=0 python tools/c2_steps.py --label warp_nochain
MCRAW_CHAIN=0 MCRAW_META_WARP=0 python tools/c2_steps.py --label cta_nochain
} > gpurun_out/${T}_c2_ab.jsonl 2> gpurun_out/${T}_c2_ab.err
cut -c1-200 gpurun_out/${T}_c2_ab.jsonl
echo done
