#!/usr/bin/env bash
# round-2 GPU call 2: parity suite, warp-specialised legacy kernel, chain / completion-word / bulk-staging A/B
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest_gpu.txt 2>&1; tail -12 gpurun_out/r2b_pytest_gpu.txt
for c in 0 3 4 5; do
  if [ $c = 0 ]; then lab=default; else export MCRAW_LGF_CTAS_PER_SM=$c; lab=$c; fi
  timeout 300 python bench.py --workload c4 --steps 30 --no-cpu-baseline 2>gpurun_out/r2b_c4_$lab.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4 ctas/sm $lab', round(d['ms_per_step'],4), 'verified', d['pixels_verified'], 'frac', round(d['roofline']['whole_step']['frac'],3))"
  unset MCRAW_LGF_CTAS_PER_SM
done
{
python tools/c2_steps.py --label base
MCRAW_CHAIN=24 python tools/c2_steps.py --label chain24_event
MCRAW_CHAIN=24 MCRAW_DONE_FLAG=1 python tools/c2_steps.py --label chain24_flag
MCRAW_CHAIN=16 MCRAW_DONE_FLAG=1 python tools/c2_steps.py --label chain16_flag
MCRAW_CHAIN=40 MCRAW_DONE_FLAG=1 python tools/c2_steps.py --label chain40_flag
MCRAW_DONE_FLAG=1 python tools/c2_steps.py --label flag_only
MCRAW_B200_LIB=libmcraw_b200_bulk.so python tools/c2_steps.py --label bulk_staging
} > gpurun_out/r2b_c2_ab.jsonl 2> gpurun_out/r2b_c2_ab.err
cat gpurun_out/r2b_c2_ab.jsonl; tail -3 gpurun_out/r2b_c2_ab.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_legacy_fused -s 6 -c 1 -f -o gpurun_out/r2b_ncu_k_legacy_fused \
    python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_ncu_k_legacy_fused.log 2>&1
echo done
