#!/usr/bin/env bash
# round-2 GPU call 6: every step under its own timeout.  Parity suite; legacy v5 CTA shapes; epilogue cost; bench line.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 150 > gpurun_out/r2f_pytest_gpu.txt 2>&1; tail -6 gpurun_out/r2f_pytest_gpu.txt
for lib in libmcraw_b200.so libmcraw_b200_lgw2.so libmcraw_b200_lgw8.so; do
  MCRAW_B200_LIB=$lib timeout 120 python bench.py --workload c4 --steps 30 --no-cpu-baseline 2>gpurun_out/r2f_c4.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4 $lib', round(d['ms_per_step'],4), 'verified', d['pixels_verified'], 'frac', round(d['roofline']['whole_step']['frac'],3))"
done
{
timeout 100 python tools/c2_steps.py --label raw
timeout 100 python tools/c2_steps.py --label black_sub_u16 --levels 1
timeout 100 python tools/c2_steps.py --label norm_f16 --levels 2
MCRAW_B200_LIB=libmcraw_b200_ldgsts.so timeout 100 python tools/c2_steps.py --label ldgsts_staging
MCRAW_CHAIN=0 timeout 100 python tools/c2_steps.py --label chain_off
} > gpurun_out/r2f_c2_ab.jsonl 2> gpurun_out/r2f_c2_ab.err
cat gpurun_out/r2f_c2_ab.jsonl; tail -2 gpurun_out/r2f_c2_ab.err
timeout 400 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/r2f_bench_all.json 2> gpurun_out/r2f_bench_all.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2f_bench_all.json').read())
    print('c2', round(d['ms_per_step'],4), d['pixels_verified'], 'whole', round(d['roofline']['whole_step']['frac'],3), 'k', round(d['roofline']['frac'],3))
    for k,v in d['workloads'].items(): print(k, round(v['ms_per_step'],4), round(v['value']), v.get('pixels_verified'), v.get('feed'))
except Exception as e: print('bench all failed', e)
PY
tail -3 gpurun_out/r2f_bench_all.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_legacy_warp -s 6 -c 1 -f -o gpurun_out/r2f_ncu_k_legacy_warp \
    python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_ncu_k_legacy_warp.log 2>&1
echo done
