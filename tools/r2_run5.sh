#!/usr/bin/env bash
# round-2 GPU call 5: every step under its own timeout.  Parity suite; legacy v4 shapes; k_meta shapes; epilogue cost.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 120 > gpurun_out/r2e_pytest_gpu.txt 2>&1; tail -8 gpurun_out/r2e_pytest_gpu.txt
timeout 200 python -m pytest tests/test_gpu_dropin.py -m gpu -q -s -k "feed" --timeout 150 2>&1 | grep "feed\[" 
for lib in libmcraw_b200.so libmcraw_b200_lgw1x256.so; do
for c in 0 8; do
  if [ $c = 0 ]; then lab=default; else export MCRAW_LGW_CTAS_PER_SM=$c; lab=$c; fi
  MCRAW_B200_LIB=$lib timeout 120 python bench.py --workload c4 --steps 30 --no-cpu-baseline 2>gpurun_out/r2e_c4_$lab.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4 $lib ctas/sm $lab', round(d['ms_per_step'],4), 'verified', d['pixels_verified'], 'frac', round(d['roofline']['whole_step']['frac'],3))"
  unset MCRAW_LGW_CTAS_PER_SM
done
done
timeout 120 python bench.py --workload c1 --steps 40 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c1 few-shape', round(d['ms_per_step'],4), d['pixels_verified'], 'idx', round(d['roofline']['index_kernels_ms_per_launch'],4))"
MCRAW_META_SMALL=1 timeout 120 python bench.py --workload c1 --steps 40 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c1 batch-shape', round(d['ms_per_step'],4), d['pixels_verified'], 'idx', round(d['roofline']['index_kernels_ms_per_launch'],4))"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_legacy_warp -s 6 -c 1 -f -o gpurun_out/r2e_ncu_k_legacy_warp \
    python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_ncu_k_legacy_warp.log 2>&1
timeout 400 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/r2e_bench_all.json 2> gpurun_out/r2e_bench_all.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2e_bench_all.json').read())
    print('c2', round(d['ms_per_step'],4), d['pixels_verified'], 'whole', round(d['roofline']['whole_step']['frac'],3))
    for k,v in d['workloads'].items(): print(k, round(v['ms_per_step'],4), round(v['value']), v.get('pixels_verified'), v.get('feed'))
except Exception as e: print('bench all failed', e)
PY
tail -3 gpurun_out/r2e_bench_all.err
echo done
