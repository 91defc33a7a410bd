#!/usr/bin/env bash
# round-2 GPU call 4: warp-per-tile legacy kernel (two segment sizes), feeds, file leg
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest_gpu.txt 2>&1; tail -12 gpurun_out/r2d_pytest_gpu.txt
grep "feed\[" gpurun_out/r2d_pytest_gpu.txt
for lib in libmcraw_b200.so libmcraw_b200_seg256.so; do
for c in 0 16 24; do
  if [ $c = 0 ]; then lab=default; else export MCRAW_LGW_CTAS_PER_SM=$c; lab=$c; fi
  MCRAW_B200_LIB=$lib timeout 300 python bench.py --workload c4 --steps 30 --no-cpu-baseline 2>gpurun_out/r2d_c4_$lab.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4 $lib ctas/sm $lab', round(d['ms_per_step'],4), 'verified', d['pixels_verified'], 'frac', round(d['roofline']['whole_step']['frac'],3))"
  unset MCRAW_LGW_CTAS_PER_SM
done
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_legacy_warp -s 6 -c 1 -f -o gpurun_out/r2d_ncu_k_legacy_warp \
    python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_ncu_k_legacy_warp.log 2>&1
python -s -m pytest tests/test_gpu_dropin.py -m gpu -q -s -k feeds 2>&1 | grep "feed\["
timeout 900 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/r2d_bench_all.json 2> gpurun_out/r2d_bench_all.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2d_bench_all.json').read())
print('c2', d['ms_per_step'], d['pixels_verified'])
for k,v in d['workloads'].items(): print(k, round(v['ms_per_step'],4), round(v['value']), v.get('pixels_verified'), v.get('feed'))
PY
tail -3 gpurun_out/r2d_bench_all.err
echo done
