#!/usr/bin/env bash
# round-2 GPU call 13: k_meta_split -- parity (split vs plain vs oracle), whole current-format suite, C1 / C3 timing with and without it
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2q}
timeout 900 python -m pytest tests/test_gpu_meta_split.py tests/test_gpu_current.py tests/test_golden.py tests/test_gpu_fuzz.py tests/test_gpu_dropin.py -m gpu -q -x --timeout 300 > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -15 gpurun_out/${T}_pytest_gpu.txt
for env in "" "MCRAW_META_SPLIT=0"; do
  for wl in c1; do
    env $env timeout 200 python bench.py --workload $wl --steps 200 --no-cpu-baseline 2>gpurun_out/${T}_$wl.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$wl [$env]', round(d['ms_per_step'],4), 'verified', d['pixels_verified'], 'idx_ms', round(r['index_kernels_ms_per_launch'],4), 'main_ms', round(r['kernel_ms_per_launch'],4))"
  done
done
echo done
MCRAW_B200_LIB=libmcraw_b200_ksdbg.so python bench.py --workload c1 --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | grep "^ks" | sort -k3,3 -k4,4n | awk '{k=$2 $3 $4 $5; if (!(k in seen)) {seen[k]=1; print}}' | head -24
