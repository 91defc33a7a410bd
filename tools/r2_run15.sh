#!/usr/bin/env bash
# round-2 GPU call 15: k_meta_split CTA shapes (256 / 512 / 1024 threads) on C1; parity with the default
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2u}
timeout 900 python -m pytest tests/test_gpu_meta_split.py tests/test_gpu_current.py tests/test_golden.py tests/test_gpu_fuzz.py tests/test_gpu_dropin.py -m gpu -q -x --timeout 300 > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -4 gpurun_out/${T}_pytest_gpu.txt
for lib in libmcraw_b200.so libmcraw_b200_ks256.so libmcraw_b200_ks1024.so; do
  MCRAW_B200_LIB=$lib timeout 200 python bench.py --workload c1 --steps 200 --no-cpu-baseline 2>gpurun_out/${T}_c1.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('c1 $lib', round(d['ms_per_step'],4), 'verified', d['pixels_verified'], 'idx_ms', round(r['index_kernels_ms_per_launch'],4), 'main_ms', round(r['kernel_ms_per_launch'],4))"
done
MCRAW_META_SPLIT=0 timeout 200 python bench.py --workload c1 --steps 200 --no-cpu-baseline 2>gpurun_out/${T}_c1.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('c1 plain', round(d['ms_per_step'],4), 'idx_ms', round(r['index_kernels_ms_per_launch'],4))"
timeout 100 python tools/dropin_latency.py 2>&1 | tail -3
echo done
