#!/usr/bin/env bash
# round-2 GPU call 16: k_meta_split window sizes on C1 (16 / 8 / 4 KiB), parity of each with the split tests
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2w}
for lib in libmcraw_b200.so libmcraw_b200_ks8192_512.so libmcraw_b200_ks8192_256.so libmcraw_b200_ks4096_256.so; do
  MCRAW_B200_LIB=$lib timeout 300 python -m pytest tests/test_gpu_meta_split.py -m gpu -q -x --timeout 200 2>&1 | tail -1
  MCRAW_B200_LIB=$lib timeout 200 python bench.py --workload c1 --steps 200 --no-cpu-baseline 2>gpurun_out/${T}_c1.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('c1 $lib', round(d['ms_per_step'],4), 'verified', d['pixels_verified'], 'idx_ms', round(r['index_kernels_ms_per_launch'],4), 'main_ms', round(r['kernel_ms_per_launch'],4))"
done
echo done
