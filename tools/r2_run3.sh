#!/usr/bin/env bash
# round-2 GPU call 3: parity suite, warp-specialised legacy kernel, chain default, bulk staging with chain, sanitizers
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest_gpu.txt 2>&1; tail -12 gpurun_out/r2c_pytest_gpu.txt
for c in 0 3 4 5; do
  if [ $c = 0 ]; then lab=default; else export MCRAW_LGF_CTAS_PER_SM=$c; lab=$c; fi
  timeout 300 python bench.py --workload c4 --steps 30 --no-cpu-baseline 2>gpurun_out/r2c_c4_$lab.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4 ctas/sm $lab', round(d['ms_per_step'],4), 'verified', d['pixels_verified'], 'frac', round(d['roofline']['whole_step']['frac'],3))"
  unset MCRAW_LGF_CTAS_PER_SM
done
{
python tools/c2_steps.py --label chain_default
MCRAW_CHAIN=0 python tools/c2_steps.py --label chain_off
MCRAW_B200_LIB=libmcraw_b200_bulk.so python tools/c2_steps.py --label bulk_chain
MCRAW_B200_LIB=libmcraw_b200_bulk.so MCRAW_CHAIN=0 python tools/c2_steps.py --label bulk_nochain
MCRAW_B200_LIB=libmcraw_b200_bulk.so MCRAW_CHAIN=32 python tools/c2_steps.py --label bulk_chain32
} > gpurun_out/r2c_c2_ab.jsonl 2> gpurun_out/r2c_c2_ab.err
cat gpurun_out/r2c_c2_ab.jsonl; tail -3 gpurun_out/r2c_c2_ab.err
timeout 900 python bench.py --steps 20 > gpurun_out/r2c_bench_all.json 2> gpurun_out/r2c_bench_all.err; cut -c1-400 gpurun_out/r2c_bench_all.json; tail -3 gpurun_out/r2c_bench_all.err
MCRAW_B200_LIB=libmcraw_b200_bulk.so timeout 900 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/r2c_bench_all_bulk.json 2> gpurun_out/r2c_bench_all_bulk.err; cut -c1-400 gpurun_out/r2c_bench_all_bulk.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_legacy_fused -s 6 -c 1 -f -o gpurun_out/r2c_ncu_k_legacy_fused \
    python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_ncu_k_legacy_fused.log 2>&1
sel="tests/test_golden.py tests/test_gpu_current.py tests/test_gpu_legacy.py tests/test_gpu_fuzz.py"
for tool in memcheck racecheck; do
    timeout 1500 compute-sanitizer --tool $tool python -m pytest $sel -m gpu -q -k "golden or vectors_batched or rejects or mixed or fuzz or encoded_width" \
        > gpurun_out/r2c_sanitizer_$tool.txt 2>&1
    tail -3 gpurun_out/r2c_sanitizer_$tool.txt
done
timeout 900 compute-sanitizer --tool racecheck python tools/c2_steps.py --frames 4 --steps 8 --label racecheck_chain > gpurun_out/r2c_sanitizer_racecheck_chain.txt 2>&1
tail -3 gpurun_out/r2c_sanitizer_racecheck_chain.txt
echo done
