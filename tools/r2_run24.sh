#!/usr/bin/env bash
# round-2 GPU call 24: HBM traffic-mix microbenchmark (what a 1 : 2 read : write stream can reach next to the copy peak)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/hbm_mix tools/ubench/hbm_mix.cu && timeout 300 gpurun_out/hbm_mix | tee gpurun_out/r2_ubench_hbm_mix.txt
rm -f gpurun_out/hbm_mix
echo done
