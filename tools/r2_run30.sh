#!/usr/bin/env bash
# round-2 GPU call 30: the tree as it stands -- GPU suite, both bench arms (1 GPU)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2r}
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/${T}_bench_all.json 2> gpurun_out/${T}_bench_all.err; python tools/bench_summary.py gpurun_out/${T}_bench_all.json
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; tail -c 200 gpurun_out/${T}_bench_reference.json
echo done
