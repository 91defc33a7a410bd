#!/usr/bin/env bash
# round-2 GPU call 43: head of the tree -- smoke(), the GPU suite, one bench line
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r2head}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/${T}_pytest_gpu.txt 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/${T}_bench_all.json 2> gpurun_out/${T}_bench_all.err; python tools/bench_summary.py gpurun_out/${T}_bench_all.json
echo done
