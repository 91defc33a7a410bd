#!/usr/bin/env bash
# round-2 GPU call 33: the bench line under torchrun on the GPUs of the box (both arms), as the driver launches it
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-2}
T=${TAG:-r2u}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 \
    > gpurun_out/${T}_bench_all_n$N.json 2> gpurun_out/${T}_bench_all_n$N.err
python tools/bench_summary.py gpurun_out/${T}_bench_all_n$N.json || tail -20 gpurun_out/${T}_bench_all_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 100 --warmup 5 \
    > gpurun_out/${T}_bench_reference_n$N.json 2> gpurun_out/${T}_bench_reference_n$N.err
tail -c 300 gpurun_out/${T}_bench_reference_n$N.json
echo done
