#!/usr/bin/env bash
# round-2 GPU call 17 (N GPUs of one box): H2D ceiling with all ranks copying at once, then the bench line at N
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-4}
T=${TAG:-r2x}
nvidia-smi topo -m > gpurun_out/${T}_topo.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/h2d_ceiling.py > gpurun_out/${T}_h2d_ceiling_n$N.json 2> gpurun_out/${T}_h2d_ceiling_n$N.err; cat gpurun_out/${T}_h2d_ceiling_n$N.json | cut -c1-600
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench_n$N.json").read().strip().splitlines()[-1])
print("c2 n", d["n_gpus"], "ms", round(d["ms_per_step"],4), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "h2d_gbs", round(d["e2e"]["h2d_gbs"],1), "peak", round(d["e2e"]["h2d_peak_gbs"],1), "frac", round(d["e2e"]["frac_of_h2d"],3), d["e2e"].get("placement"))
print("host_out", round(d["e2e_host_out"]["value"]))
for k,v in d["workloads"].items(): print(k, round(v["ms_per_step"],4), round(v["value"]), v.get("pixels_verified"), (v.get("e2e") or {}).get("value"))
PY
tail -3 gpurun_out/${T}_bench_n$N.err
echo done
