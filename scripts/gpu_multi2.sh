#!/bin/bash
set -u
N=${1:-2}
mkdir -p gpurun_out
for mode in 0 1; do
  TAN_GRAPH_NCCL=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$mode bench.py --gpus $N --steps 20 --warmup 5 --skip-cpu --skip-hbm > gpurun_out/bench_g$mode.json 2> gpurun_out/bench_g$mode.err
  echo "graph_nccl=$mode rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_g$mode.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("n_gpus","value","ms_per_step","e2e","loss","loss_api","cuda_graph")})
except Exception as e:
    print("no json", e); print(open("gpurun_out/bench_g$mode.err").read()[-3000:])
PY
done
