#!/bin/bash
# usage: gpurun --gpus N -- bash scripts/gpu_scale.sh N   (bench at N GPUs as the driver launches it)
set -u
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench n=$N rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("n_gpus","value","ms_per_step","e2e","loss","loss_api","kernel_ms_per_step","gpu_launches_per_step","clocks")})
except Exception as e:
    print("no json", e); print(open("gpurun_out/bench_n$N.err").read()[-2500:])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | cut -c1-200
