#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --skip-hbm --train-multi > gpurun_out/r02z_n2.json 2> gpurun_out/r02z_n2.err
echo "rc=$?"
python - gpurun_out/r02z_n2 <<'PY'
import json, sys
try:
    d=json.loads(open(sys.argv[1] + ".json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("n_gpus","value","ms_per_step","loss","loss_api","cuda_graph","kernel_ms_sum","comm_ms_per_step")}, "e2e", d["e2e"]["value"], "train", (d.get("train_step") or {}).get("ms_per_step"))
except Exception as e:
    print("no json", e)
print(open(sys.argv[1] + ".err").read()[-800:])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 2>&1 | tail -2 | cut -c1-400
