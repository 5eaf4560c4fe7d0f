#!/bin/bash
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/multigpu_check.py 2>&1 | grep -E "world=|Error|error" | cut -c1-160
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 scripts/multigpu_train_check.py 2>&1 | tail -1 | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --skip-hbm --train-multi > gpurun_out/r02s_n$N.json 2> gpurun_out/r02s_n$N.err
python - gpurun_out/r02s_n$N <<'PY'
import json, sys
try:
    d=json.loads(open(sys.argv[1] + ".json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("n_gpus","value","ms_per_step","loss","loss_api","kernel_ms_sum","comm_ms_per_step")}, "e2e", d["e2e"]["value"], "train", d.get("train_step"))
except Exception as e:
    print("no json", e); print(open(sys.argv[1] + ".err").read()[-2500:])
PY
# weak-scaling style small batch: 32 clips per GPU (the N=8 per-rank shape) to see the host-bound e2e path
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 --skip-hbm --skip-train --scaling weak > gpurun_out/r02s_weak_n$N.json 2> gpurun_out/r02s_weak_n$N.err
python - gpurun_out/r02s_weak_n$N <<'PY'
import json, sys
d=json.loads(open(sys.argv[1] + ".json").read().strip().splitlines()[-1])
print("weak", {k:d.get(k) for k in ("n_gpus","value","ms_per_step","loss")}, "e2e", d["e2e"]["value"])
PY
