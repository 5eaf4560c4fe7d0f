#!/bin/bash
# final 8-GPU call of the round: sharded == single-GPU loss at world 8, strong scaling N=8 (+ NCCL-in-graph variant,
# + training step) and the N=1 point ON THE SAME BOX.  usage: gpurun --gpus 8 --timeout 900 -- bash scripts/gpu_scale8_final.sh [tag]
set -u
TAG=${1:-r02y}
mkdir -p gpurun_out
run() {   # name nproc extra-env bench-args...
  local name=$1 np=$2 envs=$3; shift 3
  env $envs timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29520 \
    bench.py --gpus $np "$@" > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  echo "== $name rc=$?"
  python - gpurun_out/${TAG}_${name} <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1] + ".json").read().strip().splitlines()[-1])
    keys = ("n_gpus", "scaling", "value", "ms_per_step", "loss", "loss_api", "cuda_graph", "kernel_ms_per_step", "kernel_ms_sum", "comm_ms_per_step", "gpu_launches_per_step")
    print({k: d.get(k) for k in keys}, "e2e", d["e2e"]["value"], "train", (d.get("train_step") or {}).get("ms_per_step"), (d.get("train_step") or {}).get("value"))
except Exception as e:
    print("no json", e); print(open(sys.argv[1] + ".err").read()[-1500:])
PY
}
echo "== multigpu loss check (8)"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/multigpu_check.py 2>&1 | grep -E "world=|Error|error" | cut -c1-200
run n8 8 TAN_X=0 --steps 20 --warmup 5 --skip-hbm --train-multi
run n8_graphnccl 8 TAN_GRAPH_NCCL=1 --steps 20 --warmup 5 --skip-hbm --skip-train
env CUDA_VISIBLE_DEVICES=0 timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --skip-hbm --skip-train --skip-eager --skip-cpu > gpurun_out/${TAG}_n1.json 2> gpurun_out/${TAG}_n1.err
python - gpurun_out/${TAG}_n1 <<'PY'
import json, sys
d = json.loads(open(sys.argv[1] + ".json").read().strip().splitlines()[-1])
print("== n1 (same box)", {k: d.get(k) for k in ("n_gpus", "scaling", "value", "ms_per_step", "loss", "kernel_ms_per_step")}, "e2e", d["e2e"]["value"])
PY
