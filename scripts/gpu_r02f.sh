#!/bin/bash
# 2-GPU call: text-embedder tests (GPU 0), sharded == single-GPU loss check, bench at N=2 with / without graph-captured NCCL
set -u
N=${1:-2}
mkdir -p gpurun_out
echo "== text embedder + all gpu tests"; timeout 600 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -6
echo "== multigpu loss check"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/multigpu_check.py 2>&1 | grep -E "world=|Error|error" | cut -c1-260
echo "== multigpu train check"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 scripts/multigpu_train_check.py 2>&1 | tail -4 | cut -c1-300
for flags in "TAN_GRAPH_NCCL=0" "TAN_GRAPH_NCCL=1" "TAN_GRAPH_NCCL=0 TAN_GATHER_COALESCED=0"; do
  tag=$(echo "$flags" | tr ' =' '__')
  env $flags timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --skip-hbm > gpurun_out/r02f_n${N}_${tag}.json 2> gpurun_out/r02f_n${N}_${tag}.err
  echo "bench n=$N $flags rc=$?"; python - gpurun_out/r02f_n${N}_${tag} <<'PY'
import json, sys
try:
    d=json.loads(open(sys.argv[1] + ".json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("n_gpus","value","ms_per_step","e2e","loss","loss_api","cuda_graph","kernel_ms_per_step","kernel_ms_sum","comm_ms_per_step","gpu_launches_per_step")})
except Exception as e:
    print("no json", e); print(open(sys.argv[1] + ".err").read()[-2500:])
PY
done
