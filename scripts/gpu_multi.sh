#!/bin/bash
# usage: gpurun --gpus N -- bash scripts/gpu_multi.sh N   (sharded == single-GPU loss, then bench at N)
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/multigpu_check.py 2>&1 | grep -E "world=|Error|error" | cut -c1-260
bash scripts/gpu_scale.sh $N
