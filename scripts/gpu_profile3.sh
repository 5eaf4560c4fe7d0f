#!/bin/bash
# ncu full captures of the fused sim kernel and the streaming NCE kernel
set -u
TAG=${1:-r01c}
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 900 $NCU --set full --import-source on -k regex:sim_fused_kernel -s 4 -c 2 -o gpurun_out/${TAG}_prof_sim \
   python scripts/prof_step.py 256 3 --one-stream > gpurun_out/${TAG}_ncu_sim.log 2>&1; echo rc=$?
timeout 900 $NCU --set full --import-source on -k regex:"nce_from_logits|sim_reduce_partials" -s 2 -c 2 -o gpurun_out/${TAG}_prof_nce \
   python scripts/prof_step.py 256 1 --hbm > gpurun_out/${TAG}_ncu_nce.log 2>&1; echo rc=$?
tail -n 4 gpurun_out/${TAG}_ncu_sim.log; tail -n 4 gpurun_out/${TAG}_ncu_nce.log
