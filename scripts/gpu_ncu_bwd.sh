#!/bin/bash
# ncu --set full captures of the two backward kernels with the largest share (one launch each), small batch.
set -u
TAG=${1:-r01g}
mkdir -p gpurun_out
timeout 70 ncu --clock-control none --set full --import-source on -k regex:"SimGradEpi" -c 1 -o gpurun_out/${TAG}_prof_simgrad \
   python scripts/train_profile.py 64 256 0 > gpurun_out/${TAG}_ncu_simgrad.log 2>&1; echo rc=$?
timeout 70 ncu --clock-control none --set full --import-source on -k regex:"attn_bwd_dkv_mma|attn_bwd_dq_mma" -c 2 -o gpurun_out/${TAG}_prof_attnbwd \
   python scripts/train_profile.py 64 256 0 > gpurun_out/${TAG}_ncu_attnbwd.log 2>&1; echo rc=$?
ls -la gpurun_out/${TAG}_prof_*.ncu-rep
