#!/bin/bash
# ncu evidence of round 2: launch lists of one forward step and one training step, `--set full` captures of the kernels
# that are new this round.  Usage: gpurun --timeout 1500 -- bash scripts/gpu_profile_r02.sh [tag]
set -u
TAG=${1:-r02x}
mkdir -p gpurun_out
NCU="ncu --clock-control none"
echo "== launch list: forward step (2 eager steps at B=256, the second is the one to read)"
timeout 300 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/${TAG}_fwd_launches.csv \
   python scripts/prof_step.py 256 2 --one-stream > gpurun_out/${TAG}_fwd_list.log 2>&1; echo rc=$?; tail -2 gpurun_out/${TAG}_fwd_list.log
echo "== launch list: training step (warm-up + 1 step at B=256)"
timeout 400 $NCU --metrics gpu__time_duration.sum -c 1400 --csv --log-file gpurun_out/${TAG}_train_launches.csv \
   python scripts/train_profile.py 256 256 0 > gpurun_out/${TAG}_train_list.log 2>&1; echo rc=$?; tail -c 300 gpurun_out/${TAG}_train_list.log
full() {  # name regex count
  timeout 200 $NCU --kernel-name-base demangled --set full --import-source on -k regex:"$2" -c $3 -o gpurun_out/${TAG}_prof_$1 \
     python scripts/train_profile.py 64 256 0 > gpurun_out/${TAG}_ncu_$1.log 2>&1; echo "$1 rc=$?"
}
full attnbwd "attn_bwd_tc_kernel" 2
full gemmtn "umma_gemm_tn_kernel" 3
full simfused "sim_fused_kernel" 1
full simgrad "SimGradEpi" 1
full attnfwd "attention_kernel" 1
full lnbwd "layernorm_bwd_kernel" 1
full gemmlin "umma_gemm2_kernel<tanb::LinearEpi2<0>" 3
full resln "gemm_res_ln_kernel" 1
for f in gpurun_out/${TAG}_prof_*.ncu-rep; do python scripts/ncu_top.py $f 25 > ${f%.ncu-rep}_summary.txt 2>&1; done
python scripts/launch_summary.py gpurun_out/${TAG}_fwd_launches.csv > gpurun_out/${TAG}_fwd_launches_summary.txt 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_train_launches.csv > gpurun_out/${TAG}_train_launches_summary.txt 2>&1
rm -f gpurun_out/${TAG}_prof_*.ncu-rep
ls -la gpurun_out | grep ${TAG}_
