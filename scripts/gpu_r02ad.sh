#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 ncu --clock-control none --kernel-name-base demangled --set full --import-source on -k regex:"gemm_res_ln_kernel|LinearEpi2<.int.1" --launch-skip 3 -c 3 -f -o gpurun_out/r02ad_resln python scripts/prof_step.py 256 1 --one-stream > gpurun_out/r02ad_ncu.log 2>&1; echo rc=$?; tail -2 gpurun_out/r02ad_ncu.log
python scripts/ncu_top.py gpurun_out/r02ad_resln.ncu-rep 12 > gpurun_out/r02ad_prof_resln_cproj_summary.txt 2>&1
grep -E "kernel:|time_duration|tensor_cycles|stall reasons|registers|dram__bytes|dram_throughput" gpurun_out/r02ad_prof_resln_cproj_summary.txt | cut -c1-330
