#!/bin/bash
# ncu full captures of the two GEMM instantiations (template arguments need the demangled name base).
set -u
TAG=${1:-r01b}
mkdir -p gpurun_out
NCU="ncu --clock-control none --kernel-name-base demangled"
timeout 900 $NCU --set full --import-source on -k regex:LinearEpi2 -s 102 -c 7 -o gpurun_out/${TAG}_prof_gemm \
   python scripts/prof_step.py 256 3 --one-stream > gpurun_out/${TAG}_ncu_gemm.log 2>&1; echo rc=$?
timeout 900 $NCU --set full --import-source on -k regex:SimEpi2 -s 4 -c 2 -o gpurun_out/${TAG}_prof_sim \
   python scripts/prof_step.py 256 3 --one-stream > gpurun_out/${TAG}_ncu_sim.log 2>&1; echo rc=$?
tail -4 gpurun_out/${TAG}_ncu_gemm.log gpurun_out/${TAG}_ncu_sim.log
