#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 ncu --clock-control none --kernel-name-base demangled --set full --import-source on -k regex:"LinearEpi2<.int.0, .int.[01]>" --launch-skip 2 -c 2 -f -o gpurun_out/r02ac_gemmlin python scripts/prof_step.py 256 1 --one-stream > gpurun_out/r02ac_ncu.log 2>&1; echo rc=$?; tail -2 gpurun_out/r02ac_ncu.log
python scripts/ncu_top.py gpurun_out/r02ac_gemmlin.ncu-rep 40 > gpurun_out/r02ac_prof_gemmlin_summary.txt 2>&1
head -75 gpurun_out/r02ac_prof_gemmlin_summary.txt | cut -c1-200
