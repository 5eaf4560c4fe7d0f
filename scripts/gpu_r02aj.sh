#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 ncu --clock-control none --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r02aj_fwd_launches.csv python scripts/prof_step.py 256 2 --one-stream > gpurun_out/r02aj_fwd_list.log 2>&1; echo rc=$?
python scripts/launch_summary.py gpurun_out/r02aj_fwd_launches.csv "ncu launch list, 2 eager forward+loss steps at B=256 (scripts/prof_step.py 256 2 --one-stream), final tree of round 2 (r02aj)" > gpurun_out/r02aj_fwd_launches_summary.txt
head -16 gpurun_out/r02aj_fwd_launches_summary.txt | cut -c1-170
