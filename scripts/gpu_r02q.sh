#!/bin/bash
set -u
mkdir -p gpurun_out
TAN_ATT_FLAGS=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:attention_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/r02q_attn2 python scripts/attn_time.py 256 8 256 > gpurun_out/r02q_ncu2.log 2>&1
tail -2 gpurun_out/r02q_ncu2.log
python scripts/ncu_top.py gpurun_out/r02q_attn2.ncu-rep 30 > gpurun_out/r02q_attn2_summary.txt 2>&1
head -40 gpurun_out/r02q_attn2_summary.txt
