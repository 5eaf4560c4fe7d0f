#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -p no:cacheprovider -k "attention" 2>&1 | tail -2
timeout 300 python scripts/kernel_bench.py 2>&1 | grep attention
timeout 600 ncu --clock-control none --set full --import-source on -k regex:attention_kernel -s 2 -c 1 -o gpurun_out/r01d_prof_attn python scripts/prof_attn.py 256 > gpurun_out/r01d_ncu_attn.log 2>&1; echo rc=$?
