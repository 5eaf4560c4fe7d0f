#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "attention" -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -n 25 gpurun_out/tests.log
timeout 300 python scripts/kernel_bench.py attn > gpurun_out/kbench.log 2>&1; echo "kbench rc=$?"; cat gpurun_out/kbench.log
