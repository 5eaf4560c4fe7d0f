#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "linear" -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -n 5 gpurun_out/tests.log
for s in "8192 512 512 0 1 f32" "8192 1536 512 0 0 bf16" "9216 2048 512 1 0 bf16" "8192 512 2048 0 1 f32"; do timeout 120 python scripts/gemm_trace.py $s; done > gpurun_out/trace.log 2>&1; cat gpurun_out/trace.log
timeout 300 python scripts/kernel_bench.py linear > gpurun_out/kbench.log 2>&1; echo "kbench rc=$?"; cat gpurun_out/kbench.log
