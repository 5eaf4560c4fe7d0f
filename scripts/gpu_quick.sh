#!/bin/bash
# quick regression + perf check after a kernel change
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -q -m gpu -x -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -n 15 gpurun_out/tests.log
if [ "${1:-}" = "trace" ]; then
  for s in "8192 512 512 0 1 f32" "8192 1536 512 0 0 bf16" "8192 2048 512 1 0 bf16" "8192 512 2048 0 1 f32"; do timeout 120 python scripts/gemm_trace.py $s; done > gpurun_out/trace.log 2>&1; cat gpurun_out/trace.log
fi
timeout 300 python scripts/kernel_bench.py > gpurun_out/kbench.log 2>&1; echo "kbench rc=$?"; cat gpurun_out/kbench.log
timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
