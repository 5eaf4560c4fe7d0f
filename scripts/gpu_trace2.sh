#!/bin/bash
set -u
mkdir -p gpurun_out
for s in "8192 512 512 0 1 f32" "8192 1536 512 0 0 bf16" "9216 2048 512 1 0 bf16"; do timeout 120 python scripts/gemm_trace.py $s; done > gpurun_out/trace.log 2>&1; cat gpurun_out/trace.log
