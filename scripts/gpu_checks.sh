#!/bin/bash
# Runs the GPU test groups in separate processes (a trapped kernel kills only its own group) and
# collects logs under gpurun_out/.  Usage: gpurun -- bash scripts/gpu_checks.sh [extra pytest args]
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() {  # name, timeout, command...
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1
  local rc=$?
  echo "rc=$rc" | tee -a gpurun_out/summary.txt
  tail -n 25 "gpurun_out/$name.log" | tee -a gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
run t_cast_ln 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "cast or layernorm or device" -p no:cacheprovider
run t_linear 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "linear" -p no:cacheprovider
run t_attn 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" -p no:cacheprovider
run t_sim 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "sim or nce" -p no:cacheprovider
run t_parity 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider
run t_smoke 300 python __graft_entry__.py smoke
run kbench 300 python scripts/kernel_bench.py
