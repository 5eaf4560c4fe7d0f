#!/bin/bash
set -u
mkdir -p gpurun_out
for v in _old _oldw ""; do echo "== timings variant '$v'"; TAN_LIB_PATH=$PWD/temporalalignnet_b200/libtan_b200$v.so timeout 200 python scripts/attn_time.py 256 8 256 256 8 288 32 8 256 32 8 288 32 12 1024 32 12 1152 128 8 512 128 8 576 64 8 64 64 8 72 2>&1 | tail -10; done
echo "== attention kernel tests (old kernel, plain try_wait)"; TAN_LIB_PATH=$PWD/temporalalignnet_b200/libtan_b200_oldw.so timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_backward_kernels_gpu.py -q -x -p no:cacheprovider -k "attention" 2>&1 | tail -3
