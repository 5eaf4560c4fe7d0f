#!/bin/bash
set -u
mkdir -p gpurun_out
for v in _base "" _base ""; do echo "== timings variant '$v'"; TAN_LIB_PATH=$PWD/temporalalignnet_b200/libtan_b200$v.so timeout 200 python scripts/attn_time.py 256 8 256 256 8 288 32 12 1152 128 8 576 2>&1 | tail -4; done
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -p no:cacheprovider -k "attention" 2>&1 | tail -2
