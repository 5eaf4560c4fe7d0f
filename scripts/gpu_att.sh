#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -q -m gpu -x -p no:cacheprovider -k "attention or parity or encoder or decoder" 2>&1 | tail -3
python scripts/ab_kernels.py attn | grep attention
timeout 300 python scripts/kernel_bench.py 2>&1 | grep -E "attention"
