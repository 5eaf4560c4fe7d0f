#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== align tests"; timeout 600 python -m pytest tests/test_align_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -8
echo "== whole gpu suite"; timeout 1200 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -3
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
