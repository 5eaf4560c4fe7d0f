#!/bin/bash
# Round-end validation in one GPU call: the full GPU test suite, the bench line (incl. the train_step leg) and the
# ncu launch list of one training step.  Usage: gpurun -- bash scripts/gpu_final.sh [tag]
set -u
TAG=${1:-r01g}
mkdir -p gpurun_out
echo "== tests"; timeout 150 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1; echo rc=$?; tail -3 gpurun_out/${TAG}_tests.log
echo "== bench"; timeout 150 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo rc=$?
tail -c 900 gpurun_out/${TAG}_bench.json; tail -2 gpurun_out/${TAG}_bench.err
echo "== ncu launch list of one training step"
timeout 80 ncu --clock-control none --metrics gpu__time_duration.sum -s 1500 -c 1500 --csv --log-file gpurun_out/${TAG}_train_launches.csv \
   python scripts/train_profile.py 256 256 0 > gpurun_out/${TAG}_train_list.log 2>&1; echo rc=$?; tail -c 300 gpurun_out/${TAG}_train_list.log
