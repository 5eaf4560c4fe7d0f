#!/bin/bash
# compute-sanitizer over one small launch of every kernel family (scripts/sanitize_driver.py): memcheck (out-of-bounds /
# misaligned accesses, also through TMA descriptors), racecheck (shared-memory hazards between the warp roles of the
# mbarrier pipelines) and synccheck (barrier misuse).  Usage: gpurun --timeout 1500 -- bash scripts/gpu_sanitize.sh [tag]
set -u
TAG=${1:-r02}
TOOLS=${2:-"memcheck racecheck synccheck"}   # e.g. bash scripts/gpu_sanitize.sh r02am memcheck
TMO=${3:-600}
mkdir -p gpurun_out
python scripts/sanitize_driver.py > gpurun_out/${TAG}_sanitize_plain.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/${TAG}_sanitize_plain.log
for tool in $TOOLS; do
  timeout $TMO compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_driver.py > gpurun_out/${TAG}_sanitize_${tool}.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|sanitize driver done" gpurun_out/${TAG}_sanitize_${tool}.log | head -12
done
