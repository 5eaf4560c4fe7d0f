#!/bin/bash
# ncu evidence for one round: bench line, launch list of one eager step, one full capture per kernel class.
# Usage: gpurun -- bash scripts/gpu_profile.sh [tag]
set -u
TAG=${1:-r01e}
mkdir -p gpurun_out
NCU="ncu --clock-control none"
echo "== tests"; timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -2
echo "== bench (full line incl. cpu_baseline)"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo rc=$?
tail -c 400 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>/dev/null; echo rc=$?; cut -c1-300 gpurun_out/${TAG}_bench_ref.json
echo "== ncu launch list (3rd eager step, one stream so that the order is the program order)"
timeout 900 $NCU --metrics gpu__time_duration.sum -s 230 -c 130 --csv --log-file gpurun_out/${TAG}_launches.csv \
   python scripts/prof_step.py 256 3 --one-stream > gpurun_out/${TAG}_list.log 2>&1; echo rc=$?; tail -2 gpurun_out/${TAG}_list.log
echo "== ncu full: linear GEMMs (pre-projections + first layer)"
timeout 900 $NCU --kernel-name-base demangled --set full --import-source on -k regex:"LinearEpi2|gemm_res_ln" -s 100 -c 7 -o gpurun_out/${TAG}_prof_gemm \
   python scripts/prof_step.py 256 3 --one-stream > gpurun_out/${TAG}_ncu_gemm.log 2>&1; echo rc=$?
echo "== ncu full: fused sim+NCE"
timeout 900 $NCU --set full --import-source on -k regex:"sim_fused_kernel|sim_reduce_partials" -s 8 -c 4 -o gpurun_out/${TAG}_prof_sim \
   python scripts/prof_step.py 256 3 --one-stream > gpurun_out/${TAG}_ncu_sim.log 2>&1; echo rc=$?
echo "== ncu full: attention"
timeout 900 $NCU --set full --import-source on -k regex:attention_kernel -s 24 -c 2 -o gpurun_out/${TAG}_prof_attn \
   python scripts/prof_step.py 256 3 --one-stream > gpurun_out/${TAG}_ncu_attn.log 2>&1; echo rc=$?
echo "== ncu full: layernorm"
timeout 900 $NCU --set full --import-source on -k regex:"layernorm_kernel" -s 62 -c 3 -o gpurun_out/${TAG}_prof_ln \
   python scripts/prof_step.py 256 3 --one-stream > gpurun_out/${TAG}_ncu_ln.log 2>&1; echo rc=$?
echo "== ncu full: streaming NCE"
timeout 900 $NCU --set full --import-source on -k regex:"nce_from_logits" -s 1 -c 1 -o gpurun_out/${TAG}_prof_nce \
   python scripts/prof_step.py 256 1 --hbm > gpurun_out/${TAG}_ncu_nce.log 2>&1; echo rc=$?
ls -la gpurun_out | grep ${TAG}
