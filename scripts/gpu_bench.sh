#!/bin/bash
# bench + ncu evidence for one round.  Usage: gpurun -- bash scripts/gpu_bench.sh
set -u
mkdir -p gpurun_out
echo "== bench (graph)"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo rc=$?; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== bench (eager)"; timeout 600 python bench.py --steps 20 --warmup 5 --no-graph --skip-cpu > gpurun_out/bench_eager.json 2> gpurun_out/bench_eager.err; echo rc=$?; tail -c 1500 gpurun_out/bench_eager.json; tail -5 gpurun_out/bench_eager.err
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 420 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-graph --skip-cpu --skip-hbm > gpurun_out/ncu_list.log 2>&1; echo rc=$?
echo "== ncu full (gemm)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:umma_gemm -s 40 -c 6 -o gpurun_out/prof_gemm \
   python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-hbm > gpurun_out/ncu_gemm.log 2>&1; echo rc=$?
echo "== ncu full (attention, nce)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attention_kernel|nce_from_logits" -s 6 -c 3 -o gpurun_out/prof_attn \
   python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu > gpurun_out/ncu_attn.log 2>&1; echo rc=$?
ls -la gpurun_out
