#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -6
echo "== bench default"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo rc=$?; tail -c 600 gpurun_out/r02e_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02e_bench.json").read().strip().splitlines()[-1])
    for k in ("value", "ms_per_step", "loss", "loss_api", "e2e", "kernel_ms_per_step", "kernel_ms_sum", "loss_parity", "train_step", "eager_gpu_baseline", "cpu_baseline", "roofline_sim", "encoder_path"):
        print(k, json.dumps(d.get(k))[:600])
except Exception as e:
    print("bench parse failed", e)
PY
for cfg in 2 4 5; do
  echo "== bench config $cfg"; timeout 900 python bench.py --config $cfg --steps 5 --warmup 3 > gpurun_out/r02e_bench_c$cfg.json 2> gpurun_out/r02e_bench_c$cfg.err; echo rc=$?; tail -c 400 gpurun_out/r02e_bench_c$cfg.err
  python - $cfg <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r02e_bench_c{sys.argv[1]}.json").read().strip().splitlines()[-1])
    for k in ("value", "ms_per_step", "loss", "e2e", "kernel_ms_per_step", "loss_parity", "train_step", "eager_gpu_baseline", "cpu_baseline"):
        print(k, json.dumps(d.get(k))[:500])
except Exception as e:
    print("bench parse failed", e)
PY
done
