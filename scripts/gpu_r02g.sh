#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -12
for flags in "TAN_COMPACT_COLUMNS=0" "TAN_COMPACT_COLUMNS=1"; do
  tag=$(echo "$flags" | tr ' =' '__')
  env $flags timeout 600 python bench.py --steps 10 --warmup 3 --skip-eager --skip-cpu --skip-hbm > gpurun_out/r02g_${tag}.json 2> gpurun_out/r02g_${tag}.err; echo "$flags rc=$?"; tail -c 300 gpurun_out/r02g_${tag}.err
  python - gpurun_out/r02g_${tag}.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    for k in ("value", "ms_per_step", "loss", "loss_api", "e2e", "kernel_ms_per_step", "roofline_sim", "train_step"):
        print(k, json.dumps(d.get(k))[:500])
except Exception as e:
    print("bench parse failed", e)
PY
done
