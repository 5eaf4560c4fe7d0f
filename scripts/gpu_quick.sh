#!/bin/bash
# full GPU test suite + kernel bench + bench
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -n 25 gpurun_out/tests.log
timeout 300 python scripts/kernel_bench.py > gpurun_out/kbench.log 2>&1; echo "kbench rc=$?"; grep -E "sim_nce|attention" gpurun_out/kbench.log
timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
for k in ("value","ms_per_step","e2e","kernel_ms_per_step","roofline_sim","roofline_nce_hbm","roofline_attention","loss","clocks"): print(k, d.get(k))
print("linear", d["roofline"]["achieved"], d["roofline"]["frac"])
PY
tail -3 gpurun_out/bench.err
