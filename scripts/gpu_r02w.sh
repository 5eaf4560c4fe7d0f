#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== gpu suite"; timeout 1200 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -3
for i in 1 2; do
timeout 120 python scripts/train_profile.py 256 256 3 > gpurun_out/r02w_train.json 2> gpurun_out/r02w_train.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02w_train.json"))
print(d["ms_per_train_step"], d["launches"], {k: v["ms"] for k, v in d["classes"].items() if v["ms"] > 0.4})
PY
done
