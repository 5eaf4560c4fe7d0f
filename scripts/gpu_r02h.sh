#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -6
for flags in "TAN_FUSE_GELU=0" "TAN_FUSE_GELU=1" "TAN_FUSE_GELU=0" "TAN_FUSE_GELU=1"; do
  tag=$(echo "$flags" | tr ' =' '__')
  env $flags timeout 120 python scripts/train_profile.py 256 256 3 > gpurun_out/ab_${tag}.json 2> gpurun_out/ab_${tag}.err
  python - "$flags" gpurun_out/ab_${tag}.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    c = d["classes"]
    print(f'{sys.argv[1]:40s} {d["ms_per_train_step"]:8.2f} ms  mem {d["mem_gb"]}  ' + "  ".join(f'{k} {v["ms"]}' for k, v in c.items()) + f'  loss {d["loss"]:.6f}')
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  tail -2 gpurun_out/ab_${tag}.err
done
