#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== backward kernel + train tests"; timeout 900 python -m pytest tests/test_backward_kernels_gpu.py tests/test_train_gpu.py tests/test_text_embed_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -3
for v in _base "" _base ""; do
  TAN_LIB_PATH=$PWD/temporalalignnet_b200/libtan_b200$v.so timeout 120 python scripts/train_profile.py 256 256 3 > gpurun_out/r02al_train$v.json 2> gpurun_out/r02al_train$v.err
  python - "$v" gpurun_out/r02al_train$v.json <<'PY'
import json, sys
d = json.load(open(sys.argv[2]))
print(f"variant '{sys.argv[1]}':", d["ms_per_train_step"], d["launches"], {k: v["ms"] for k, v in d["classes"].items() if k in ("colsum", "ln_bwd", "layernorm", "wgrad")})
PY
done
