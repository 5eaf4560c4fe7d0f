#!/bin/bash
# TN GEMM (tan_gemm_tn_bf16) validation + A/B of the training step with / without it.
set -u
mkdir -p gpurun_out
echo "== kernel tests"; timeout 200 python -m pytest tests/test_backward_kernels_gpu.py -q -p no:cacheprovider -k "gemm_tn" 2>&1 | tail -15
echo "== train parity (TN on)"; timeout 300 python -m pytest tests/test_train_gpu.py -q -p no:cacheprovider 2>&1 | tail -3
for flags in "TAN_TN_GEMM=0" "TAN_TN_GEMM=1" "TAN_TN_GEMM=1 TAN_ATTN_BWD=pipe"; do
  tag=$(echo "$flags" | tr ' =' '__')
  env $flags timeout 120 python scripts/train_profile.py 256 256 3 > gpurun_out/tn_${tag}.json 2> gpurun_out/tn_${tag}.err
  python - "$flags" gpurun_out/tn_${tag}.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    c = d["classes"]
    print(f'{sys.argv[1]:40s} {d["ms_per_train_step"]:8.2f} ms  ' + "  ".join(f'{k} {v["ms"]}' for k, v in c.items()) + f'  loss {d["loss"]:.6f}')
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
