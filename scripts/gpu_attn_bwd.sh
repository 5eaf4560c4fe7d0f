#!/bin/bash
# tcgen05 attention backward: kernel test, training-step parity, A/B against the pipelined mma.sync version.
set -u
mkdir -p gpurun_out
echo "== kernel tests"; timeout 300 python -m pytest tests/test_backward_kernels_gpu.py tests/test_kernels_gpu.py -q -p no:cacheprovider -k "attention" 2>&1 | tail -25
echo "== train parity"; timeout 300 python -m pytest tests/test_train_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -8
for flags in "TAN_ATTN_BWD=pipe" "TAN_ATTN_BWD=tc"; do
  tag=$(echo "$flags" | tr ' =' '__')
  env $flags timeout 120 python scripts/train_profile.py 256 256 3 > gpurun_out/ab_${tag}.json 2> gpurun_out/ab_${tag}.err
  python - "$flags" gpurun_out/ab_${tag}.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    c = d["classes"]
    print(f'{sys.argv[1]:40s} {d["ms_per_train_step"]:8.2f} ms  ' + "  ".join(f'{k} {v["ms"]}' for k, v in c.items()) + f'  loss {d["loss"]:.6f}')
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  tail -3 gpurun_out/ab_${tag}.err
done
