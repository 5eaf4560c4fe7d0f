#!/bin/bash
# One GPU call that validates and A/Bs the experiments prepared at the end of round 1 (none of them is the default):
#   TAN_ATTN_BWD=pipe      cp.async-pipelined mma.sync attention backward      (attention_bwd_pipe.cu)
#   TAN_SIM_GRAD_GT=1      similarity-gradient epilogue that also writes G^T   (sim_grad_gemm_gt.cu)
#   TAN_FUSE_BIAS_SUM=1    dY transpose that also produces the bias gradients  (backward_fused.cu)
# Usage: gpurun --timeout 400 -- bash scripts/gpu_experiments.sh      (about 2-3 minutes of box time)
# A variant is ready to become the default when its tests pass and its ms_per_train_step is lower.
set -u
mkdir -p gpurun_out
echo "== kernel-level tests of the experimental variants"
TAN_TEST_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_backward_kernels_gpu.py -q -p no:cacheprovider 2>&1 | tail -4
for flags in "TAN_ATTN_BWD=pipe" "TAN_SIM_GRAD_GT=1" "TAN_FUSE_BIAS_SUM=1"; do
  echo "== training-step parity with $flags"
  env $flags timeout 200 python -m pytest tests/test_train_gpu.py -q -p no:cacheprovider 2>&1 | tail -2
done
echo "== A/B at the bench shape (ms_per_train_step)"
for flags in "TAN_NONE=1" "TAN_ATTN_BWD=pipe" "TAN_SIM_GRAD_GT=1" "TAN_FUSE_BIAS_SUM=1" "TAN_ATTN_BWD=pipe TAN_SIM_GRAD_GT=1 TAN_FUSE_BIAS_SUM=1"; do
  tag=$(echo "$flags" | tr ' =' '__')
  env $flags timeout 120 python scripts/train_profile.py 256 256 3 > gpurun_out/exp_${tag}.json 2> gpurun_out/exp_${tag}.err
  python - "$flags" gpurun_out/exp_${tag}.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    c = d["classes"]
    print(f'{sys.argv[1]:70s} {d["ms_per_train_step"]:8.2f} ms  attention_bwd {c.get("attention_bwd", {}).get("ms")}  '
          f'transpose {c.get("transpose", {}).get("ms")}  colsum {c.get("colsum", {}).get("ms")}  loss {d["loss"]:.6f}')
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
