#!/bin/bash
# round-2 ab: linear epilogue specialised by activation at compile time: A/B
set -u
mkdir -p gpurun_out
for v in _base ""; do
  echo "== bench variant '$v'"
  TAN_LIB_PATH=$PWD/temporalalignnet_b200/libtan_b200$v.so timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-hbm --skip-eager > gpurun_out/r02ab_bench$v.json 2> gpurun_out/r02ab_bench$v.err
  python - gpurun_out/r02ab_bench$v.json <<'PY'
import json, sys
for line in open(sys.argv[1]):
    if line.startswith("{"):
        d = json.loads(line)
        print({k: d.get(k) for k in ("value", "ms_per_step", "loss", "kernel_ms_per_step")})
        print("  e2e", d["e2e"]["value"], "linear frac", d["roofline"]["frac"], "sim", d["roofline_sim"]["frac"], "attn", d["roofline_attention"]["frac"], "enc", d["encoder_path"]["frac"], "train", d.get("train_step", {}).get("ms_per_step"), "clk", d["clocks"]["sm_mhz"])
PY
done
echo "== gpu tests"; timeout 1200 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -3
