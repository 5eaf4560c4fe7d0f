#!/bin/bash
set -u
mkdir -p gpurun_out
for v in _base "" _base ""; do
  echo "== bench variant '$v'"
  TAN_LIB_PATH=$PWD/temporalalignnet_b200/libtan_b200$v.so timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-hbm --skip-eager --skip-train > gpurun_out/r02ah_bench$v.json 2> gpurun_out/r02ah_bench$v.err
  python - gpurun_out/r02ah_bench$v.json <<'PY'
import json, sys
for line in open(sys.argv[1]):
    if line.startswith("{"):
        d = json.loads(line)
        print({k: d.get(k) for k in ("value", "ms_per_step", "loss", "kernel_ms_per_step")}, "clk", d["clocks"]["sm_mhz"])
PY
done
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -p no:cacheprovider -k "linear" 2>&1 | tail -2
