#!/bin/bash
# round-2 r: per-(clip, head) attention kernel with elect-issued MMAs + plain waits: A/B, whole GPU suite, bench
set -u
mkdir -p gpurun_out
for v in _noelect ""; do echo "== timings variant '$v'"; TAN_LIB_PATH=$PWD/temporalalignnet_b200/libtan_b200$v.so timeout 200 python scripts/attn_time.py 256 8 256 256 8 288 32 8 256 32 8 288 32 12 1024 32 12 1152 128 8 512 128 8 576 64 8 64 2>&1 | tail -9; done
echo "== gpu tests"; timeout 1200 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -4
echo "== bench"; timeout 900 python bench.py > gpurun_out/r02r_bench.json 2> gpurun_out/r02r_bench.err; tail -c 600 gpurun_out/r02r_bench.err
python - <<'PY'
import json
for line in open("gpurun_out/r02r_bench.json"):
    if line.startswith("{"):
        d = json.loads(line)
        print({k: d.get(k) for k in ("value", "ms_per_step", "loss", "kernel_ms_per_step", "gpu_launches_per_step")})
        print("e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], "attn", d["roofline_attention"], "enc", d["encoder_path"]["frac"])
        print("train", d.get("train_step", {}).get("ms_per_step"), d.get("train_step", {}).get("value"))
PY
