#!/bin/bash
# what the driver runs at round end, on the final tree: GPU suite, smoke, default bench (+ reference arm)
set -u
TAG=${1:-r02ai}
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1200 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -3
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo rc=$?
python - gpurun_out/${TAG}_bench.json <<'PY'
import json, sys
for line in open(sys.argv[1]):
    if line.startswith("{"):
        d = json.loads(line)
        print({k: d.get(k) for k in ("value", "ms_per_step", "loss", "kernel_ms_per_step", "gpu_launches_per_step")})
        print("e2e", d["e2e"], "roofline", d["roofline"]["frac"], "sim", d["roofline_sim"]["frac"], "attn", d["roofline_attention"]["frac"], "enc", d["encoder_path"]["frac"], "step", d["whole_step_tensor_frac"])
        print("train", d["train_step"]["ms_per_step"], d["train_step"]["value"], "eager", d["eager_gpu_baseline"]["fp16_autocast"]["value"], d["eager_gpu_baseline"]["fp16_autocast_train_step"]["value"], "cpu", d["cpu_baseline"]["value"], "parity", d["loss_parity"]["rel_err"], "clk", d["clocks"])
PY
