#!/bin/bash
# final single-GPU records of the round: whole GPU suite, default bench, the other BASELINE configs
set -u
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1200 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -3
echo "== bench (default)"; timeout 900 python bench.py > gpurun_out/r02ae_bench.json 2> gpurun_out/r02ae_bench.err; tail -c 300 gpurun_out/r02ae_bench.err
for c in 2 4 5; do
  timeout 600 python bench.py --config $c --skip-cpu --skip-eager --skip-hbm > gpurun_out/r02ae_bench_c$c.json 2> gpurun_out/r02ae_bench_c$c.err
done
python - <<'PY'
import json
for f in ["r02ae_bench", "r02ae_bench_c2", "r02ae_bench_c4", "r02ae_bench_c5"]:
    try:
        for line in open(f"gpurun_out/{f}.json"):
            if line.startswith("{"):
                d = json.loads(line)
                print(f, {k: d.get(k) for k in ("value", "ms_per_step", "loss", "kernel_ms_per_step")}, "e2e", d["e2e"]["value"], "lin", d["roofline"]["frac"], "sim", d["roofline_sim"]["frac"], "attn", d["roofline_attention"]["frac"], "enc", d["encoder_path"]["frac"], "step", d.get("whole_step_tensor_frac"), "train", (d.get("train_step") or {}).get("ms_per_step"), "clk", d["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "failed", e)
PY
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | cut -c1-300
