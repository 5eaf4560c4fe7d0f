#!/bin/bash
# usage: gpurun --gpus N -- bash scripts/gpu_multi_b.sh N   (tests with the current kernels + bench at 1 and N)
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -3
timeout 300 python scripts/kernel_bench.py 2>&1 | grep attention
for n in 1 $N; do
  if [ $n = 1 ]; then timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --skip-cpu --skip-hbm > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 20 --warmup 5 --skip-cpu --skip-hbm > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; fi
  echo "bench n=$n rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n$n.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("n_gpus","value","ms_per_step","e2e","loss","loss_api","kernel_ms_per_step")})
except Exception as e:
    print("no json", e); print(open("gpurun_out/bench_n$n.err").read()[-1500:])
PY
done
