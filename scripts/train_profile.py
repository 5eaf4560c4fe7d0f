"""Training-step timing at a given shape: ms per fwd+loss+bwd step (CUDA events) and a per-kernel-class
breakdown (event pairs around every C-ABI call; eager launches, so small classes include launch gaps).
  python scripts/train_profile.py [B] [T] [steps]
"""
import json
import sys
from collections import defaultdict

import torch

sys.path.insert(0, ".")
from temporalalignnet_b200 import ops  # noqa: E402
from temporalalignnet_b200.runner import TanStepRunner  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 256
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
r = TanStepRunner(6, 6, B_loc=B, T=T, use_graph=False)
for _ in range(2 if steps > 0 else 1):          # steps == 0: one warm-up + the profiled step (ncu launch lists)
    loss = r.step_train()
torch.cuda.synchronize()
ms = float("nan")
if steps > 0:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = r.step_train()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
n0 = ops.launches()
with ops.profile() as prof:
    r.step_train()
torch.cuda.synchronize()
launches = ops.launches() - n0
acc = defaultdict(lambda: [0.0, 0.0, 0])
for name, work, a, b in prof:
    acc[name][0] += a.elapsed_time(b)
    acc[name][1] += work
    acc[name][2] += 1
classes = {k: {"ms": round(v[0], 3), "calls": v[2], "tflops": round(v[1] / v[0] / 1e9, 1) if v[1] else None}
           for k, v in sorted(acc.items(), key=lambda kv: -kv[1][0])}
print(json.dumps({"B": B, "T": T, "ms_per_train_step": round(ms, 2), "clips_per_s": round(B / ms * 1e3, 1),
                  "loss": float(loss), "launches": launches, "mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 1),
                  "classes": classes}))
