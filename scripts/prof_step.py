"""Profiling driver: a few eager (no CUDA graph) steps of the bench workload, for ncu.
Usage: python scripts/prof_step.py [B_loc] [steps] [--hbm]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from temporalalignnet_b200 import ops  # noqa: E402
from temporalalignnet_b200.runner import TanStepRunner  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
r = TanStepRunner(6, 6, B, 256, 32, 512, 1024, device="cuda:0", use_graph=False)
r.model.two_streams = "--one-stream" not in sys.argv and r.model.two_streams
for i in range(steps):
    n0 = ops.launches()
    loss = r._step_kernels()
    torch.cuda.synchronize()
    print("step", i, "launches", ops.launches() - n0, "loss", float(loss), flush=True)
if "--hbm" in sys.argv:
    out = r.model(r.d_video, r.d_text, video_padding_mask=r.d_vpm, lang_padding_mask=r.d_tpm)
    dense = out["logits_joint"].materialize()
    Bq, S, T, B2, N = dense.shape
    g = ops.sim_geom(Bq, S, T, B2 * N, N, 1, 0)
    rs = torch.empty(2, Bq * S * T, dtype=torch.float32, device=dense.device)
    cs = torch.empty(2, S, B2 * N, dtype=torch.float32, device=dense.device)
    ws = torch.empty(ops.sim_workspace_bytes(g), dtype=torch.uint8, device=dense.device)
    from temporalalignnet_b200 import loss as loss_mod
    nce = loss_mod.prepare_nce_inputs(r.batch["start"], r.batch["end"], r.d_tpm, r.T, r.N, r.device, False, compact=False)
    for _ in range(2):
        ops.nce_from_logits(dense, g, nce.posbits, nce.col_valid, rs, cs, ws)
    torch.cuda.synchronize()
