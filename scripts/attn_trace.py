"""Per-CTA timeline of one tan_attention_bf16 launch (tan_debug_set_trace; slot layout: csrc/attention_pp.cu).
usage: attn_trace.py B H L"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from temporalalignnet_b200 import ops, _lib
B, H, L = (int(x) for x in sys.argv[1:4])
d = H * 64
qkv = torch.randn(B * L, 3 * d, device="cuda").to(torch.bfloat16)
out = torch.empty(B * L, d, dtype=torch.bfloat16, device="cuda")
run = lambda: ops.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], None, out, B, H, L, L)
for _ in range(3):
    run()
torch.cuda.synchronize()
NCTA, SLOTS = 148, 256
tr = torch.zeros(NCTA * SLOTS, dtype=torch.int64, device="cuda")
_lib.check(_lib.lib().tan_debug_set_trace(tr.data_ptr()))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); flush.zero_()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record(); run(); e1.record()
torch.cuda.synchronize()
_lib.check(_lib.lib().tan_debug_set_trace(None))
t = tr.cpu().numpy().reshape(NCTA, SLOTS)
t = t[t[:, 1] != 0]
print(f"B={B} H={H} L={L}: event time {e0.elapsed_time(e1) * 1e3:.1f} us, traced CTAs {len(t)}")
rel = lambda col: (t[:, col] - t[:, 1])
print(f"prologue done (clk): median {np.median(rel(2)):.0f}   CTA lifetime: median {np.median(rel(3)):.0f} max {rel(3).max():.0f}")
names = ["K_iss", "V_iss", "QK_A", "QK_B", "prdyA", "PV_A", "prdyB", "PV_B", "sfulA", "SregA", "PstgA", "sfulB", "SregB",
         "PstgB", "epiA", "epiB"]
print("block  " + " ".join(f"{n:>6s}" for n in names))
for s in range(14):
    vals = []
    for k in range(16):
        c = 8 + 16 * s + k
        ok = t[:, c] != 0
        vals.append(np.median(t[ok][:, c] - t[ok][:, 1]) if ok.any() else float("nan"))
    if all(np.isnan(v) for v in vals):
        break
    print(f"{s:5d}  " + " ".join(f"{v:6.0f}" for v in vals))
