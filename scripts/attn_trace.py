"""Per-CTA timeline of one tan_attention_bf16 launch (tan_debug_set_trace).  Needs the trace build of the attention
kernel: make -C temporalalignnet_b200/csrc variant SRC=attention NAME=trace FLAGS="-DTAN_WAIT_HINT_ALL=0 -DTAN_ATT_TRACE"
and TAN_LIB_PATH=$PWD/temporalalignnet_b200/libtan_b200_trace.so.  usage: B H L"""
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
tr = torch.zeros(512 * 128, dtype=torch.int64, device="cuda")
_lib.check(_lib.lib().tan_debug_set_trace(tr.data_ptr()))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); flush.zero_()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record(); run(); e1.record()
torch.cuda.synchronize()
_lib.check(_lib.lib().tan_debug_set_trace(None))
t = tr.cpu().numpy().reshape(512, 128)
t = t[t[:, 1] != 0]
print(f"B={B} H={H} L={L}: event time {e0.elapsed_time(e1) * 1e3:.1f} us, traced CTAs {len(t)}")
rel = lambda col: (t[:, col] - t[:, 1])
print(f"prologue done (clk): median {np.median(rel(2)):.0f}   CTA lifetime: median {np.median(rel(3)):.0f} max {rel(3).max():.0f}")
names = ["K_issued", "V_issued", "QK_issued", "p_ready_seen", "PV_issued", "s_full_seen", "S_in_regs", "P_staged"]
nb = (L + 63) // 64
for g in range(min(10, 2 * nb)):
    vals = []
    for k in range(8):
        c = 8 + 10 * g + k
        ok = t[:, c] != 0
        vals.append(np.median(t[ok][:, c] - t[ok][:, 1]) if ok.any() else float("nan"))
    print(f"block {g}: " + "  ".join(f"{n}={v:7.0f}" for n, v in zip(names, vals)))
# first-wave vs later-wave CTAs
first = t[:, 0] <= np.sort(t[:, 0])[min(len(t) - 1, 295)]
print(f"first-wave CTAs lifetime median {np.median(rel(3)[first]):.0f}, later {np.median(rel(3)[~first]) if (~first).any() else float('nan'):.0f}")
