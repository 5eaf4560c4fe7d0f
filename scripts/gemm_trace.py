"""Per-CTA timeline of one tan_linear_bf16 launch (uses tan_debug_set_trace).  usage: M N K act res out"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from temporalalignnet_b200 import ops, _lib
M, N, K, act, res = (int(x) for x in sys.argv[1:6])
out = sys.argv[6]
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
bias = torch.randn(N, device="cuda")
r = torch.randn(M, N, device="cuda") if res else None
of = torch.empty(M, N, device="cuda") if out == "f32" else None
ob = torch.empty(M, N, dtype=torch.bfloat16, device="cuda") if out == "bf16" else None
run = lambda: ops.linear(a, w, bias=bias, residual=r, out_f32=(r if res else of), out_bf16=ob, act=act)
for _ in range(3):
    run()
torch.cuda.synchronize()
tr = torch.zeros(148 * 64, dtype=torch.int64, device="cuda")
_lib.check(_lib.lib().tan_debug_set_trace(tr.data_ptr()))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); flush.zero_()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record(); run(); e1.record()
torch.cuda.synchronize()
_lib.check(_lib.lib().tan_debug_set_trace(None))
t = tr.cpu().numpy().reshape(148, 64)
print(f"shape M={M} N={N} K={K} act={act} res={res} out={out}: event time {e0.elapsed_time(e1)*1e3:.1f} us")
used = t[:, 1] != 0
t = t[used]
gt0 = t[:, 0].min()
print(f"CTAs {used.sum()}; globaltimer start spread {(t[:,0].max()-gt0)/1e3:.2f} us")
clk = 1.9  # GHz nominal, cycles -> us approx
def us(c): return c / 1e3 / clk
rel = lambda col: us(t[:, col] - t[:, 1])
print(f"prologue done: median {np.median(rel(2)):.2f} us  max {rel(2).max():.2f}")
print(f"kernel exit:   median {np.median(rel(3)):.2f} us  max {rel(3).max():.2f}")
names = ["prod_first", "prod_last", "mma_first", "mma_last", "epi_pre", "epi_bufs_free", "epi_acc_ready", "epi_end"]
for it in range(3):
    cols = [4 + it * 8 + j for j in range(8)]
    have = t[:, cols[0]] != 0
    if not have.any():
        break
    def med(c):
        ok = have & (t[:, c] != 0)
        return np.median(us(t[ok][:, c] - t[ok][:, 1])) if ok.any() else float("nan")
    s = " ".join(f"{n}={med(c):6.2f}" for n, c in zip(names, cols))
    print(f"tile {it} ({have.sum():3d} CTAs): {s}")
