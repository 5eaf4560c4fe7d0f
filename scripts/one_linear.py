"""Launch a few tan_linear_bf16 calls (for ncu captures).  usage: one_linear.py M N K act res out"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from temporalalignnet_b200 import ops
M, N, K, act, res = (int(x) for x in sys.argv[1:6])
out = sys.argv[6]
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
bias = torch.randn(N, device="cuda")
r = torch.randn(M, N, device="cuda") if res else None
of = torch.empty(M, N, device="cuda") if out == "f32" else None
ob = torch.empty(M, N, dtype=torch.bfloat16, device="cuda") if out == "bf16" else None
for _ in range(4):
    ops.linear(a, w, bias=bias, residual=r, out_f32=(r if res else of), out_bf16=ob, act=act)
torch.cuda.synchronize()
