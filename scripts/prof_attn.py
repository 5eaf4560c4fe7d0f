"""Profiling driver: the attention kernel alone at the bench shapes (for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from temporalalignnet_b200 import ops  # noqa: E402

B, H = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 8
d = H * 64
for L in (256, 288):
    qkv = torch.randn(B * L, 3 * d, device="cuda").to(torch.bfloat16)
    out = torch.empty(B * L, d, dtype=torch.bfloat16, device="cuda")
    for _ in range(3):
        ops.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], None, out, B, H, L, L)
    torch.cuda.synchronize()
