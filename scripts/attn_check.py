"""Where is tan_attention_bf16 wrong?  max |error| per (clip, head, 32-row group).  usage: attn_check.py B H L [masked]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from temporalalignnet_b200 import ops
B, H, L = (int(x) for x in sys.argv[1:4])
masked = len(sys.argv) > 4 and sys.argv[4] == "1"
d = H * 64
g = torch.Generator(device="cuda").manual_seed(20)
qkv = torch.randn(B * L, 3 * d, device="cuda", generator=g).to(torch.bfloat16)
kpm = None
if masked:
    kpm = torch.zeros(B, L, dtype=torch.uint8, device="cuda"); kpm[0, L - L // 4:] = 1; kpm[-1, 1::3] = 1
out = torch.empty(B * L, d, dtype=torch.bfloat16, device="cuda")
ops.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], kpm, out, B, H, L, L)
q, k, v = (x.float().view(B, L, H, 64).transpose(1, 2) for x in (qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]))
s = q @ k.transpose(-1, -2) / 8.0
if kpm is not None:
    s = s.masked_fill(kpm.bool()[:, None, None, :], float("-inf"))
ref = torch.softmax(s, -1) @ v                                  # [B, H, L, 64]
got = out.float().view(B, L, H, 64).transpose(1, 2)
err = (got - ref).abs().amax(-1)                                # [B, H, L]
ng = (L + 31) // 32
bad = 0
for b in range(B):
    for h in range(H):
        e = [err[b, h, g0 * 32:(g0 + 1) * 32].max().item() for g0 in range(ng)]
        if max(e) > 0.05:
            bad += 1
            if bad <= 24:
                print(f"clip {b} head {h}: " + " ".join(f"{x:5.2f}" for x in e))
print(f"B={B} H={H} L={L} masked={masked}: {bad} bad (clip, head) of {B * H}; max err {err.max().item():.3f}, nan {torch.isnan(got).sum().item()}")
