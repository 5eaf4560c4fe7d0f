"""Event-timed tan_attention_bf16 launches (L2 flushed between launches).  usage: attn_time.py B H L [B H L ...]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from temporalalignnet_b200 import ops
a = [int(x) for x in sys.argv[1:]]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for B, H, L in zip(a[0::3], a[1::3], a[2::3]):
    d = H * 64
    qkv = torch.randn(B * L, 3 * d, device="cuda").to(torch.bfloat16)
    kpm = torch.zeros(B, L, dtype=torch.uint8, device="cuda")
    kpm[:, L - 3:] = 1
    out = torch.empty(B * L, d, dtype=torch.bfloat16, device="cuda")
    run = lambda: ops.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], kpm, out, B, H, L, L)
    for _ in range(3):
        run()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); run(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    fl = 4.0 * B * H * L * L * 64
    print(f"B={B} H={H} L={L}: median {ts[5]:.1f} us, min {ts[0]:.1f} us, {fl / ts[5] * 1e-6:.0f} TFLOP/s")
