"""A/B timing of single kernels at the bench shapes (B=256): CUDA-graph chains, L2-cold singles."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from temporalalignnet_b200 import ops
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from kernel_bench import timeit, timeit_chain  # noqa
DEV = "cuda"
tag = os.environ.get("TAN_GEMM_STREAMING", "resident")
for (M, N, K, act) in ((65536, 1536, 512, 0), (65536, 2048, 512, 1), (73728, 1536, 512, 0), (73728, 2048, 512, 1)):
    a = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    w = (torch.randn(N, K, device=DEV) * K ** -0.5).to(torch.bfloat16)
    bias = torch.randn(N, device=DEV)
    ob = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    fn = lambda: ops.linear(a, w, bias=bias, out_bf16=ob, act=act)
    ms, msc = timeit(fn), timeit_chain(fn, n=10)
    print(json.dumps({"mode": tag, "M": M, "N": N, "K": K, "ms_cold": round(ms, 4), "ms_chain": round(msc, 4),
                      "tflops_chain": round(2 * M * N * K / msc / 1e9, 1)}), flush=True)
if "attn" in sys.argv:
    for (B, H, L) in ((256, 8, 256), (256, 8, 288)):
        d = H * 64
        qkv = torch.randn(B * L, 3 * d, device=DEV).to(torch.bfloat16)
        out = torch.empty(B * L, d, dtype=torch.bfloat16, device=DEV)
        fn = lambda: ops.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], None, out, B, H, L, L)
        ms, msc = timeit(fn), timeit_chain(fn, n=10)
        print(json.dumps({"kernel": "attention", "B": B, "L": L, "ms_cold": round(ms, 4), "ms_chain": round(msc, 4),
                          "tflops_chain": round(4 * B * H * L * L * 64 / msc / 1e9, 1)}), flush=True)
if "sim" in sys.argv:
    B, S, T, N, d = 256, 6, 256, 32, 512
    C = B * N
    v = torch.randn(B, S, T, d, device=DEV); v = (v / v.norm(dim=-1, keepdim=True)).to(torch.bfloat16)
    t = torch.randn(S, C, d, device=DEV); t = (t / t.norm(dim=-1, keepdim=True)).to(torch.bfloat16)
    start = torch.randint(0, T, (B * N,), device=DEV).float()
    posbits = ops.pos_from_time(start, start + 4, None, B, T, N)
    valid = (torch.rand(C, device=DEV) < 0.75).to(torch.uint8)
    g = ops.sim_geom(B, S, T, C, N, d, 0)
    rs = torch.empty(2, B * S * T, device=DEV); cs = torch.empty(2, S, C, device=DEV)
    ws = torch.empty(ops.sim_workspace_bytes(g), dtype=torch.uint8, device=DEV)
    fn = lambda: ops.sim_nce_fwd(v, t, C * d, g, posbits, valid, None, rs, cs, ws)
    ms, msc = timeit(fn), timeit_chain(fn, n=5)
    print(json.dumps({"kernel": "sim_nce_fwd", "ms_cold": round(ms, 4), "ms_chain": round(msc, 4),
                      "tflops_chain": round(2 * B * S * T * C * d / msc / 1e9, 1)}), flush=True)
