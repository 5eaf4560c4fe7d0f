"""A/B: fused projection+residual+LayerNorm kernels vs the unfused pair, per shape (CUDA-graph chains)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from temporalalignnet_b200 import ops
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from kernel_bench import timeit, timeit_chain  # noqa
DEV = "cuda"
for M in (8192, 9216, 16384, 32768, 65536, 73728):
    for K in (512, 2048):
        L = 288 if M % 288 == 0 else 256
        a = torch.randn(M, K, device=DEV).to(torch.bfloat16)
        w = (torch.randn(512, K, device=DEV) * K ** -0.5).to(torch.bfloat16)
        bias = torch.randn(512, device=DEV); gamma = torch.ones(512, device=DEV); beta = torch.zeros(512, device=DEV)
        x = torch.randn(M, 512, device=DEV); xn = torch.empty(M, 512, dtype=torch.bfloat16, device=DEV)
        nrm = torch.empty(M, 512, dtype=torch.bfloat16, device=DEV)
        def unfused():
            ops.linear(a, w, bias=bias, residual=x, out_f32=x)
            ops.layernorm(x, M, 512, gamma=gamma, beta=beta, L_in=L, out_bf16=xn, l_split=L, strideA=L,
                          nrmA_bf16=nrm if K == 2048 else None)
        def fused():
            if K == 2048:
                ops.linear_res_ln_stage(a, w, bias, x, gamma, beta, xn, L, L, nrm, L, None, 0)
            else:
                ops.linear_res_ln(a, w, bias, x, gamma, beta, xn)
        u, f = timeit_chain(unfused, n=10), timeit_chain(fused, n=10)
        uc, fc = timeit(unfused), timeit(fused)
        print(json.dumps({"M": M, "K": K, "unfused_chain_us": round(u * 1e3, 1), "fused_chain_us": round(f * 1e3, 1),
                          "unfused_cold_us": round(uc * 1e3, 1), "fused_cold_us": round(fc * 1e3, 1)}), flush=True)
