"""Per-kernel timing at the BASELINE config-3 per-GPU shapes (CUDA events, L2 flushed between
iterations).  Development aid; prints one line per kernel with achieved TFLOP/s or GB/s."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from temporalalignnet_b200 import ops  # noqa: E402

DEV = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def timeit_chain(fn, n=20, iters=5):
    """Steady-state per-launch time: a CUDA graph of n back-to-back launches (warm L2, PDL overlap)."""
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / n)
    ts.sort()
    return ts[len(ts) // 2]


def bench_linear(M, N, K, act=0, res=False, out="bf16"):
    a = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    w = (torch.randn(N, K, device=DEV) * K ** -0.5).to(torch.bfloat16)
    bias = torch.randn(N, device=DEV)
    r = torch.randn(M, N, device=DEV) if res else None
    of = torch.empty(M, N, device=DEV) if out == "f32" else None
    ob = torch.empty(M, N, dtype=torch.bfloat16, device=DEV) if out == "bf16" else None
    ms = timeit(lambda: ops.linear(a, w, bias=bias, residual=r, out_f32=of if not res else r, out_bf16=ob, act=act))
    ms_t = timeit(lambda: torch.matmul(a, w.t()))
    ms_c = timeit_chain(lambda: ops.linear(a, w, bias=bias, residual=r, out_f32=of if not res else r, out_bf16=ob, act=act))
    oc = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    ms_tc = timeit_chain(lambda: torch.matmul(a, w.t(), out=oc))
    print(json.dumps({"kernel": "linear", "M": M, "N": N, "K": K, "act": act, "res": res, "ms": round(ms, 4),
                      "tflops": round(2 * M * N * K / ms / 1e9, 1), "cublas_ms": round(ms_t, 4),
                      "cublas_tflops": round(2 * M * N * K / ms_t / 1e9, 1),
                      "chain_ms": round(ms_c, 4), "chain_tflops": round(2 * M * N * K / ms_c / 1e9, 1),
                      "cublas_chain_ms": round(ms_tc, 4), "cublas_chain_tflops": round(2 * M * N * K / ms_tc / 1e9, 1)}),
          flush=True)


def bench_attn(B, H, L):
    d = H * 64
    qkv = torch.randn(B * L, 3 * d, device=DEV).to(torch.bfloat16)
    out = torch.empty(B * L, d, dtype=torch.bfloat16, device=DEV)
    ms = timeit(lambda: ops.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], None, out, B, H, L, L))
    fl = 4 * B * H * L * L * 64
    ms_c = timeit_chain(lambda: ops.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], None, out, B, H, L, L))
    print(json.dumps({"kernel": "attention", "B": B, "H": H, "L": L, "ms": round(ms, 4),
                      "tflops": round(fl / ms / 1e9, 1), "chain_ms": round(ms_c, 4),
                      "chain_tflops": round(fl / ms_c / 1e9, 1)}), flush=True)


def bench_ln(rows, d):
    x = torch.randn(rows, d, device=DEV)
    g, b = torch.ones(d, device=DEV), torch.zeros(d, device=DEV)
    ob = torch.empty(rows, d, dtype=torch.bfloat16, device=DEV)
    ms = timeit(lambda: ops.layernorm(x, rows, d, gamma=g, beta=b, out_bf16=ob))
    ms_c = timeit_chain(lambda: ops.layernorm(x, rows, d, gamma=g, beta=b, out_bf16=ob))
    print(json.dumps({"kernel": "layernorm", "rows": rows, "d": d, "ms": round(ms, 4),
                      "GBps": round(rows * d * 6 / ms / 1e6, 1), "chain_ms": round(ms_c, 4),
                      "chain_GBps": round(rows * d * 6 / ms_c / 1e6, 1)}), flush=True)


def bench_sim(B, S, T, N, d, Bglob, store):
    C = Bglob * N
    v = torch.randn(B, S, T, d, device=DEV)
    v = (v / v.norm(dim=-1, keepdim=True)).to(torch.bfloat16)
    t = torch.randn(S, C, d, device=DEV)
    t = (t / t.norm(dim=-1, keepdim=True)).to(torch.bfloat16)
    start = torch.randint(0, T, (B * N,), device=DEV).float()
    end = start + 4
    valid = torch.ones(C, dtype=torch.uint8, device=DEV)
    posbits = ops.pos_from_time(start, end, None, B, T, N)
    g = ops.sim_geom(B, S, T, C, N, d, 0)
    rs = torch.empty(2, B * S * T, device=DEV)
    cs = torch.empty(2, S, C, device=DEV)
    ws = torch.empty(ops.sim_workspace_bytes(g), dtype=torch.uint8, device=DEV)
    lg = torch.empty(B * S * T, C, dtype=torch.bfloat16, device=DEV) if store else None
    ms = timeit(lambda: ops.sim_nce_fwd(v, t, C * d, g, posbits, valid, lg, rs, cs, ws))
    fl = 2 * B * S * T * C * d
    rec = {"kernel": "sim_nce_fwd", "B": B, "S": S, "T": T, "C": C, "store": store, "ms": round(ms, 4),
           "tflops": round(fl / ms / 1e9, 1)}
    if store:
        rec["logit_write_GBps"] = round(B * S * T * C * 2 / ms / 1e6, 1)
        ms2 = timeit(lambda: ops.nce_from_logits(lg.view(B, S, T, Bglob, N), g, posbits, valid, rs, cs, ws))
        rec["nce_from_logits_ms"] = round(ms2, 4)
        rec["nce_from_logits_GBps"] = round(B * S * T * C * 2 / ms2 / 1e6, 1)
    print(json.dumps(rec), flush=True)


def sweep_linear():
    M = 32 * 256
    for (N, K, act, res, out) in ((1536, 512, 0, False, "bf16"), (512, 512, 0, True, "f32"),
                                  (2048, 512, 1, False, "bf16"), (512, 2048, 0, True, "f32")):
        for bn in (128, 256):
            for cs in (1, 2, 4):
                os.environ["TAN_GEMM_CS"], os.environ["TAN_GEMM_BN"] = str(cs), str(bn)
                print(f"cs={cs} bn={bn} ", end="")
                bench_linear(M, N, K, act=act, res=res, out=out)
    os.environ.pop("TAN_GEMM_CS"); os.environ.pop("TAN_GEMM_BN")
    for cs in (1, 2, 4):
        os.environ["TAN_SIM_CS"] = str(cs)
        print(f"cs={cs} ", end="")
        bench_sim(32, 6, 256, 32, 512, 32, False)
        print(f"cs={cs} ", end="")
        bench_sim(32, 6, 256, 32, 512, 256, False)
    os.environ.pop("TAN_SIM_CS")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "sweep":
        sweep_linear()
        sys.exit(0)
    M = 32 * 256
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    if only == "attn":
        bench_attn(32, 8, 256)
        bench_attn(32, 8, 288)
        bench_attn(4, 12, 1152)
        bench_ln(M, 512)
        sys.exit(0)
    bench_linear(M, 512, 1024, out="f32")
    bench_linear(M, 1536, 512)
    bench_linear(M, 512, 512, res=True, out="f32")
    bench_linear(M, 2048, 512, act=1)
    bench_linear(M, 512, 2048, res=True, out="f32")
    bench_linear(32 * 288, 1536, 512)
    bench_linear(32 * 288, 2048, 512, act=1)
    if only == "linear":
        sys.exit(0)
    bench_attn(32, 8, 256)
    bench_attn(32, 8, 288)
    bench_attn(4, 12, 1152)
    bench_ln(M, 512)
    bench_sim(32, 6, 256, 32, 512, 32, False)
    bench_sim(32, 6, 256, 32, 512, 32, True)
    bench_sim(32, 6, 256, 32, 512, 256, False)
    bench_sim(32, 6, 256, 32, 512, 256, True)
