"""Per-kernel numerics tests (GPU): every C-ABI entry point against a plain PyTorch fp32 reference
of the same op on the same (bf16-rounded) inputs.  Tolerances are written next to each check."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    from temporalalignnet_b200 import ops
    return ops


def _rand(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def test_device_is_sm100():
    from temporalalignnet_b200 import _lib
    assert _lib.lib().tan_device_check() == 0, _lib.lib().tan_last_error_string()


def test_cast():
    ops = _ops()
    x = _rand(1000, 64)
    y = ops.cast_bf16(x)
    assert torch.equal(y, x.to(torch.bfloat16))          # round-to-nearest-even, bit exact


@pytest.mark.parametrize("M,N,K", [
    (128, 256, 64), (128, 128, 512), (256, 512, 512), (8192, 512, 512), (1000, 1536, 512),
    (77, 2048, 512), (4608, 512, 2048), (300, 768, 3072), (128, 128, 64), (9216, 1536, 512), (33, 384, 128),
])
def test_linear_plain(M, N, K):
    ops = _ops()
    a = _rand(M, K, seed=1).to(torch.bfloat16)
    w = _rand(N, K, scale=K ** -0.5, seed=2).to(torch.bfloat16)
    out = torch.empty(M, N, dtype=torch.float32, device=DEV)
    ops.linear(a, w, out_f32=out)
    ref = a.float() @ w.float().t()
    # fp32 accumulation of exact bf16 products: only summation-order noise
    assert (out - ref).abs().max().item() < 2e-3 * max(1.0, ref.abs().max().item())
    assert torch.isfinite(out).all()


@pytest.mark.parametrize("act", [0, 1])
@pytest.mark.parametrize("with_res", [False, True])
def test_linear_epilogue(act, with_res):
    ops = _ops()
    M, N, K = 1000, 512, 512
    a = _rand(M, K, seed=3).to(torch.bfloat16)
    w = _rand(N, K, scale=K ** -0.5, seed=4).to(torch.bfloat16)
    bias = _rand(N, seed=5)
    res = _rand(M, N, seed=6) if with_res else None
    out_f = torch.empty(M, N, dtype=torch.float32, device=DEV)
    out_b = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    ops.linear(a, w, bias=bias, residual=res, out_f32=out_f, out_bf16=out_b, act=act)
    ref = a.float() @ w.float().t() + bias
    if act:
        ref = ref * torch.sigmoid(1.702 * ref)
    if with_res:
        ref = ref + res
    # tanh.approx-based QuickGELU: abs error <= ~1e-3 * |x|
    tol = 5e-3 if act else 2e-3
    assert (out_f - ref).abs().max().item() < tol * max(1.0, ref.abs().max().item())
    assert (out_b.float() - ref).abs().max().item() < 1e-2 * max(1.0, ref.abs().max().item())


def test_linear_residual_in_place_and_strided_views():
    ops = _ops()
    M, d = 640, 512
    x = _rand(M, d, seed=7)
    x0 = x.clone()
    a = _rand(M, d, seed=8).to(torch.bfloat16)
    w = _rand(d, d, scale=d ** -0.5, seed=9).to(torch.bfloat16)
    ops.linear(a, w, residual=x, out_f32=x)                          # aliasing allowed by the ABI
    ref = x0 + a.float() @ w.float().t()
    assert (x - ref).abs().max().item() < 2e-3 * ref.abs().max().item()
    # column-sliced output (ld > N) and row-sliced weight
    big = torch.zeros(M, 3 * d, dtype=torch.bfloat16, device=DEV)
    w3 = _rand(3 * d, d, scale=d ** -0.5, seed=10).to(torch.bfloat16)
    ops.linear(a, w3[d:2 * d], out_bf16=big[:, d:2 * d])
    ref = a.float() @ w3[d:2 * d].float().t()
    assert (big[:, d:2 * d].float() - ref).abs().max().item() < 1e-2 * ref.abs().max().item()
    assert big[:, :d].abs().max().item() == 0 and big[:, 2 * d:].abs().max().item() == 0


@pytest.mark.parametrize("M,K", [(256, 512), (8192, 512), (1000, 512), (300, 2048), (37, 64)])
def test_linear_res_ln(M, K):
    """Fused out-projection + residual + LayerNorm against torch fp32 on the bf16-rounded operands."""
    ops = _ops()
    g = torch.Generator().manual_seed(M + K)
    a = torch.randn(M, K, generator=g).to(DEV).to(torch.bfloat16)
    w = (torch.randn(512, K, generator=g) * K ** -0.5).to(DEV).to(torch.bfloat16)
    bias = torch.randn(512, generator=g).to(DEV)
    x0 = (torch.randn(M, 512, generator=g) * 1.5 + 0.3).to(DEV)
    gamma = (1 + 0.1 * torch.randn(512, generator=g)).to(DEV)
    beta = (0.1 * torch.randn(512, generator=g)).to(DEV)
    x = x0.clone()
    out = torch.full((M, 512), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.linear_res_ln(a, w, bias, x, gamma, beta, out)
    x_ref = x0 + a.float() @ w.float().t() + bias
    y_ref = torch.nn.functional.layer_norm(x_ref, (512,), gamma, beta, 1e-5)
    assert (x - x_ref).abs().max().item() < 2e-3
    assert (out.float() - y_ref).abs().max().item() < 3e-2          # bf16 output of O(1..4) values
    assert ((out.float() - y_ref).norm() / y_ref.norm()).item() < 4e-3


@pytest.mark.parametrize("B,L,l_split,K,want_out", [(3, 288, 256, 2048, True), (5, 64, 64, 512, True), (2, 288, 256, 512, False)])
def test_linear_res_ln_stage(B, L, l_split, K, want_out):
    """c_proj + residual + LayerNorm + L2-normalised bf16 stage features scattered by clip (video | text parts)."""
    ops = _ops()
    M = B * L
    g = torch.Generator().manual_seed(B * L + K)
    a = torch.randn(M, K, generator=g).to(DEV).to(torch.bfloat16)
    w = (torch.randn(512, K, generator=g) * K ** -0.5).to(DEV).to(torch.bfloat16)
    bias = torch.randn(512, generator=g).to(DEV)
    x0 = (torch.randn(M, 512, generator=g) * 1.5 + 0.3).to(DEV)
    gamma = (1 + 0.1 * torch.randn(512, generator=g)).to(DEV)
    beta = (0.1 * torch.randn(512, generator=g)).to(DEV)
    x = x0.clone()
    out = torch.full((M, 512), float("nan"), dtype=torch.bfloat16, device=DEV) if want_out else None
    S, s_idx = 3, 1                                      # features land in stage 1 of a [B, S, T, d] / [S, B*N, d] pair
    T, N = l_split, L - l_split
    nrmA = torch.full((B, S, T, 512), float("nan"), dtype=torch.bfloat16, device=DEV)
    nrmB = torch.full((S, B * N, 512), float("nan"), dtype=torch.bfloat16, device=DEV) if N > 0 else None
    vA = nrmA.view(-1, 512)[s_idx * T:]
    vB = nrmB.view(-1, 512)[s_idx * B * N:] if N > 0 else None
    ops.linear_res_ln_stage(a, w, bias, x, gamma, beta, out, L, l_split, vA, S * T, vB, N)
    x_ref = x0 + a.float() @ w.float().t() + bias
    y_ref = torch.nn.functional.layer_norm(x_ref, (512,), gamma, beta, 1e-5)
    n_ref = (y_ref / y_ref.norm(dim=-1, keepdim=True)).view(B, L, 512)
    assert (x - x_ref).abs().max().item() < 2e-3
    if want_out:
        assert ((out.float() - y_ref).norm() / y_ref.norm()).item() < 4e-3
    assert (nrmA[:, s_idx].float() - n_ref[:, :T]).abs().max().item() < 2e-3       # |values| <= 1, bf16
    assert torch.isnan(nrmA[:, 0].float()).all() and torch.isnan(nrmA[:, 2].float()).all()   # other stages untouched
    if N > 0:
        assert (nrmB[s_idx].view(B, N, 512).float() - n_ref[:, T:]).abs().max().item() < 2e-3
        assert torch.isnan(nrmB[0].float()).all() and torch.isnan(nrmB[2].float()).all()


def test_linear_rejects_bad_shapes():
    from temporalalignnet_b200 import TanError
    ops = _ops()
    a = torch.zeros(128, 100, dtype=torch.bfloat16, device=DEV)      # K % 64 != 0
    w = torch.zeros(128, 100, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(TanError, match="TAN_ERR_SHAPE"):
        ops.linear(a, w, out_f32=torch.empty(128, 128, device=DEV))
    a = torch.zeros(128, 128, dtype=torch.bfloat16, device=DEV)      # N % 128 != 0
    w = torch.zeros(64, 128, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(TanError, match="TAN_ERR_SHAPE"):
        ops.linear(a, w, out_f32=torch.empty(128, 64, device=DEV))


@pytest.mark.parametrize("d", [128, 512, 768])
@pytest.mark.parametrize("in_bf16", [False, True])
def test_layernorm_basic(d, in_bf16):
    ops = _ops()
    rows = 777
    x = _rand(rows, d, scale=3.0, seed=11) + 0.5
    if in_bf16:
        x = x.to(torch.bfloat16)
    g, b = _rand(d, seed=12) * 0.1 + 1, _rand(d, seed=13) * 0.1
    of = torch.empty(rows, d, dtype=torch.float32, device=DEV)
    ob = torch.empty(rows, d, dtype=torch.bfloat16, device=DEV)
    ops.layernorm(x, rows, d, gamma=g, beta=b, out_f32=of, out_bf16=ob)
    ref = torch.nn.functional.layer_norm(x.float(), (d,), g, b, 1e-5)
    assert (of - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item())     # fp32 vs fp32
    assert torch.equal(ob, of.to(torch.bfloat16))


def test_layernorm_scatter_add_and_stage_emission():
    ops = _ops()
    B, T, N, d, S = 3, 10, 4, 512, 2
    L = T + N
    xin = _rand(B * L, d, seed=14)
    g, b = _rand(d, seed=15) * 0.1 + 1, _rand(d, seed=16) * 0.1
    add = _rand(L, d, seed=17)
    out = torch.zeros(B * (L + 3), d, device=DEV)
    rawA = torch.zeros(B, S, T, d, device=DEV)
    rawB = torch.zeros(S, B, N, d, device=DEV)
    nA = torch.zeros(B, S, T, d, dtype=torch.bfloat16, device=DEV)
    nB = torch.zeros(S, B * N, d, dtype=torch.bfloat16, device=DEV)
    nBf = torch.zeros(S, B * N, d, device=DEV)
    s = 1
    ops.layernorm(xin, B * L, d, gamma=g, beta=b, add=add, add_rows=L, L_in=L, L_out=L + 3, l_off=2, out_f32=out,
                  l_split=T, strideA=S * T, strideB=N,
                  rawA=rawA.view(-1, d)[s * T:], rawB=rawB.view(-1, d)[s * B * N:],
                  nrmA_bf16=nA.view(-1, d)[s * T:], nrmB_bf16=nB.view(-1, d)[s * B * N:],
                  nrmB_f32=nBf.view(-1, d)[s * B * N:])
    y = torch.nn.functional.layer_norm(xin, (d,), g, b, 1e-5).view(B, L, d) + add[None]
    got = out.view(B, L + 3, d)
    assert (got[:, 2:2 + L] - y).abs().max().item() < 3e-5
    assert got[:, :2].abs().max().item() == 0 and got[:, 2 + L:].abs().max().item() == 0
    assert (rawA[:, s] - y[:, :T]).abs().max().item() < 3e-5
    assert rawA[:, 0].abs().max().item() == 0
    assert (rawB[s] - y[:, T:]).abs().max().item() < 3e-5
    yn = y / y.norm(dim=-1, keepdim=True)
    assert (nA[:, s].float() - yn[:, :T]).abs().max().item() < 4e-3 * yn.abs().max().item() + 1e-3
    assert (nBf[s].view(B, N, d) - yn[:, T:]).abs().max().item() < 1e-6
    assert torch.equal(nB[s], nBf[s].to(torch.bfloat16))


def _attn_ref(q, k, v, kpm, B, H, Lq, Lk):
    d = H * 64
    qh = q.float().view(B, Lq, H, 64).transpose(1, 2)
    kh = k.float().view(B, Lk, H, 64).transpose(1, 2)
    vh = v.float().view(B, Lk, H, 64).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) / 8.0
    if kpm is not None:
        s = s.masked_fill(kpm.bool()[:, None, None, :], float("-inf"))
    return (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(B * Lq, d)


# (40, 8, 256), (24, 8, 288), (5, 8, 600): more (clip, head, query-tile pair) tasks than SMs, so every persistent CTA
# walks several tasks (pipelines running across task boundaries, pairs with and without a second tile)
@pytest.mark.parametrize("B,H,L", [(2, 8, 32), (3, 8, 72), (2, 8, 256), (2, 12, 288), (1, 8, 1152), (4, 2, 37),
                                   (40, 8, 256), (24, 8, 288), (5, 8, 600), (3, 4, 129)])
@pytest.mark.parametrize("masked", [False, True])
def test_self_attention(B, H, L, masked):
    ops = _ops()
    d = H * 64
    qkv = _rand(B * L, 3 * d, seed=20).to(torch.bfloat16)
    kpm = None
    if masked:
        kpm = torch.zeros(B, L, dtype=torch.uint8, device=DEV)
        kpm[0, L - L // 4:] = 1                      # padded suffix
        kpm[-1, 1::3] = 1                            # arbitrary pattern
    out = torch.empty(B * L, d, dtype=torch.bfloat16, device=DEV)
    ops.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], kpm, out, B, H, L, L)
    ref = _attn_ref(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], kpm, B, H, L, L)
    # bf16 P and bf16 output: abs error ~ 2^-8 of the output scale
    assert (out.float() - ref).abs().max().item() < 2e-2 * max(ref.abs().max().item(), 0.1)
    assert ((out.float() - ref).norm() / ref.norm()).item() < 1e-2


def test_attention_rising_scores_rescale_the_accumulator():
    # keys whose scores grow along the sequence: every later key block exceeds the running reference maximum by far
    # more than the lazy-rescale slack of 2^8, so the O accumulator in TMEM is rescaled at every block
    ops = _ops()
    B, H, L = 20, 8, 520
    d = H * 64
    qkv = _rand(B * L, 3 * d, seed=23)
    ramp = torch.linspace(0.2, 3.0, L, device=DEV).repeat(B)[:, None]
    qkv[:, :d] += 1.0                                                  # a common query direction ...
    qkv[:, d:2 * d] = 3.0 * ramp + 0.3 * qkv[:, d:2 * d]               # ... along which the keys grow: score ~ 24 ramp
    qkv = qkv.to(torch.bfloat16)
    out = torch.empty(B * L, d, dtype=torch.bfloat16, device=DEV)
    lse = torch.empty(B, H, (L + 63) // 64 * 64, dtype=torch.float32, device=DEV)
    ops.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], None, out, B, H, L, L, lse=lse)
    ref = _attn_ref(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], None, B, H, L, L)
    assert ((out.float() - ref).norm() / ref.norm()).item() < 1e-2
    qh = qkv[:, :d].float().view(B, L, H, 64).transpose(1, 2)
    kh = qkv[:, d:2 * d].float().view(B, L, H, 64).transpose(1, 2)
    lse_ref = torch.logsumexp(qh @ kh.transpose(-1, -2) / 8.0, -1) * 1.4426950408889634     # log2 domain
    assert (lse[:, :, :L] - lse_ref).abs().max().item() < 2e-3 * lse_ref.abs().max().item() + 1e-2
    assert torch.isinf(lse[:, :, L:]).all()


def test_cross_attention_and_all_masked_row_is_nan():
    ops = _ops()
    B, H, Lq, Lk = 2, 2, 10, 12
    d = H * 64
    q = _rand(B * Lq, d, seed=21).to(torch.bfloat16)
    kv = _rand(B * Lk, 2 * d, seed=22).to(torch.bfloat16)
    kpm = torch.zeros(B, Lk, dtype=torch.uint8, device=DEV)
    kpm[0, 9:] = 1
    out = torch.empty(B * Lq, d, dtype=torch.bfloat16, device=DEV)
    ops.attention(q, kv[:, :d], kv[:, d:], kpm, out, B, H, Lq, Lk)
    ref = _attn_ref(q, kv[:, :d], kv[:, d:], kpm, B, H, Lq, Lk)
    assert (out.float() - ref).abs().max().item() < 2e-2 * ref.abs().max().item()
    kpm[1, :] = 1                                     # every key masked -> NaN like torch.softmax
    ops.attention(q, kv[:, :d], kv[:, d:], kpm, out, B, H, Lq, Lk)
    assert torch.isnan(out.float()[Lq:]).all() and torch.isfinite(out.float()[:Lq]).all()


def _sim_inputs(B, S, T, N, d, seed, shared):
    g = torch.Generator().manual_seed(seed)
    v = torch.randn(B, S, T, d, generator=g)
    v = (v / v.norm(dim=-1, keepdim=True)).to(DEV).to(torch.bfloat16)
    tshape = (B * N, d) if shared else (S, B * N, d)
    t = torch.randn(*tshape, generator=g)
    t = (t / t.norm(dim=-1, keepdim=True)).to(DEV).to(torch.bfloat16)
    n_b = torch.randint(max(N // 2, 1), N + 1, (B,), generator=g)
    start = torch.full((B, N), T + 100.0)
    end = torch.full((B, N), -100.0)
    valid = torch.zeros(B, N, dtype=torch.uint8)
    for b in range(B):
        nb = int(n_b[b])
        s = torch.sort(torch.randint(0, max(T - 1, 1), (nb,), generator=g)).values.float()
        start[b, :nb] = s
        end[b, :nb] = torch.minimum(s + torch.randint(1, 9, (nb,), generator=g).float(), torch.tensor(float(T)))
        valid[b, :nb] = 1
    return v, t, start.view(-1).to(DEV), end.view(-1).to(DEV), valid.view(-1).to(DEV)


def _posbits(ops, start, end, valid, B, T, N, b_off=0, B_loc=None):
    """Packed target bits of the local clips from tan_pos_from_time, checked against the torch packer."""
    from tests.helpers import cpu_pos_from_time
    B_loc = B if B_loc is None else B_loc
    sl = slice(b_off * N, (b_off + B_loc) * N)
    bits = ops.pos_from_time(start[sl].contiguous(), end[sl].contiguous(), valid[sl].contiguous(), B_loc, T, N)
    ref = cpu_pos_from_time(start[sl].cpu(), end[sl].cpu(), valid[sl].cpu(), B_loc, T, N)
    assert torch.equal(bits.cpu(), ref)
    return bits


def _sim_ref(v, t, start, end, valid, B, S, T, N, shared, b_off=0, B_loc=None, kill=None):
    """fp32 torch reference of the statistics: exp-sums with the fixed shift 1/0.07.  kill [B_loc, T] bool:
    own-clip entries of those frames count as exp(-inf)."""
    B_loc = B if B_loc is None else B_loc
    vf = v.float()                                                  # [B_loc,S,T,d]
    tf = t.float() if not shared else t.float()[None].expand(S, -1, -1)     # [S,C,d]
    cos = torch.einsum("bstd,scd->bstc", vf, tf)                    # [B_loc,S,T,C]
    e = torch.exp((cos - 1.0) / 0.07) * valid.float()[None, None, None, :]
    C = tf.shape[1]
    tt = torch.arange(T, device=v.device).float()
    pos_bt = (start[None, :] <= tt[:, None]) & (tt[:, None] < end[None, :]) & valid.bool()[None, :]    # [T,C]
    own = (torch.arange(C, device=v.device) // N)[None, :] == (b_off + torch.arange(B_loc, device=v.device))[:, None]
    pos = pos_bt[None, :, :] & own[:, None, :]                      # [B_loc,T,C]
    if kill is not None:
        e = e * (~(kill[:, :, None] & own[:, None, :]))[:, None].float()
    pe = e * pos[:, None].float()
    row = torch.stack((e.sum(-1).reshape(-1), pe.sum(-1).reshape(-1)))
    col = torch.stack((e.sum(dim=(0, 2)), pe.sum(dim=(0, 2))))     # [2,S,C]
    return cos, row, col


@pytest.mark.parametrize("B,S,T,N,d,shared", [
    (4, 1, 32, 4, 512, True), (3, 3, 24, 5, 512, False), (2, 6, 64, 8, 512, False),
    (8, 2, 256, 32, 512, True), (8, 2, 256, 32, 512, False), (2, 2, 200, 70, 768, False), (5, 1, 130, 3, 512, True),
])
@pytest.mark.parametrize("store,kill", [(False, False), (True, False), (False, True), (True, True)])
def test_sim_nce_fwd(B, S, T, N, d, shared, store, kill):
    ops = _ops()
    v, t, start, end, valid = _sim_inputs(B, S, T, N, d, 30, shared)
    C = B * N
    g = ops.sim_geom(B, S, T, C, N, d, 0)
    rs = torch.empty(2, B * S * T, device=DEV)
    cs = torch.empty(2, S, C, device=DEV)
    ws = torch.empty(ops.sim_workspace_bytes(g), dtype=torch.uint8, device=DEV)
    logits = torch.full((B, S, T, B, N), float("nan"), dtype=torch.bfloat16, device=DEV) if store else None
    posbits = _posbits(ops, start, end, valid, B, T, N)
    km = None
    if kill:                                         # padded suffix on every other clip
        km = torch.zeros(B, T, dtype=torch.bool, device=DEV)
        km[::2, T - max(T // 4, 1):] = True
    ops.sim_nce_fwd(v, t, 0 if shared else C * d, g, posbits, valid, logits, rs, cs, ws,
                    row_kill=km.to(torch.uint8) if kill else None)
    cos, row, col = _sim_ref(v, t, start, end, valid, B, S, T, N, shared, kill=km)
    # ex2.approx on |z| <= 14.3: relative error ~1e-6 per term; sums of positives
    assert ((rs - row).abs() / row.clamp_min(1e-30)).max().item() < 1e-3
    assert ((cs - col).abs() / col.clamp_min(1e-30)).max().item() < 1e-3
    assert ((rs[1] > 0) == (row[1] > 0)).all() and ((cs[1] > 0) == (col[1] > 0)).all()
    if store:
        assert torch.equal(logits.view(B, S, T, C), cos.to(torch.bfloat16)) or \
            (logits.view(B, S, T, C).float() - cos).abs().max().item() < 8e-3    # bf16 rounding of fp32 accum


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,S,T,N", [(4, 1, 32, 4), (3, 3, 24, 5), (8, 2, 256, 32), (2, 1, 100, 150)])
def test_nce_from_logits(B, S, T, N, dtype):
    ops = _ops()
    d = 128
    v, t, start, end, valid = _sim_inputs(B, S, T, N, d, 31, False)
    C = B * N
    cos, _, _ = _sim_ref(v, t, start, end, valid, B, S, T, N, False)
    logits = cos.to(dtype).contiguous()
    # reference statistics from the ROUNDED logits the kernel actually sees
    e = torch.exp((logits.float() - 1.0) / 0.07) * valid.float()
    tt = torch.arange(T, device=DEV).float()
    pos_bt = (start[None, :] <= tt[:, None]) & (tt[:, None] < end[None, :]) & valid.bool()[None, :]
    own = (torch.arange(C, device=DEV) // N)[None, :] == torch.arange(B, device=DEV)[:, None]
    pe = e * (pos_bt[None] & own[:, None, :])[:, None].float()
    row = torch.stack((e.sum(-1).reshape(-1), pe.sum(-1).reshape(-1)))
    col = torch.stack((e.sum(dim=(0, 2)), pe.sum(dim=(0, 2))))
    g = ops.sim_geom(B, S, T, C, N, d, 0)
    rs = torch.empty(2, B * S * T, device=DEV)
    cs = torch.empty(2, S, C, device=DEV)
    ws = torch.empty(ops.sim_workspace_bytes(g), dtype=torch.uint8, device=DEV)
    posbits = _posbits(ops, start, end, valid, B, T, N)
    ops.nce_from_logits(logits.view(B, S, T, B, N), g, posbits, valid, rs, cs, ws)
    assert ((rs - row).abs() / row.clamp_min(1e-30)).max().item() < 1e-4
    assert ((cs - col).abs() / col.clamp_min(1e-30)).max().item() < 1e-4
    out4 = torch.zeros(4, dtype=torch.float64, device=DEV)
    ops.nce_reduce(rs, cs, out4, S, T, C)
    rm, cm = row[1] > 0, col[1].reshape(-1) > 0
    ref_v = (row[0][rm].log() - row[1][rm].log()).double().sum()
    ref_t = (col[0].reshape(-1)[cm].log() - col[1].reshape(-1)[cm].log()).double().sum()
    assert abs(out4[1].item() - rm.sum().item()) == 0 and abs(out4[3].item() - cm.sum().item()) == 0
    assert abs(out4[0].item() - ref_v.item()) < 1e-4 * abs(ref_v.item())
    assert abs(out4[2].item() - ref_t.item()) < 1e-4 * abs(ref_t.item())
    # row / column selections of the thresholded loss (train/loss.py:277-304)
    gsel = torch.Generator().manual_seed(7)
    row_sel = (torch.rand(B, T, generator=gsel) < 0.5).to(DEV)
    col_sel = (torch.rand(C, generator=gsel) < 0.5).to(DEV)
    out4.zero_()
    ops.nce_reduce(rs, cs, out4, S, T, C, row_sel.to(torch.uint8), col_sel.to(torch.uint8))
    rm2 = rm & row_sel[:, None, :].expand(B, S, T).reshape(-1)
    cm2 = cm & col_sel[None, :].expand(S, C).reshape(-1)
    ref_v2 = (row[0][rm2].log() - row[1][rm2].log()).double().sum()
    ref_t2 = (col[0].reshape(-1)[cm2].log() - col[1].reshape(-1)[cm2].log()).double().sum()
    assert out4[1].item() == rm2.sum().item() and out4[3].item() == cm2.sum().item()
    assert abs(out4[0].item() - ref_v2.item()) <= 1e-4 * abs(ref_v2.item()) + 1e-9
    assert abs(out4[2].item() - ref_t2.item()) <= 1e-4 * abs(ref_t2.item()) + 1e-9


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_nce_from_logits_killed_rows_and_many_clips(dtype):
    """Row-kill (the reference's -6e4 fill of padded frames) and the multi-clip-per-CTA walk."""
    ops = _ops()
    B, S, T, N, d = 40, 2, 72, 8, 128
    v, t, start, end, valid = _sim_inputs(B, S, T, N, d, 33, False)
    C = B * N
    km = torch.zeros(B, T, dtype=torch.bool, device=DEV)
    km[1::3, T - 20:] = True
    cos, _, _ = _sim_ref(v, t, start, end, valid, B, S, T, N, False)
    logits = cos.to(dtype).contiguous()
    e = torch.exp((logits.float() - 1.0) / 0.07) * valid.float()
    tt = torch.arange(T, device=DEV).float()
    pos_bt = (start[None, :] <= tt[:, None]) & (tt[:, None] < end[None, :]) & valid.bool()[None, :]
    own = (torch.arange(C, device=DEV) // N)[None, :] == torch.arange(B, device=DEV)[:, None]
    e = e * (~(km[:, :, None] & own[:, None, :]))[:, None].float()
    pe = e * (pos_bt[None] & own[:, None, :])[:, None].float()
    row = torch.stack((e.sum(-1).reshape(-1), pe.sum(-1).reshape(-1)))
    col = torch.stack((e.sum(dim=(0, 2)), pe.sum(dim=(0, 2))))
    g = ops.sim_geom(B, S, T, C, N, d, 0)
    rs = torch.empty(2, B * S * T, device=DEV)
    cs = torch.empty(2, S, C, device=DEV)
    ws = torch.empty(ops.sim_workspace_bytes(g), dtype=torch.uint8, device=DEV)
    posbits = _posbits(ops, start, end, valid, B, T, N)
    ops.nce_from_logits(logits.view(B, S, T, B, N), g, posbits, valid, rs, cs, ws, row_kill=km.to(torch.uint8))
    assert ((rs - row).abs() / row.clamp_min(1e-30)).max().item() < 1e-4
    assert ((cs - col).abs() / col.clamp_min(1e-30)).max().item() < 1e-4
    assert ((rs[1] > 0) == (row[1] > 0)).all() and ((cs[1] > 0) == (col[1] > 0)).all()


@pytest.mark.parametrize("B_loc,b_off,B_glob,S,T,N,shared", [
    (2, 17, 40, 2, 200, 32, False),      # few row blocks, 5 column tiles: the fused kernel splits the column sweep
    (3, 0, 24, 1, 520, 64, True),        # 3 row blocks per segment, ragged last one; N = 64 (two target words)
    (1, 9, 10, 3, 64, 100, False),       # sentences straddle 32-column chunks and 256-column tiles
])
def test_sim_nce_fwd_local_rows_global_columns(B_loc, b_off, B_glob, S, T, N, shared):
    """The multi-GPU geometry: local rows x global columns (b_off > 0), column chunking of the resident-row
    kernel, own-clip blocks in the middle of the column range."""
    ops = _ops()
    d = 512
    v, t, start, end, valid = _sim_inputs(B_glob, S, T, N, d, 41, shared)
    v = v[b_off:b_off + B_loc].contiguous()
    C = B_glob * N
    g = ops.sim_geom(B_loc, S, T, C, N, d, b_off)
    rs = torch.empty(2, B_loc * S * T, device=DEV)
    cs = torch.empty(2, S, C, device=DEV)
    ws = torch.empty(ops.sim_workspace_bytes(g), dtype=torch.uint8, device=DEV)
    posbits = _posbits(ops, start, end, valid, B_glob, T, N, b_off, B_loc)
    ops.sim_nce_fwd(v, t, 0 if shared else C * d, g, posbits, valid, None, rs, cs, ws)
    _, row, col = _sim_ref(v, t, start, end, valid, B_glob, S, T, N, shared, b_off=b_off, B_loc=B_loc)
    assert ((rs - row).abs() / row.clamp_min(1e-30)).max().item() < 1e-3
    assert ((cs - col).abs() / col.clamp_min(1e-30)).max().item() < 1e-3
    assert ((rs[1] > 0) == (row[1] > 0)).all() and ((cs[1] > 0) == (col[1] > 0)).all()


def test_sim_sharded_rows_add_up():
    """Column sums are additive over row shards (the multi-GPU exchange relies on it)."""
    ops = _ops()
    B, S, T, N, d = 4, 2, 64, 8, 512
    v, t, start, end, valid = _sim_inputs(B, S, T, N, d, 32, False)
    C = B * N

    def run(vv, b_off, B_loc):
        g = ops.sim_geom(B_loc, S, T, C, N, d, b_off)
        rs = torch.empty(2, B_loc * S * T, device=DEV)
        cs = torch.empty(2, S, C, device=DEV)
        ws = torch.empty(ops.sim_workspace_bytes(g), dtype=torch.uint8, device=DEV)
        posbits = _posbits(ops, start, end, valid, B, T, N, b_off, B_loc)
        ops.sim_nce_fwd(vv.contiguous(), t, C * d, g, posbits, valid, None, rs, cs, ws)
        return rs, cs

    rs_full, cs_full = run(v, 0, B)
    rs0, cs0 = run(v[:2], 0, 2)
    rs1, cs1 = run(v[2:], 2, 2)
    assert torch.allclose(cs0 + cs1, cs_full, rtol=1e-5, atol=0)
    assert torch.allclose(torch.cat((rs0.view(2, 2, -1), rs1.view(2, 2, -1)), 1).reshape(2, -1), rs_full, rtol=1e-6)
