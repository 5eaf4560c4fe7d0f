"""Per-kernel numerics tests of the backward-pass entry points (GPU): each against torch autograd / a plain
fp32 torch formula of the same op on the same (bf16-rounded) inputs.  Tolerances next to each check."""
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


@pytest.mark.parametrize("R,C", [(64, 64), (144, 512), (1000, 1536), (130, 2), (4096, 2048)])
def test_transpose(R, C):
    ops = _ops()
    x = _rand(R, C, seed=1).to(torch.bfloat16)
    out = torch.full((C, ops.pad64(R)), 7.0, dtype=torch.bfloat16, device=DEV)
    ops.transpose_bf16(x, out)
    assert torch.equal(out[:, :R], x.t())                       # bit exact
    assert (out[:, R:] == 0).all()                              # contraction padding is zero


def test_transpose_pitched_input():
    ops = _ops()
    big = _rand(300, 1536, seed=2).to(torch.bfloat16)
    x = big[:, 512:1024]                                        # column slice: row pitch 1536
    out = ops.transpose_bf16(x)
    assert torch.equal(out[:, :300], x.t())


@pytest.mark.parametrize("M,N,bf", [(1000, 512, True), (33, 1536, True), (70000, 2048, True), (513, 512, False)])
def test_colsum(M, N, bf):
    ops = _ops()
    x = _rand(M, N, seed=3)
    if bf:
        x = x.to(torch.bfloat16)
    out = torch.ones(N, dtype=torch.float32, device=DEV)
    ops.colsum(x, out, accumulate=True)
    ref = 1.0 + x.double().sum(0)
    assert (out.double() - ref).abs().max().item() < 1e-3 * max(1.0, ref.abs().max().item())
    ops.colsum(x, out, accumulate=False)
    assert (out.double() - (ref - 1.0)).abs().max().item() < 1e-3 * max(1.0, ref.abs().max().item())


def test_quickgelu_fwd_bwd():
    ops = _ops()
    u = _rand(1000, 2048, scale=2.0, seed=4).to(torch.bfloat16)
    dh = _rand(1000, 2048, seed=5).to(torch.bfloat16)
    h = torch.empty_like(u)
    du = torch.empty_like(u)
    ops.quickgelu_fwd(u, h)
    ops.quickgelu_bwd(dh, u, du)
    uf = u.float().requires_grad_(True)
    hf = uf * torch.sigmoid(1.702 * uf)
    hf.backward(dh.float())
    # bf16 output rounding (2^-9 relative) + fast exp
    assert (h.float() - hf).abs().max().item() < 1e-2 * max(1.0, hf.abs().max().item())
    assert (du.float() - uf.grad).abs().max().item() < 1e-2 * max(1.0, uf.grad.abs().max().item())
    assert ((h.float() - hf).norm() / hf.norm()).item() < 3e-3
    assert ((du.float() - uf.grad).norm() / uf.grad.norm()).item() < 3e-3


@pytest.mark.parametrize("rows,d,accumulate", [(1000, 512, True), (37, 768, False), (5000, 512, False)])
def test_layernorm_bwd(rows, d, accumulate):
    ops = _ops()
    x = _rand(rows, d, scale=2.0, seed=6) + 0.5
    dy = _rand(rows, d, seed=7)
    gamma = _rand(d, seed=8) * 0.2 + 1.0
    beta = _rand(d, seed=9)
    dx0 = _rand(rows, d, seed=10)
    dx = dx0.clone()
    dgamma = torch.full((d,), 0.5, dtype=torch.float32, device=DEV)
    dbeta = torch.full((d,), -0.5, dtype=torch.float32, device=DEV)
    ops.layernorm_bwd(dy, x, gamma, dx, accumulate, rows, d, dgamma, dbeta)
    xr = x.clone().requires_grad_(True)
    gr = gamma.clone().requires_grad_(True)
    br = beta.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (d,), gr, br, 1e-5).backward(dy)
    ref_dx = xr.grad + (dx0 if accumulate else 0)
    assert (dx - ref_dx).abs().max().item() < 1e-4 * max(1.0, ref_dx.abs().max().item())
    assert (dgamma - 0.5 - gr.grad).abs().max().item() < 1e-3 * max(1.0, gr.grad.abs().max().item())
    assert (dbeta + 0.5 - br.grad).abs().max().item() < 1e-3 * max(1.0, br.grad.abs().max().item())
    # fused outputs: bf16 copy of the updated dx and its column sums (accumulated into the given vector)
    dx2 = dx0.clone()
    dxb = torch.zeros(rows, d, dtype=torch.bfloat16, device=DEV)
    csum = torch.full((d,), 2.0, dtype=torch.float32, device=DEV)
    dg2, db2 = torch.zeros(d, device=DEV), torch.zeros(d, device=DEV)
    ops.layernorm_bwd(dy, x, gamma, dx2, accumulate, rows, d, dg2, db2, dx_bf16=dxb, dx_colsum=csum)
    assert torch.equal(dx2, dx)
    assert torch.equal(dxb, dx.to(torch.bfloat16))
    ref_c = 2.0 + dx.double().sum(0)
    assert (csum.double() - ref_c).abs().max().item() < 1e-3 * max(1.0, ref_c.abs().max().item())


def test_layernorm_bwd_row_map():
    """dy lives in the concatenated [B, T+N, d] token buffer (joint stack): rows l_off .. l_off+L_in of each clip."""
    ops = _ops()
    B, T, N, d = 3, 20, 4, 512
    L = T + N
    x = _rand(B * N, d, seed=11)
    dy_full = _rand(B * L, d, seed=12)
    gamma = _rand(d, seed=13) * 0.1 + 1.0
    dx = torch.empty(B * N, d, dtype=torch.float32, device=DEV)
    dgamma = torch.zeros(d, dtype=torch.float32, device=DEV)
    dbeta = torch.zeros(d, dtype=torch.float32, device=DEV)
    ops.layernorm_bwd(dy_full, x, gamma, dx, False, B * N, d, dgamma, dbeta, L_in=N, L_out=L, l_off=T)
    dy = dy_full.view(B, L, d)[:, T:].reshape(B * N, d)
    xr = x.clone().requires_grad_(True)
    gr = gamma.clone().requires_grad_(True)
    br = torch.zeros(d, device=DEV, requires_grad=True)
    torch.nn.functional.layer_norm(xr, (d,), gr, br, 1e-5).backward(dy)
    assert (dx - xr.grad).abs().max().item() < 1e-4 * max(1.0, xr.grad.abs().max().item())
    assert (dgamma - gr.grad).abs().max().item() < 1e-3 * max(1.0, gr.grad.abs().max().item())
    assert (dbeta - br.grad).abs().max().item() < 1e-3 * max(1.0, br.grad.abs().max().item())


def test_l2norm_bwd_and_batch_sum():
    ops = _ops()
    B, S, T, N, d = 3, 2, 10, 4, 512
    L = T + N
    x = _rand(B, S, T, d, seed=14)                              # raw stage features, layout [B, S, T, d]
    g = _rand(B, S, T, d, seed=15)
    s = 1
    dst0 = _rand(B * L, d, seed=16)
    dst = dst0.clone()
    ops.l2norm_bwd(x.view(-1, d)[s * T:], g.view(-1, d)[s * T:], dst, False, B * T, d, T, S * T, L, 0)
    xr = x[:, s].clone().requires_grad_(True)
    (xr / xr.norm(dim=-1, keepdim=True)).backward(g[:, s])
    got = dst.view(B, L, d)[:, :T]
    assert (got - xr.grad).abs().max().item() < 1e-5 * max(1.0, xr.grad.abs().max().item())
    assert torch.equal(dst.view(B, L, d)[:, T:], dst0.view(B, L, d)[:, T:])     # other rows untouched
    dst2 = dst0.clone()
    ops.l2norm_bwd(x.view(-1, d)[s * T:], g.view(-1, d)[s * T:], dst2, True, B * T, d, T, S * T, L, 0)
    assert (dst2.view(B, L, d)[:, :T] - (dst0.view(B, L, d)[:, :T] + xr.grad)).abs().max().item() < 1e-5
    # gradient stored stage-major ([S, B, T, d]: clip stride T) while x keeps the [B, S, T, d] layout
    g_sm = g.permute(1, 0, 2, 3).contiguous()
    dst3 = torch.empty(B * L, d, dtype=torch.float32, device=DEV)
    ops.l2norm_bwd(x.view(-1, d)[s * T:], g_sm[s].view(-1, d), dst3, False, B * T, d, T, S * T, L, 0, g_stride=T)
    assert (dst3.view(B, L, d)[:, :T] - xr.grad).abs().max().item() < 1e-5 * max(1.0, xr.grad.abs().max().item())
    # batch sum over the clips of rows [T, T+N) of every clip
    out = torch.ones(N, d, dtype=torch.float32, device=DEV)
    ops.batch_sum(dst0, out, B, N, d, L, T, True)
    ref = 1.0 + dst0.view(B, L, d)[:, T:].sum(0)
    assert (out - ref).abs().max().item() < 1e-5


def _attn_ref(q, k, v, kpm, dO):
    """fp32 torch attention on [B, H, L, 64] + autograd."""
    q, k, v = (t.clone().requires_grad_(True) for t in (q, k, v))
    s = (q @ k.transpose(-1, -2)) * 0.125
    if kpm is not None:
        s = s.masked_fill(kpm[:, None, None, :].bool(), float("-inf"))
    o = torch.softmax(s, dim=-1) @ v
    o.backward(dO)
    return o.detach(), q.grad, k.grad, v.grad


@pytest.mark.parametrize("B,H,Lq,Lk,masked", [(2, 8, 64, 64, False), (3, 8, 36, 36, True), (2, 12, 288, 288, True),
                                              (2, 8, 100, 160, True), (1, 8, 256, 256, False)])
def test_attention_bwd(B, H, Lq, Lk, masked):
    ops = _ops()
    d = H * 64
    qkv = (_rand(B * Lq, 3 * d, seed=20)).to(torch.bfloat16)
    kv = (_rand(B * Lk, 3 * d, seed=21)).to(torch.bfloat16) if Lk != Lq else qkv
    q2, k2, v2 = qkv[:, :d], kv[:, d:2 * d], kv[:, 2 * d:]
    dO2 = _rand(B * Lq, d, seed=22).to(torch.bfloat16)
    kpm = None
    if masked:
        kpm = torch.zeros(B, Lk, dtype=torch.uint8, device=DEV)
        for b in range(B):
            kpm[b, Lk - 3 - 5 * b:] = 1                       # padded suffix of different lengths
        kpm[0, 1] = 1                                          # and a hole

    def heads(x2, L):
        return x2.float().view(B, L, H, 64).permute(0, 2, 1, 3)

    o_ref, dq_ref, dk_ref, dv_ref = _attn_ref(heads(q2, Lq), heads(k2, Lk), heads(v2, Lk), kpm, heads(dO2, Lq))
    # the forward kernel provides o and the rows' log2-domain log-sum-exp ([B, H, pad64(Lq)], +inf padding)
    Lp = ops.pad64(Lq)
    o2 = torch.empty(B * Lq, d, dtype=torch.bfloat16, device=DEV)
    lse = torch.full((B, H, Lp), 7.0, dtype=torch.float32, device=DEV)
    ops.attention(q2, k2, v2, kpm, o2, B, H, Lq, Lk, lse=lse)
    o_chk = o_ref.permute(0, 2, 1, 3).reshape(B * Lq, d)
    assert ((o2.float() - o_chk).norm() / o_chk.norm()).item() < 1e-2
    dqkv = torch.zeros(B * Lq, d, dtype=torch.bfloat16, device=DEV)
    dkv = torch.zeros(B * Lk, 2 * d, dtype=torch.bfloat16, device=DEV)
    delta = torch.empty(B, H, Lp, dtype=torch.float32, device=DEV)
    ops.attention_bwd(q2, k2, v2, o2, dO2, kpm, dqkv, dkv[:, :d], dkv[:, d:], lse, delta, B, H, Lq, Lk)
    torch.cuda.synchronize()

    def flat(x4, L):
        return x4.permute(0, 2, 1, 3).reshape(B * L, d)

    s = (heads(q2, Lq) @ heads(k2, Lk).transpose(-1, -2)) * 0.125
    if kpm is not None:
        s = s.masked_fill(kpm[:, None, None, :].bool(), float("-inf"))
    assert (lse[:, :, :Lq] * 0.6931471805599453 - torch.logsumexp(s, -1)).abs().max().item() < 2e-3
    assert torch.isinf(lse[:, :, Lq:]).all() and (lse[:, :, Lq:] > 0).all()
    d_ref = (heads(dO2, Lq) * heads(o2, Lq)).sum(-1)
    assert (delta[:, :, :Lq] - d_ref).abs().max().item() < 1e-2 and (delta[:, :, Lq:] == 0).all()
    for got, ref, L in ((dqkv, dq_ref, Lq), (dkv[:, :d], dk_ref, Lk), (dkv[:, d:], dv_ref, Lk)):
        ref2 = flat(ref, L)
        err = ((got.float() - ref2).norm() / ref2.norm()).item()
        assert err < 1.5e-2, err                               # bf16 P / dS operands and bf16 outputs
        assert torch.isfinite(got.float()).all()


@pytest.mark.parametrize("with_kill", [False, True])
def test_sim_grad_tiles_and_gemms(with_kill):
    """dA = G @ B, dB = G^T @ A through tan_sim_grad_tiles + tan_linear_bf16 against autograd of the closed-form
    MIL-NCE loss (train/loss.py:231-275) on the same bf16 features, one stage.  with_kill: some frames lose their
    own-clip entries (row_kill, the -6e4 fill of padded frames under --learn_agreement)."""
    ops = _ops()
    B, T, N, d = 4, 40, 6, 512
    C = B * N
    Cp = (C + 127) // 128 * 128
    R = B * T
    a = torch.nn.functional.normalize(_rand(R, d, seed=30), dim=-1).to(torch.bfloat16)
    t = torch.nn.functional.normalize(_rand(C, d, seed=31), dim=-1).to(torch.bfloat16)
    gen = torch.Generator().manual_seed(5)
    valid = torch.ones(B, N, dtype=torch.bool)
    valid[1, 4:] = False
    valid[3, 5:] = False
    start = torch.randint(0, T - 8, (B, N), generator=gen)
    end = start + torch.randint(1, 8, (B, N), generator=gen)
    tt = torch.arange(T)[None, None, :]
    mask = (start[:, :, None] <= tt) & (tt < end[:, :, None]) & valid[:, :, None]        # [B, N, T]
    from tests.helpers import pack_posbits
    posbits = pack_posbits(mask).to(DEV)
    col_valid = valid.view(-1).to(torch.uint8).to(DEV)

    # reference: closed-form loss on fp32 cosines with autograd
    af = a.float().requires_grad_(True)
    tf = t.float().requires_grad_(True)
    z = (af @ tf.t()) / 0.07
    pos = torch.zeros(R, C, dtype=torch.bool)
    for b in range(B):
        pos[b * T:(b + 1) * T, b * N:(b + 1) * N] = mask[b].t()
    pos = pos.to(DEV)
    cv = col_valid.bool()
    zv = z.masked_fill(~cv[None, :], float("-inf"))
    row_kill = None
    if with_kill:
        kill = torch.zeros(B, T, dtype=torch.bool)
        kill[1, T - 6:] = True
        kill[2, T - 3:] = True
        mask = mask & ~kill[:, None, :]                           # killed frames carry no positives
        posbits = pack_posbits(mask).to(DEV)
        pos = torch.zeros(R, C, dtype=torch.bool)
        for b in range(B):
            pos[b * T:(b + 1) * T, b * N:(b + 1) * N] = mask[b].t()
        pos = pos.to(DEV)
        own = torch.zeros(R, C, dtype=torch.bool)
        for b in range(B):
            own[b * T:(b + 1) * T, b * N:(b + 1) * N] = True
        zv = zv.masked_fill((own & kill.view(-1)[:, None]).to(DEV), float("-inf"))
        row_kill = kill.to(torch.uint8).to(DEV).contiguous()
    row_all = torch.logsumexp(zv, 1)
    row_pos = torch.logsumexp(zv.masked_fill(~pos, float("-inf")), 1)
    rsel = pos.any(1)
    col_all = torch.logsumexp(zv, 0)
    col_pos = torch.logsumexp(zv.masked_fill(~pos, float("-inf")), 0)
    csel = pos.any(0) & cv
    loss = 0.5 * ((row_all - row_pos)[rsel].mean() + (col_all - col_pos)[csel].mean())
    loss.backward()

    # ours: coefficient vectors from the same sums (fixed shift 1/0.07)
    with torch.no_grad():
        e = torch.exp(zv - 1.0 / 0.07)
        ra_sum, rp_sum = e.sum(1), (e * pos).sum(1)
        ca_sum, cp_sum = e.sum(0), (e * pos).sum(0)
        wr, wc = 0.5 / rsel.sum(), 0.5 / csel.sum()
        ra = torch.where(rsel, wr / ra_sum, torch.zeros_like(ra_sum)).contiguous()
        rap = torch.where(rsel, wr / rp_sum.clamp_min(1e-30), torch.zeros_like(ra_sum)).contiguous()
        cb = torch.where(csel, wc / ca_sum.clamp_min(1e-30), torch.zeros_like(ca_sum)).contiguous()
        cbp = torch.where(csel, wc / cp_sum.clamp_min(1e-30), torch.zeros_like(ca_sum)).contiguous()
    g = ops.sim_geom(B, 1, T, C, N, d, 0)
    tpad = torch.zeros(Cp, d, dtype=torch.bfloat16, device=DEV)
    tpad[:C] = t
    zbuf = torch.empty(R, Cp, dtype=torch.float32, device=DEV)
    ops.linear(a, tpad, out_f32=zbuf)
    G = torch.empty(R, Cp, dtype=torch.bfloat16, device=DEV)
    GT = torch.empty(C, ops.pad64(R), dtype=torch.bfloat16, device=DEV)
    ops.sim_grad_tiles(zbuf, R, 0, g, posbits, col_valid, row_kill, ra, rap, cb, cbp, G, GT)
    assert torch.equal(GT[:, :R], G[:, :C].t())
    assert (GT[:, R:] == 0).all() and (G[:, C:] == 0).all()
    # the fused variant: G in the epilogue of the recomputation GEMM (no fp32 cosines in HBM)
    G2 = torch.full((R, Cp), 3.0, dtype=torch.bfloat16, device=DEV)
    ops.sim_grad_gemm(a, tpad, 0, g, posbits, col_valid, row_kill, ra, rap, cb, cbp, G2)
    torch.cuda.synchronize()
    assert (G2[:, C:] == 0).all()
    assert ((G2.float() - G.float()).norm() / G.float().norm()).item() < 1e-2
    assert (G2.float() - G.float()).abs().max().item() < 2e-2 * G.float().abs().max().item()
    tT = ops.transpose_bf16(tpad)                               # [d, Cp]
    aT = ops.transpose_bf16(a)                                  # [d, pad64(R)]
    dA = torch.empty(R, d, dtype=torch.float32, device=DEV)
    dB = torch.empty(C, d, dtype=torch.float32, device=DEV)
    ops.linear(G, tT, out_f32=dA)
    ops.linear(GT, aT, out_f32=dB)
    dB_tn = torch.zeros(Cp, d, dtype=torch.float32, device=DEV)               # the product's route: G^T @ a straight
    ops.gemm_tn(G, a, dB_tn, accumulate=True)                                  # from G (MN-major operands)
    torch.cuda.synchronize()
    assert (dB_tn[C:] == 0).all()
    for got, ref in ((dA, af.grad), (dB, tf.grad), (dB_tn[:C], tf.grad)):
        err = ((got - ref).norm() / ref.norm()).item()
        assert err < 1e-2, err                                  # G is bf16 (2^-9 per element, averaged)


@pytest.mark.parametrize("R,P,Q,acc", [(64, 64, 64, False), (1000, 512, 512, True), (4096, 1536, 512, False),
                                       (8200, 512, 2048, True), (300, 24, 128, True), (20000, 512, 1024, False),
                                       (8192, 1024, 512, True)])
def test_gemm_tn(R, P, Q, acc):
    """out (+)= a^T @ b on MN-major operands (no transposes) vs torch fp32 matmul of the same bf16 inputs."""
    ops = _ops()
    a = _rand(R, P, seed=31).to(torch.bfloat16)
    b = _rand(R, Q, seed=32).to(torch.bfloat16)
    out0 = _rand(P, Q, seed=33)
    out = out0.clone()
    ops.gemm_tn(a, b, out, accumulate=acc)
    ref = a.float().t() @ b.float() + (out0 if acc else 0)
    err = (out - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), err
    # deterministic: same bits on a second run
    out2 = out0.clone()
    ops.gemm_tn(a, b, out2, accumulate=acc)
    assert torch.equal(out, out2)


def test_gemm_tn_pitched_operands():
    ops = _ops()
    big = _rand(3000, 1536, seed=34).to(torch.bfloat16)
    a, b = big[:, 512:1024], big[:, 1024:1536]
    out = torch.zeros(512, 512, dtype=torch.float32, device=DEV)
    ops.gemm_tn(a, b, out, accumulate=False)
    ref = a.float().t() @ b.float()
    assert (out - ref).abs().max().item() < 2e-3 * ref.abs().max().item()


@pytest.mark.parametrize("M,N,K", [(300, 2048, 512), (4096, 2048, 512), (1000, 512, 128)])
def test_linear_dual_and_gelu_bwd_epilogues(M, N, K):
    """tan_linear_dual_bf16 (pre-activation + QuickGELU from one GEMM) and tan_linear_gelu_bwd_bf16
    ((A W^T) o gelu'(u)) against torch on the same bf16 inputs."""
    ops = _ops()
    a = _rand(M, K, seed=51).to(torch.bfloat16)
    w = (_rand(N, K, seed=52) * 0.05).to(torch.bfloat16)
    b = _rand(N, seed=53) * 0.1
    act = torch.full((M, N), 7.0, dtype=torch.bfloat16, device=DEV)
    pre = torch.full((M, N), 7.0, dtype=torch.bfloat16, device=DEV)
    ops.linear_dual(a, w, b, act, pre)
    y = a.float() @ w.float().t() + b
    assert ((pre.float() - y).norm() / y.norm()).item() < 4e-3
    h = y * torch.sigmoid(1.702 * y)
    assert ((act.float() - h).norm() / h.norm()).item() < 5e-3
    # backward: dh = g @ w2^T, du = dh * gelu'(u)
    g = _rand(M, K, seed=54).to(torch.bfloat16)
    du = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    ops.linear_gelu_bwd(g, w, pre, du)
    x = pre.float()
    s = torch.sigmoid(1.702 * x)
    ref = (g.float() @ w.float().t()) * s * (1 + 1.702 * x * (1 - s))
    assert ((du.float() - ref).norm() / ref.norm()).item() < 6e-3
