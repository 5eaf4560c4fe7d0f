"""CPU stand-ins for the C-ABI wrappers of `temporalalignnet_b200.ops` (TEST INFRASTRUCTURE).

Each function has the signature of its `ops` counterpart and the same contract (same output buffers, row maps,
bf16 rounding of bf16 outputs), computed with plain torch on CPU tensors.  `install(monkeypatch)` swaps them in so
that the HOST logic of the product (kernel sequence, buffer layouts, stage-gradient injection, row maps, positional
table slices, the autograd hook) can be exercised by `-m "not gpu"` tests against the oracle.  They are not a
fallback: the product never imports this module and keeps raising on CPU tensors.
"""
import torch

from tests.helpers import cpu_pos_from_time, unpack_posbits

BF = torch.bfloat16


def _rows(t):
    return t if t.dim() == 2 else t.reshape(-1, t.shape[-1])


def cast_bf16(x, out=None):
    y = x.to(BF)
    if out is None:
        return y
    out.copy_(y)
    return out


def linear(a, w, bias=None, residual=None, out_f32=None, out_bf16=None, act=0, tag="linear"):
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias.float()
    if act == 1:
        y = y * torch.sigmoid(1.702 * y)
    if residual is not None:
        y = y + residual
    if out_f32 is not None:
        out_f32.copy_(y)
    if out_bf16 is not None:
        out_bf16.copy_(y.to(BF))


def layernorm(x, rows, d, gamma=None, beta=None, add=None, add_rows=0, L_in=None, L_out=None, l_off=0,
              out_f32=None, out_bf16=None, l_split=0, strideA=0, strideB=0, rawA=None, rawB=None, nrmA_bf16=None,
              nrmB_bf16=None, nrmA_f32=None, nrmB_f32=None, raw_strideA=0, raw_strideB=0):
    L_in = rows if L_in is None else L_in
    L_out = L_in if L_out is None else L_out
    y = _rows(x)[:rows].float()
    if gamma is not None:
        y = torch.nn.functional.layer_norm(y, (d,), gamma.float(), beta.float(), 1e-5)
    r = torch.arange(rows)
    b, l = r // L_in, r % L_in
    if add is not None:
        y = y + _rows(add)[l % add_rows].float()
    dst = b * L_out + l_off + l
    if out_f32 is not None:
        _rows(out_f32)[dst] = y
    if out_bf16 is not None:
        _rows(out_bf16)[dst] = y.to(BF)
    partA = l < l_split
    for part, raw, nb, nf, stride, off, rstride in ((partA, rawA, nrmA_bf16, nrmA_f32, strideA, 0, raw_strideA),
                                                    (~partA, rawB, nrmB_bf16, nrmB_f32, strideB, l_split, raw_strideB)):
        if raw is None and nb is None and nf is None:
            continue
        idx = part.nonzero().squeeze(1)
        if idx.numel() == 0:
            continue
        srow = b[idx] * stride + (l[idx] - off)
        yy = y[idx]
        if raw is not None:
            _rows(raw)[srow if rstride == 0 else b[idx] * rstride + (l[idx] - off)] = yy
        if nb is not None or nf is not None:
            n = yy / yy.norm(dim=-1, keepdim=True)
            if nf is not None:
                _rows(nf)[srow] = n
            if nb is not None:
                _rows(nb)[srow] = n.to(BF)


def _heads(x2, B, L, H):
    return x2.float().reshape(B, L, H, 64).permute(0, 2, 1, 3)


def _scores(q, k, kpm, B, H, Lq, Lk):
    s = (_heads(q, B, Lq, H) @ _heads(k, B, Lk, H).transpose(-1, -2)) * 0.125
    if kpm is not None:
        s = s.masked_fill(kpm.view(B, 1, 1, Lk).bool(), float("-inf"))
    return s


def attention(q, k, v, kpm_u8, out, B, H, Lq, Lk, lse=None):
    s = _scores(q, k, kpm_u8, B, H, Lq, Lk)
    p = torch.softmax(s, dim=-1)
    if lse is not None:                                  # log2 domain, pitch pad64(Lq), +inf padding
        lse.fill_(float("inf"))
        lse.view(B, H, -1)[:, :, :Lq] = torch.logsumexp(s, -1) * 1.4426950408889634
    o = p @ _heads(v, B, Lk, H)
    out.copy_(o.permute(0, 2, 1, 3).reshape(B * Lq, H * 64).to(BF))


def attention_bwd(q, k, v, o, d_out, kpm_u8, dq, dk, dv, lse, delta, B, H, Lq, Lk):
    s = _scores(q, k, kpm_u8, B, H, Lq, Lk)
    p = torch.softmax(s, dim=-1)
    do = _heads(d_out, B, Lq, H)
    dl = (do * _heads(o, B, Lq, H)).sum(-1, keepdim=True)
    dp = do @ _heads(v, B, Lk, H).transpose(-1, -2)
    ds = p * (dp - dl)

    def flat(x4, L):
        return x4.permute(0, 2, 1, 3).reshape(B * L, H * 64).to(BF)

    dq.copy_(flat(ds @ _heads(k, B, Lk, H) * 0.125, Lq))
    dk.copy_(flat(ds.transpose(-1, -2) @ _heads(q, B, Lq, H) * 0.125, Lk))
    dv.copy_(flat(p.transpose(-1, -2) @ do, Lk))
    delta.view(B, H, -1)[:, :, :Lq] = dl.squeeze(-1)


def quickgelu_fwd(u, h):
    x = u.float()
    h.copy_((x * torch.sigmoid(1.702 * x)).to(BF))


def quickgelu_bwd(dh, u, du):
    x = u.float()
    s = torch.sigmoid(1.702 * x)
    du.copy_((dh.float() * s * (1 + 1.702 * x * (1 - s))).to(BF))


def pad64(n):
    return (n + 63) // 64 * 64


def transpose_bf16(x, out=None):
    R, C = x.shape
    if out is None:
        out = torch.empty(C, pad64(R), dtype=BF)
    out.zero_()
    out[:, :R] = x.t()
    return out


def linear_dual(a, w, bias, out_act, out_pre, act=1):
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias.float()
    out_pre.copy_(y.to(BF))
    out_act.copy_((y * torch.sigmoid(1.702 * y) if act == 1 else y).to(BF))


def linear_gelu_bwd(a, w, u, out):
    y = a.float() @ w.float().t()
    x = u.float()
    s = torch.sigmoid(1.702 * x)
    out.copy_((y * s * (1 + 1.702 * x * (1 - s))).to(BF))


def embed_gather(ids, table, out=None):
    r = table[ids.long()]
    if out is not None:
        out.copy_(r)
        return out
    return r


def ema_update(target_params, online_params, m):
    """Stand-in of optim.ema_update (tan_ema_update): in place on the Parameters, so version counters advance."""
    with torch.no_grad():
        for t, o in zip(target_params, online_params):
            t.copy_(t * m + o.detach() * (1.0 - m))


def gemm_tn(a, b, out, accumulate=True, tag="wgrad"):
    r = a.float().t() @ b.float()
    out.copy_(out + r if accumulate else r)


def colsum(x, out, accumulate=True):
    s = x.float().sum(0)
    out.copy_(out + s if accumulate else s)


def layernorm_bwd(dy, x, gamma, dx, accumulate_dx, rows, d, dgamma, dbeta, L_in=None, L_out=None, l_off=0,
                  dx_bf16=None, dx_colsum=None):
    L_in = rows if L_in is None else L_in
    L_out = L_in if L_out is None else L_out
    r = torch.arange(rows)
    g = _rows(dy)[(r // L_in) * L_out + l_off + r % L_in].float()
    xr = _rows(x)[:rows].detach().clone().requires_grad_(True)
    gm = gamma.detach().clone().requires_grad_(True)
    bt = torch.zeros(d, requires_grad=True)
    with torch.enable_grad():
        torch.nn.functional.layer_norm(xr, (d,), gm, bt, 1e-5).backward(g)
    tgt = _rows(dx)[:rows]
    tgt.copy_(tgt + xr.grad if accumulate_dx else xr.grad)
    if dgamma is not None:
        dgamma.add_(gm.grad)
        dbeta.add_(bt.grad)
    if dx_bf16 is not None:
        _rows(dx_bf16)[:rows].copy_(tgt.to(BF))
    if dx_colsum is not None:
        dx_colsum.add_(tgt.float().sum(0))


def l2norm_bwd(x, g, dst, accumulate, rows, d, L_in, src_stride, L_out, l_off, g_stride=None):
    g_stride = src_stride if g_stride is None else g_stride
    r = torch.arange(rows)
    b, l = r // L_in, r % L_in
    xv = _rows(x)[b * src_stride + l].float()
    gv = _rows(g)[b * g_stride + l].float()
    n = xv.norm(dim=-1, keepdim=True)
    y = xv / n
    out = (gv - y * (y * gv).sum(-1, keepdim=True)) / n
    dr = b * L_out + l_off + l
    _rows(dst)[dr] = _rows(dst)[dr] + out if accumulate else out


def batch_sum(x, out, B, L, d, L_out, l_off, accumulate):
    s = _rows(x).view(B, L_out, d)[:, l_off:l_off + L].sum(0)
    out.copy_(out + s if accumulate else s)


def _own_cols(g, b):
    """Columns [c0, c1) of local clip b (ragged layout when g.col_off points at a HOST int32 prefix array)."""
    if g.col_off:
        import ctypes
        arr = (ctypes.c_int32 * (g.b_off + b + 2)).from_address(g.col_off)
        return int(arr[g.b_off + b]), int(arr[g.b_off + b + 1])
    return (g.b_off + b) * g.N, (g.b_off + b + 1) * g.N


def _pos_matrix(g, posbits, Rc, r0):
    """[Rc, C] positives of rows r0.. of one stage (row = b * T + t, local clips)."""
    pos_bnt = unpack_posbits(posbits, g.N)                               # [B, N, T]
    pos = torch.zeros(Rc, g.C)
    for i in range(Rc):
        b, t = divmod(r0 + i, g.T)
        c0, c1 = _own_cols(g, b)
        pos[i, c0:c1] = pos_bnt[b, :c1 - c0, t].float()
    return pos


def _grad_matrix(cos, Rc, r0, g, posbits, col_valid, row_kill, ra, rap, cb, cbp):
    e = torch.exp((cos[:, :g.C] - 1.0) / 0.07) * col_valid.float()[None]
    pos = _pos_matrix(g, posbits, Rc, r0)
    if row_kill is not None:
        own = torch.zeros(Rc, g.C)
        for i in range(Rc):
            b = (r0 + i) // g.T
            c0, c1 = _own_cols(g, b)
            own[i, c0:c1] = 1.0
        e = e * (1.0 - own * row_kill.reshape(-1)[r0:r0 + Rc].float()[:, None])
    rr = slice(r0, r0 + Rc)
    return e * (ra[rr][:, None] + cb[None, :g.C] - pos * (rap[rr][:, None] + cbp[None, :g.C])) / 0.07


def sim_grad_gemm(a, t_pad, r0, g, posbits, col_valid, row_kill, ra, rap, cb, cbp, G, GT=None):
    Rc = a.shape[0]
    Gm = _grad_matrix(a.float() @ t_pad.float().t(), Rc, r0, g, posbits, col_valid, row_kill, ra, rap, cb, cbp)
    G.zero_()
    G[:, :g.C] = Gm.to(BF)
    if GT is not None:
        GT.zero_()
        GT[:g.C, :Rc] = Gm.to(BF).t()


def sim_grad_tiles(z, Rc, r0, g, posbits, col_valid, row_kill, ra, rap, cb, cbp, G, GT):
    Gm = _grad_matrix(z[:Rc].float(), Rc, r0, g, posbits, col_valid, row_kill, ra, rap, cb, cbp).to(BF)
    G[:Rc].zero_()
    G[:Rc, :g.C] = Gm
    GT[:, :pad64(Rc)].zero_()
    GT[:g.C, :Rc] = Gm.t()


def pos_from_time(start, end, valid_u8, B, T, N, out=None):
    return cpu_pos_from_time(start, end, valid_u8, B, T, N)


def sim_workspace_bytes(g):
    return 16


def sim_nce_fwd(vfeat, tfeat, tfeat_stage_stride, g, posbits, col_valid, logits_out, row_sums, col_sums, workspace,
                row_kill=None):
    B, S, T, C, N, d = g.B_loc, g.S, g.T, g.C, g.N, g.d
    v = vfeat.float().reshape(B, S, T, d)
    t = tfeat.float().reshape(-1, C, d)
    t = t.expand(S, C, d) if t.shape[0] == 1 else t
    cos = torch.einsum("bstd,scd->bstc", v, t)
    e = torch.exp((cos - 1.0) / 0.07) * col_valid.float()
    pos = _pos_matrix(g, posbits, B * T, 0).view(B, 1, T, C)
    if row_kill is not None:
        own = torch.zeros(B, 1, 1, C)
        for b in range(B):
            c0, c1 = _own_cols(g, b)
            own[b, 0, 0, c0:c1] = 1.0
        e = e * (1.0 - own * row_kill.view(B, 1, T, 1).float())
    pe = e * pos
    row_sums.copy_(torch.stack((e.sum(-1).reshape(-1), pe.sum(-1).reshape(-1))))
    col_sums.copy_(torch.stack((e.sum(dim=(0, 2)), pe.sum(dim=(0, 2)))))
    if logits_out is not None:
        logits_out.copy_(cos.reshape(logits_out.shape).to(logits_out.dtype))


def nce_reduce(row_sums, col_sums, out4_f64, S, T, C_, row_sel=None, col_sel=None):
    if row_sums is not None:
        B = row_sums.shape[1] // (S * T)
        rs = row_sums.view(2, B, S, T)
        m = rs[1] > 0
        if row_sel is not None:
            m = m & row_sel.view(B, 1, T).bool()
        out4_f64[0] += (rs[0][m].log() - rs[1][m].log()).double().sum()
        out4_f64[1] += m.sum()
    if col_sums is not None:
        cs = col_sums.view(2, -1, C_)
        m = cs[1] > 0
        if col_sel is not None:
            m = m & col_sel.view(1, -1).bool()
        out4_f64[2] += (cs[0][m].log() - cs[1][m].log()).double().sum()
        out4_f64[3] += m.sum()


def own_clip_sim(vfeat, tfeat, shared_text, B, S, T, N, d, s_first=0, s_count=None, out=None):
    s_count = S - s_first if s_count is None else s_count
    v = vfeat.float().reshape(B, S, T, d)[:, s_first:s_first + s_count]
    t = tfeat.float().reshape(-1, B, N, d)
    t = t.expand(S, B, N, d) if t.shape[0] == 1 else t
    return torch.einsum("bstd,sbnd->bstn", v, t[s_first:s_first + s_count]).contiguous()


def agree_scan(own, posbits, vpm_u8, tpm_u8, B, T, N, fill_max):
    """Only what the threshold / head branches read (max_logit without the padding fill); the self-labelling scan
    itself is a kernel-level concern (tests/test_loss_full_gpu.py)."""
    if fill_max:
        raise NotImplementedError("cpu stand-in: agreement self-labelling is tested on the GPU")
    z = own.float() / 0.07                                               # [B, T, N]
    win = torch.zeros(B, N, 2, dtype=torch.int32)
    return win, torch.zeros(B, N), z.max(dim=1).values


def align_stitch(blk_joint, blk_dual, windows_i32, sim_joint, sim_dual, cover, accumulate, finalize):
    """csrc/align.cu: running sums of blk / 0.07 over the windows covering (sentence, frame), divided when `finalize`."""
    if not accumulate:
        for t in (sim_joint, sim_dual, cover):
            t.zero_()
    for w, (t0, t1, n0, n1) in enumerate(windows_i32.tolist()):
        sim_joint[n0:n1, t0:t1] += blk_joint[w, :t1 - t0, :n1 - n0].t() / 0.07
        sim_dual[n0:n1, t0:t1] += blk_dual[w, :t1 - t0, :n1 - n0].t() / 0.07
        cover[n0:n1, t0:t1] += 1
    if finalize:
        den = cover.clamp(min=1e-5)
        sim_joint /= den
        sim_dual /= den


def align_argmax(sim):
    return torch.where(sim != 0, sim, torch.full_like(sim, -6e4)).softmax(-1).argmax(-1)


NAMES = ["align_stitch", "align_argmax", "gemm_tn", "linear_dual", "linear_gelu_bwd", "embed_gather", "own_clip_sim", "agree_scan", "cast_bf16", "linear", "layernorm", "attention", "attention_bwd", "quickgelu_fwd", "quickgelu_bwd",
         "transpose_bf16", "colsum", "layernorm_bwd", "l2norm_bwd", "batch_sum", "sim_grad_gemm", "sim_grad_tiles",
         "pos_from_time", "sim_workspace_bytes", "sim_nce_fwd", "nce_reduce"]


def install(monkeypatch):
    """Swap the stand-ins into temporalalignnet_b200.ops and lift the product's CUDA-only guard (tests only)."""
    import sys

    from temporalalignnet_b200 import ops, tan_model
    me = sys.modules[__name__]
    for n in NAMES:
        monkeypatch.setattr(ops, n, getattr(me, n))
    monkeypatch.setattr(tan_model.TemporalAligner, "_check_device", lambda self, t: None)
    from temporalalignnet_b200 import align
    monkeypatch.setattr(align, "_require_cuda", lambda t, what: None)


def install_plain():
    """The same swap without pytest's monkeypatch (spawned worker processes of the gloo tests)."""
    import sys

    from temporalalignnet_b200 import ops, tan_model
    me = sys.modules[__name__]
    for n in NAMES:
        setattr(ops, n, getattr(me, n))
    tan_model.TemporalAligner._check_device = lambda self, t: None
