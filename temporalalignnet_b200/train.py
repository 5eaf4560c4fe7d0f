"""Training step of the TAN hot path: forward with saved activations ("tape") and the hand-written backward
pass, hooked into torch autograd as ONE node so that `loss.backward()` (train/main.py:112), GradScaler and any
torch optimizer work unchanged.

    model.enable_autograd(True)          # or TAN_AUTOGRAD=1; round-1 status: opt-in, first correct path
    out  = model(video, text, ...)       # same dict as the inference forward; handles carry the tape
    loss = get_loss(..., out, args)['loss']
    loss.backward()                      # -> .grad of every parameter on the path

What autograd does in the reference (over nn.MultiheadAttention / LayerNorm / Linear / einsum / logsumexp,
model/tfm_model.py:17-55, model/tan_model.py:100-234, train/loss.py:231-275) is done here by kernels of
libtan_b200.so: every GEMM-shaped gradient (dgrad, wgrad, the similarity recomputation and its two products) is
a `tan_linear_bf16` call on transposed operands (tcgen05 pair GEMM), the rest are the kernels of backward.cu.
torch only allocates buffers and carries the result into `.grad`.

Not differentiated (raise): `interpolate_from` (evaluation-time option); sine positions are constants.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch

from . import ops
from ._lib import ACT_NONE, TanError
from .tfm_model import StageSink, _f32, _mask_u8

AUTOGRAD_DEFAULT = os.environ.get("TAN_AUTOGRAD", "0") != "0"
# rows of one similarity-gradient chunk (G = rows x C bf16).  Measured on B200 at the bench shape (R = 65 536 rows,
# C = 8192): 29.1 / 20.9 / 18.8 / 16.6 ms of similarity-backward GEMMs per step at 4096 / 8192 / 16384 / 65536 rows --
# a chunk's dA / dB products have only 64 output tiles for 74 CTA pairs, larger chunks amortise that; G is capped
# at SIM_BWD_G_BYTES
SIM_BWD_ROWS = int(os.environ.get("TAN_SIM_BWD_ROWS", "65536"))
SIM_BWD_G_BYTES = 2 << 30
SIM_GRAD_FUSED = os.environ.get("TAN_SIM_GRAD_FUSED", "1") != "0"     # G in the epilogue of the recomputation GEMM
# QuickGELU inside the GEMM epilogues of the training step: c_fc writes pre-activation AND activation
# (tan_linear_dual_bf16), the dgrad through c_proj multiplies with gelu'(u) (tan_linear_gelu_bwd_bf16) -- instead of
# two elementwise passes over the [M, 4d] activations per layer
FUSE_GELU = os.environ.get("TAN_FUSE_GELU", "1") != "0"


# ------------------------------------------------------------------------------------------------------
# tape
# ------------------------------------------------------------------------------------------------------
class LayerTape:
    __slots__ = ("x_in", "xn", "qkv", "att", "lse", "x1", "xn2", "u", "h")


class StackTape:
    """One encoder stack: per-layer activations, the final residual stream and the raw stage features."""

    def __init__(self, enc, post_ln, B: int, L: int, l_split: int, kpm):
        self.enc, self.post_ln, self.B, self.L, self.l_split, self.kpm = enc, post_ln, B, L, l_split, kpm
        self.layers: List[LayerTape] = []
        self.x_out = None
        self.rawA = None          # [S, B, l_split, d] fp32 stage-major raw features of the first part (video)
        self.rawB = None          # [S, B, L - l_split, d] fp32 of the second part (text), or None


class StepTape:
    """Everything one backward pass needs (one forward of one TemporalAligner)."""

    def __init__(self, model):
        self.model = model
        self.video = None         # StackTape
        self.joint = None         # StackTape
        self.consumed = False

    def release(self) -> None:
        """Drop the saved activations (15 GB at the bench shape) as soon as the backward pass has used them: the
        loss tensor -- and with it this tape -- usually stays referenced until the next step's loss replaces it."""
        self.video = self.joint = None
        self.input_leaves = []
        for name in ("vb", "tb", "pre_v", "pre_t", "text_raw"):
            setattr(self, name, None)


def run_encoder_stack_train(enc, x0: torch.Tensor, kpm, B: int, L: int, l_split: int, post_ln,
                            nrm_sink: StageSink) -> StackTape:
    """TemporalEncoder.forward (model/tfm_model.py:48-55) with every activation the backward needs kept:
    unfused kernel sequence (the residual stream is written to a new buffer per half block instead of in place;
    c_fc stores the pre-activation and QuickGELU runs as its own pass).  Stage features: L2-normalised bf16
    into `nrm_sink` (the layout the similarity kernel reads), raw fp32 stage-major into the tape."""
    blocks = list(enc.resblocks)
    S = len(blocks)
    M, d = x0.shape
    dev = x0.device
    cache = enc._cache
    tape = StackTape(enc, post_ln, B, L, l_split, kpm)
    f32 = dict(dtype=torch.float32, device=dev)
    bf = dict(dtype=torch.bfloat16, device=dev)
    nB = L - l_split
    tape.rawA = torch.empty(S, B, l_split, d, **f32)
    tape.rawB = torch.empty(S, B, nB, d, **f32) if nB > 0 else None

    def emit(s: int) -> dict:
        out = dict(nrm_sink.stage_views(s))
        out["rawA"] = tape.rawA[s].view(-1, d)
        if tape.rawB is not None:
            out["rawB"] = tape.rawB[s].view(-1, d)
        return out

    # raw features are stage-major (stride = part length), the normalised ones keep the sink's layout: the
    # LayerNorm kernel takes a second stride pair for the raw rows, so both are emitted by ONE pass over the row
    def ln_emit(x, gamma, beta, out_bf16, s: int) -> None:
        views = emit(s) if s >= 0 else {}
        ops.layernorm(x, M, d, gamma=gamma, beta=beta, L_in=L, out_bf16=out_bf16, l_split=l_split,
                      strideA=nrm_sink.strideA, strideB=nrm_sink.strideB, raw_strideA=l_split, raw_strideB=max(nB, 1),
                      **views)

    x = x0
    for i, blk in enumerate(blocks):
        lt = LayerTape()
        lt.x_in = x
        lt.xn = torch.empty(M, d, **bf)
        ln_emit(x, _f32(blk.ln_1.weight), _f32(blk.ln_1.bias), lt.xn, i - 1)
        lt.qkv = torch.empty(M, 3 * d, **bf)
        ops.linear(lt.xn, cache.get(blk.attn.in_proj_weight), _f32(blk.attn.in_proj_bias), out_bf16=lt.qkv)
        lt.att = torch.empty(M, d, **bf)
        lt.lse = torch.empty(B, blk.n_head, ops.pad64(L), **f32)       # row statistics for the attention backward
        ops.attention(lt.qkv[:, 0:d], lt.qkv[:, d:2 * d], lt.qkv[:, 2 * d:3 * d], kpm, lt.att, B, blk.n_head, L, L,
                      lse=lt.lse)
        lt.x1 = torch.empty(M, d, **f32)
        ops.linear(lt.att, cache.get(blk.attn.out_proj.weight), _f32(blk.attn.out_proj.bias), residual=x, out_f32=lt.x1)
        lt.xn2 = torch.empty(M, d, **bf)
        ops.layernorm(lt.x1, M, d, gamma=_f32(blk.ln_2.weight), beta=_f32(blk.ln_2.bias), L_in=L, out_bf16=lt.xn2)
        lt.u = torch.empty(M, 4 * d, **bf)
        lt.h = torch.empty(M, 4 * d, **bf)
        if FUSE_GELU:
            ops.linear_dual(lt.xn2, cache.get(blk.mlp.c_fc.weight), _f32(blk.mlp.c_fc.bias), out_act=lt.h, out_pre=lt.u)
        else:
            ops.linear(lt.xn2, cache.get(blk.mlp.c_fc.weight), _f32(blk.mlp.c_fc.bias), out_bf16=lt.u, act=ACT_NONE)
            ops.quickgelu_fwd(lt.u, lt.h)
        x_next = torch.empty(M, d, **f32)
        ops.linear(lt.h, cache.get(blk.mlp.c_proj.weight), _f32(blk.mlp.c_proj.bias), residual=lt.x1, out_f32=x_next)
        x = x_next
        tape.layers.append(lt)
    tape.x_out = x
    ln_emit(x, _f32(post_ln.weight), _f32(post_ln.bias), None, S - 1)
    return tape


def forward_train(model, video_embed, lang_embed, video_padding_mask=None, lang_padding_mask=None,
                  interpolate_from=None) -> dict:
    """TemporalAligner.forward (model/tan_model.py:100-149) with a tape.  Same returned dict as the inference
    path (LazyLogits handles; they carry `.tape`)."""
    from .tan_model import LazyLogits
    if interpolate_from:
        raise TanError("interpolate_from is an evaluation-time option; the training forward does not support it")
    if model.use_alignability_head and model.num_decoder_layers < 3:
        raise TanError("the alignability head reads joint stage 2 (train/loss.py:341): num_decoder_layers >= 3")
    B, T, Din = video_embed.shape
    N = lang_embed.shape[1]
    dev = video_embed.device
    d, E, D = model.width, model.num_encoder_layers, model.num_decoder_layers
    L = T + N
    kpm_v = _mask_u8(video_padding_mask, B, T, dev)
    kpm_t = _mask_u8(lang_padding_mask, B, N, dev)
    tape = StepTape(model)
    tape.B, tape.T, tape.N, tape.Din, tape.Dt = B, T, N, Din, lang_embed.shape[2]
    # inputs that take part in autograd (the text backbone's fc1 / fc2 train through `lang_embed` in the reference,
    # train/main.py:58-60): they become inputs of the step's autograd node next to the parameters
    tape.input_leaves = [t_ for t_ in (lang_embed, video_embed) if t_.requires_grad]
    tape.want_text_grad, tape.want_video_grad = bool(lang_embed.requires_grad), bool(video_embed.requires_grad)
    f32 = dict(dtype=torch.float32, device=dev)
    bf = dict(dtype=torch.bfloat16, device=dev)

    # RNG draws in the reference's call order (see TemporalAligner._forward_impl)
    ps_v = model._pos_start(T)
    ps_t = model._pos_start(N) if model.use_text_pos_enc else None
    ps_j = model._pos_start(T)
    tape.ps_v, tape.ps_t, tape.ps_j = ps_v, ps_t, ps_j
    pos_ln_v = model._pos_ln(model.temporal_pos_embed, T, ps_v, None, "v")
    pos_ln_j = pos_ln_v if ps_j == ps_v else model._pos_ln(model.temporal_pos_embed, T, ps_j, None, "j")
    pos_ln_t = model._pos_ln(model.text_temporal_pos_embed, N, ps_t, None, "t") if ps_t is not None else None

    # pre-projections (own buffers: the tape outlives the model's scratch)
    v = video_embed.detach()
    tape.vb = v.contiguous().view(B * T, Din) if v.dtype == torch.bfloat16 else \
        ops.cast_bf16(v.float().contiguous().view(B * T, Din))
    t = lang_embed.detach()
    tape.tb = t.contiguous().view(B * N, tape.Dt) if t.dtype == torch.bfloat16 else \
        ops.cast_bf16(t.float().contiguous().view(B * N, tape.Dt))
    tape.pre_v = torch.empty(B * T, d, **f32)
    tape.pre_t = torch.empty(B * N, d, **f32)
    ops.linear(tape.vb, model._cache.get(model.video_pre_proj.weight), out_f32=tape.pre_v)
    ops.linear(tape.tb, model._cache.get(model.text_pre_proj.weight), out_f32=tape.pre_t)

    # ---- video (dual) stack -------------------------------------------------------------------------
    x0v = torch.empty(B * T, d, **f32)
    ops.layernorm(tape.pre_v, B * T, d, gamma=_f32(model.ln_video_init.weight), beta=_f32(model.ln_video_init.bias),
                  add=pos_ln_v, add_rows=T, L_in=T, out_f32=x0v)
    vfeat_dual = torch.empty(B, E, T, d, **bf)
    sink_v = StageSink(E, l_split=T, strideA=E * T, nrmA_bf16=vfeat_dual, offA=T)
    tape.video = run_encoder_stack_train(model.video_temporal_encoder, x0v, kpm_v, B, T, T, model.ln_video_post_enc,
                                         sink_v)
    # ---- joint stack --------------------------------------------------------------------------------
    x0j = torch.empty(B * L, d, **f32)
    ops.layernorm(tape.pre_v, B * T, d, gamma=_f32(model.ln_video_init.weight), beta=_f32(model.ln_video_init.bias),
                  add=pos_ln_j, add_rows=T, L_in=T, L_out=L, l_off=0, out_f32=x0j)
    ops.layernorm(tape.pre_t, B * N, d, gamma=_f32(model.ln_text_init.weight), beta=_f32(model.ln_text_init.bias),
                  add=pos_ln_t, add_rows=N, L_in=N, L_out=L, l_off=T, out_f32=x0j)
    if kpm_v is None and kpm_t is None:
        kpm = None
    else:
        kv = kpm_v if kpm_v is not None else torch.zeros(B, T, dtype=torch.uint8, device=dev)
        kt = kpm_t if kpm_t is not None else torch.zeros(B, N, dtype=torch.uint8, device=dev)
        kpm = torch.cat((kv, kt), dim=1).contiguous()
    vfeat_joint = torch.empty(B, D, T, d, **bf)
    tfeat_joint = torch.empty(D, B * N, d, **bf)
    sink_j = StageSink(D, l_split=T, strideA=D * T, strideB=N, nrmA_bf16=vfeat_joint, nrmB_bf16=tfeat_joint, offA=T,
                       offB=B * N)
    tape.joint = run_encoder_stack_train(model.joint_temporal_encoder, x0j, kpm, B, L, T, model.ln_joint_post_enc,
                                         sink_j)
    # ---- dual text features -------------------------------------------------------------------------
    tape.text_raw = torch.empty(B, N, d, **f32)
    tfeat_dual = torch.empty(B * N, d, **bf)
    tfeat_dual_f32 = torch.empty(B, N, d, **f32) if model.return_dual_feature else None
    ops.layernorm(tape.pre_t, B * N, d, gamma=_f32(model.ln_text_init.weight), beta=_f32(model.ln_text_init.bias),
                  L_in=N, l_split=N, strideA=N, rawA=tape.text_raw, nrmA_bf16=tfeat_dual, nrmA_f32=tfeat_dual_f32)

    logits_dual = LazyLogits(vfeat_dual, tfeat_dual, shared_text=True, N=N)
    logits_joint = LazyLogits(vfeat_joint, tfeat_joint, shared_text=False, N=N)
    logits_dual.tape = tape
    logits_joint.tape = tape
    out = {'logits_dual': logits_dual, 'logits_joint': logits_joint}
    if model.return_dual_feature:
        out['dual_feature_video'] = vfeat_dual
        out['dual_feature_text'] = tfeat_dual_f32
    if model.use_alignability_head:      # Linear(d, 1) on the raw text features (model/tan_model.py:146-148): tiny glue
        out['dual_logits_alignability'] = model._binary_head(tape.text_raw)
        out['joint_logits_alignability'] = model._binary_head(tape.joint.rawB.permute(1, 0, 2, 3))    # [B,D,N,1]
    return out


# ------------------------------------------------------------------------------------------------------
# backward
# ------------------------------------------------------------------------------------------------------
class _Grads:
    """fp32 gradient buffers per parameter (zero-initialised; kernels accumulate into them).  The parameters of one
    transformer block live in ONE flat buffer (`bucket`), so that with several GPUs a finished block's gradients are
    all-reduced by a single asynchronous collective that overlaps the backward pass of the blocks below it
    ("bucketing per layer")."""

    def __init__(self, dist=None):
        self.g = {}
        self.dist = dist
        self.buckets = {}          # id(block) -> flat tensor
        self.pending = []          # async all-reduce handles
        self.reduced = set()       # ids of parameters whose gradient is already summed over the ranks

    def of(self, p: torch.Tensor) -> torch.Tensor:
        t = self.g.get(id(p))
        if t is None:
            t = torch.zeros(p.shape, dtype=torch.float32, device=p.device)
            self.g[id(p)] = t
        return t

    def get(self, p):
        return self.g.get(id(p))

    def bucket(self, block) -> None:
        """Place the gradient buffers of every trainable parameter of `block` in one flat tensor (before first use)."""
        if id(block) in self.buckets:
            return
        ps = [p for p in block.parameters() if p.requires_grad and id(p) not in self.g]
        if not ps:
            return
        flat = torch.zeros(sum(p.numel() for p in ps), dtype=torch.float32, device=ps[0].device)
        off = 0
        for p in ps:
            self.g[id(p)] = flat[off:off + p.numel()].view(p.shape)
            off += p.numel()
        self.buckets[id(block)] = (flat, ps)

    def reduce_bucket(self, block) -> None:
        """All ranks have finished this block's gradients: sum them over the ranks, asynchronously."""
        ent = self.buckets.get(id(block))
        if self.dist is None or ent is None:
            return
        flat, ps = ent
        self.pending.append(self.dist.all_reduce(flat, async_op=True))
        self.reduced.update(id(p) for p in ps)

    def wait(self) -> None:
        for w in self.pending:
            w.wait()
        self.pending = []


def _wgrad(dy_bf16: torch.Tensor, x_bf16: torch.Tensor, gw: torch.Tensor, gb: Optional[torch.Tensor] = None) -> None:
    """gw [N, K] += dy^T @ x   (dy [M, N], x [M, K] bf16): tan_gemm_tn_bf16 consumes both operands as they lie in HBM
    (MN-major UMMA operands) and splits the long contraction over the CTA pairs inside the one launch (fixed-order
    partial sums: deterministic).  gb [N]: the bias gradient (column sums of dy)."""
    if gb is not None:
        ops.colsum(dy_bf16, gb)
    ops.gemm_tn(dy_bf16, x_bf16, gw, accumulate=True)


def _dgrad(dy_bf16: torch.Tensor, wT: torch.Tensor, out_f32=None, out_bf16=None, residual=None) -> None:
    """dx = dy @ W  (dy [M, N], wT = W^T [K, N] bf16)."""
    ops.linear(dy_bf16, wT, residual=residual, out_f32=out_f32, out_bf16=out_bf16, tag="dgrad")


def stack_backward(tape: StackTape, stage_grads: List[Optional[torch.Tensor]], grads: _Grads) -> torch.Tensor:
    """Backward of run_encoder_stack_train.  stage_grads[s]: fp32 [B*L, d] token-major gradient with respect to the
    RAW stage-s features (ln_1 of block s+1 / post_ln output).  Returns the gradient of the stack input [B*L, d]."""
    blocks = list(tape.enc.resblocks)
    cache = tape.enc._cache
    S = len(blocks)
    B, L = tape.B, tape.L
    M, d = tape.x_out.shape
    dev = tape.x_out.device
    f32 = dict(dtype=torch.float32, device=dev)
    bf = dict(dtype=torch.bfloat16, device=dev)
    dx = torch.empty(M, d, **f32)
    dxb = torch.empty(M, d, **bf)
    post = tape.post_ln
    for blk in blocks:
        grads.bucket(blk)
    # Every LayerNorm backward also emits the bf16 copy of the residual-stream gradient it has just updated and that
    # gradient's column sums: the operand and the bias gradient of the linear layer whose backward comes next (c_proj
    # of the block above / out_proj of this block) -- instead of a cast pass and a column-sum pass each.
    ops.layernorm_bwd(stage_grads[S - 1], tape.x_out, _f32(post.weight), dx, False, M, d, grads.of(post.weight),
                      grads.of(post.bias), dx_bf16=dxb, dx_colsum=grads.of(blocks[S - 1].mlp.c_proj.bias))
    dh = torch.empty(M, 4 * d, **bf)
    dy32 = torch.empty(M, d, **f32)
    datt = torch.empty(M, d, **bf)
    dqkv = torch.empty(M, 3 * d, **bf)
    delta = torch.empty(B * blocks[0].n_head * ops.pad64(L), **f32)
    for i in range(S - 1, -1, -1):
        blk, lt = blocks[i], tape.layers[i]
        H = blk.n_head
        # ---- MLP: x_out = x1 + c_proj(gelu(c_fc(ln_2(x1)))) ------------------------------------------
        if FUSE_GELU:        # du = (dx @ W_proj) o gelu'(u) in one GEMM
            ops.linear_gelu_bwd(dxb, ops.transpose_bf16(cache.get(blk.mlp.c_proj.weight)), lt.u, dh)
            _wgrad(dxb, lt.h, grads.of(blk.mlp.c_proj.weight))
        else:
            _dgrad(dxb, ops.transpose_bf16(cache.get(blk.mlp.c_proj.weight)), out_bf16=dh)
            _wgrad(dxb, lt.h, grads.of(blk.mlp.c_proj.weight))
            ops.quickgelu_bwd(dh, lt.u, dh)                                    # du, in place
        _dgrad(dh, ops.transpose_bf16(cache.get(blk.mlp.c_fc.weight)), out_f32=dy32)
        _wgrad(dh, lt.xn2, grads.of(blk.mlp.c_fc.weight), grads.of(blk.mlp.c_fc.bias))
        ops.layernorm_bwd(dy32, lt.x1, _f32(blk.ln_2.weight), dx, True, M, d, grads.of(blk.ln_2.weight),
                          grads.of(blk.ln_2.bias), dx_bf16=dxb, dx_colsum=grads.of(blk.attn.out_proj.bias))
        # ---- attention: x1 = x + out_proj(attn(in_proj(ln_1(x)))) --------------------------------------
        _dgrad(dxb, ops.transpose_bf16(cache.get(blk.attn.out_proj.weight)), out_bf16=datt)
        _wgrad(dxb, lt.att, grads.of(blk.attn.out_proj.weight))
        q, k, v = lt.qkv[:, 0:d], lt.qkv[:, d:2 * d], lt.qkv[:, 2 * d:3 * d]
        ops.attention_bwd(q, k, v, lt.att, datt, tape.kpm, dqkv[:, 0:d], dqkv[:, d:2 * d], dqkv[:, 2 * d:3 * d], lt.lse,
                          delta, B, H, L, L)
        sg = stage_grads[i - 1] if i >= 1 else None      # ln_1 of block i IS stage i-1 (model/tfm_model.py:50-53)
        _dgrad(dqkv, ops.transpose_bf16(cache.get(blk.attn.in_proj_weight)), out_f32=dy32, residual=sg)
        _wgrad(dqkv, lt.xn, grads.of(blk.attn.in_proj_weight), grads.of(blk.attn.in_proj_bias))
        below = blocks[i - 1].mlp.c_proj.bias if i >= 1 else None      # the next consumer of dx
        ops.layernorm_bwd(dy32, lt.x_in, _f32(blk.ln_1.weight), dx, True, M, d, grads.of(blk.ln_1.weight),
                          grads.of(blk.ln_1.bias), dx_bf16=dxb if i >= 1 else None,
                          dx_colsum=grads.of(below) if below is not None else None)
        grads.reduce_bucket(blk)       # several GPUs: this block's gradients travel while the next block computes
    return dx


class SimCtx:
    """What get_loss hands to the backward for ONE model: the exp-sums of the forward (columns already summed over
    all ranks), the targets and the selections."""

    def __init__(self, logits, row_sums, col_sums, nce, row_sel=None, col_sel=None):
        self.logits, self.row_sums, self.col_sums, self.nce = logits, row_sums, col_sums, nce
        self.row_sel, self.col_sel = row_sel, col_sel


def sim_coefficients(ctx: SimCtx, scale: torch.Tensor, dist):
    """Row / column coefficient vectors of dL/dz (train/loss.py:248-256 differentiated): with A = sum_all e,
    P = sum_pos e of a counted row, v = log A - log P and dL/dv = scale * 0.5 / n_rows:
    ra = w / A, rap = w / P (0 for rows that do not count); the same for columns.  [R]- and [S*C]-sized vectors."""
    lg, nce = ctx.logits, ctx.nce
    B, S, T, _ = lg.vfeat.shape
    rs = ctx.row_sums.view(2, B, S, T)
    cs = ctx.col_sums                                              # [2, S, C]
    rsel = rs[1] > 0
    if ctx.row_sel is not None:
        rsel = rsel & ctx.row_sel.view(B, 1, T).bool()
    csel = cs[1] > 0
    if ctx.col_sel is not None:
        csel = csel & ctx.col_sel.view(1, -1).bool()
    n_r = rsel.sum().to(torch.float32)
    if dist is not None:
        dist.all_reduce(n_r)
    n_c = csel.sum().to(torch.float32)
    wr = scale * 0.5 / n_r
    wc = scale * 0.5 / n_c
    zero_r = torch.zeros_like(rs[0])
    ra = torch.where(rsel, wr / rs[0].clamp_min(1e-37), zero_r)
    rap = torch.where(rsel, wr / rs[1].clamp_min(1e-37), zero_r)
    zero_c = torch.zeros_like(cs[0])
    cb = torch.where(csel, wc / cs[0].clamp_min(1e-37), zero_c)
    cbp = torch.where(csel, wc / cs[1].clamp_min(1e-37), zero_c)
    # stage-major rows: [S, B*T]
    ra = ra.permute(1, 0, 2).reshape(S, B * T).contiguous()
    rap = rap.permute(1, 0, 2).reshape(S, B * T).contiguous()
    return ra, rap, cb.contiguous(), cbp.contiguous()


def sim_backward(ctx: SimCtx, scale: torch.Tensor, dist):
    """Gradients of `scale * loss_x` with respect to the L2-normalised features of one model:
    (d_v [S, B*T, d] fp32 stage-major, d_t [S_t, C_pad, d] fp32 over the GLOBAL columns, partial over local rows)."""
    from .loss import gather_text_features
    lg, nce = ctx.logits, ctx.nce
    vfeat, tfeat = lg.vfeat, lg.tfeat
    B, S, T, d = vfeat.shape
    dev = vfeat.device
    if dist is not None:
        tfeat = gather_text_features(tfeat, lg.shared_text, dist)
    tfeat = nce.compact_features(tfeat)            # ragged columns: the padded sentences are not computed at all
    S_t = 1 if lg.shared_text else S
    tfeat = tfeat.view(S_t, -1, d)
    C = tfeat.shape[1]
    Cp = (C + 127) // 128 * 128
    if Cp != C:
        tpad = torch.zeros(S_t, Cp, d, dtype=torch.bfloat16, device=dev)
        tpad[:, :C] = tfeat
    else:
        tpad = tfeat.contiguous()
    ra, rap, cb, cbp = sim_coefficients(ctx, scale, dist)
    vsm = vfeat.permute(1, 0, 2, 3).reshape(S, B * T, d).contiguous()         # stage-major rows
    R = B * T
    Rc = min(R, SIM_BWD_ROWS, max(256, (SIM_BWD_G_BYTES // (2 * Cp)) // 256 * 256))
    g = nce.geom(B, 1, T, d)
    fused = SIM_GRAD_FUSED and nce.N <= 64 and d % 64 == 0
    z = None if fused else torch.empty(Rc, Cp, dtype=torch.float32, device=dev)
    G = torch.empty(Rc, Cp, dtype=torch.bfloat16, device=dev)
    GT = None if fused else torch.empty(C, ops.pad64(Rc), dtype=torch.bfloat16, device=dev)
    d_v = torch.empty(S, R, d, dtype=torch.float32, device=dev)
    d_t = torch.zeros(S_t, Cp, d, dtype=torch.float32, device=dev)
    for s in range(S):
        si = 0 if lg.shared_text else s
        ts = tpad[si]
        tT = ops.transpose_bf16(ts)                                           # [d, Cp] (small: the text side)
        for r0 in range(0, R, Rc):
            rc = min(Rc, R - r0)
            a = vsm[s, r0:r0 + rc]
            if fused:
                # cosines recomputed and turned into G inside ONE GEMM (epilogue); the fp32 cosines never reach HBM
                ops.sim_grad_gemm(a, ts, r0, g, nce.posbits, nce.col_valid, nce.row_kill, ra[s], rap[s], cb[s], cbp[s],
                                  G[:rc])
            else:
                ops.linear(a, ts, out_f32=z[:rc], tag="sim_bwd")                             # cosines of the chunk
                ops.sim_grad_tiles(z, rc, r0, g, nce.posbits, nce.col_valid, nce.row_kill, ra[s], rap[s], cb[s],
                                   cbp[s], G, GT)
            ops.linear(G[:rc], tT, out_f32=d_v[s, r0:r0 + rc], tag="sim_bwd")                # dA = G @ text
            # dB += G^T @ video straight from G and the video rows as they lie (MN-major operands): no transposes
            ops.gemm_tn(G[:rc], a, d_t[si], accumulate=True, tag="sim_bwd")
    return d_v, nce.scatter_columns(d_t)       # text gradients back in the padded [B_glob * N] column layout


def step_backward(tape: StepTape, ctx_dual: SimCtx, ctx_joint: SimCtx, grad_out: torch.Tensor, nce_weight: float,
                  dist, params: List[torch.Tensor], bce_dx: Optional[torch.Tensor] = None) -> List[Optional[torch.Tensor]]:
    """d (grad_out * (nce_weight * (loss_dual + loss_joint) / 2 + loss_bce)) / d params, in the order of `params`.
    bce_dx [B_loc, N]: d loss_bce / d joint_logits_alignability[:, 2, :, 0] of the local clips (train/loss.py:341-351)."""
    model = tape.model
    B, T, N = tape.B, tape.T, tape.N
    d, E, D = model.width, model.num_encoder_layers, model.num_decoder_layers
    L = T + N
    dev = tape.pre_v.device
    f32 = dict(dtype=torch.float32, device=dev)
    grads = _Grads(dist)
    scale = (grad_out.detach().to(torch.float32) * (0.5 * nce_weight)).reshape(())
    b_off = ctx_dual.nce.b_off

    # ---- similarity + NCE -----------------------------------------------------------------------------
    dv_dual, dt_dual = sim_backward(ctx_dual, scale, dist)        # [E, B*T, d], [1, Cp, d]
    dv_joint, dt_joint = sim_backward(ctx_joint, scale, dist)     # [D, B*T, d], [D, Cp, d]
    if dist is not None:                                          # text-feature gradients: sum over the ranks' rows
        dist.all_reduce(dt_dual)
        dist.all_reduce(dt_joint)
    own = slice(b_off * N, (b_off + B) * N)

    # ---- L2 normalisation -> gradients of the raw stage features, token-major per stage -----------------
    vt, jt = tape.video, tape.joint
    sg_v = []
    for s in range(E):
        buf = torch.empty(B * T, d, **f32)
        ops.l2norm_bwd(vt.rawA[s].view(-1, d), dv_dual[s], buf, False, B * T, d, T, T, T, 0)
        sg_v.append(buf)
    sg_j = []
    for s in range(D):
        buf = torch.empty(B * L, d, **f32)
        ops.l2norm_bwd(jt.rawA[s].view(-1, d), dv_joint[s], buf, False, B * T, d, T, T, L, 0)
        ops.l2norm_bwd(jt.rawB[s].view(-1, d), dt_joint[s, own], buf, False, B * N, d, N, N, L, T)
        sg_j.append(buf)
    sg_text = torch.empty(B * N, d, **f32)
    ops.l2norm_bwd(tape.text_raw.view(-1, d), dt_dual[0, own], sg_text, False, B * N, d, N, N, N, 0)

    # ---- alignability head: Linear(d, 1) on the raw joint text features of stage 2 ([B*N]-sized glue) -------
    if bce_dx is not None:
        head = model.binary_head
        gx = (bce_dx.to(torch.float32) * grad_out.detach().to(torch.float32)).reshape(B * N, 1)
        feat = tape.joint.rawB[2].reshape(B * N, d)
        grads.of(head.weight).add_((gx * feat).sum(0, keepdim=True))
        grads.of(head.bias).add_(gx.sum().reshape(1))
        sg_j[2].view(B, L, d)[:, T:, :].add_((gx * head.weight.detach().float().view(1, d)).view(B, N, d))

    # ---- encoder stacks -------------------------------------------------------------------------------
    dx0v = stack_backward(vt, sg_v, grads)                        # [B*T, d]
    dx0j = stack_backward(jt, sg_j, grads)                        # [B*L, d]

    # ---- input LayerNorms, positional tables, pre-projections -------------------------------------------
    lv, lt_, lp = model.ln_video_init, model.ln_text_init, model.ln_position_init
    dpre_v = torch.empty(B * T, d, **f32)
    dpre_t = torch.empty(B * N, d, **f32)
    ops.layernorm_bwd(dx0v, tape.pre_v, _f32(lv.weight), dpre_v, False, B * T, d, grads.of(lv.weight), grads.of(lv.bias))
    ops.layernorm_bwd(dx0j, tape.pre_v, _f32(lv.weight), dpre_v, True, B * T, d, grads.of(lv.weight), grads.of(lv.bias),
                      L_in=T, L_out=L, l_off=0)
    ops.layernorm_bwd(dx0j, tape.pre_t, _f32(lt_.weight), dpre_t, False, B * N, d, grads.of(lt_.weight),
                      grads.of(lt_.bias), L_in=N, L_out=L, l_off=T)
    ops.layernorm_bwd(sg_text, tape.pre_t, _f32(lt_.weight), dpre_t, True, B * N, d, grads.of(lt_.weight),
                      grads.of(lt_.bias))

    def pos_backward(table, start, dsum, n):
        """ln_position_init(table[start:start+n]) was added to every clip: LayerNorm backward of the batch sum."""
        if not (isinstance(table, torch.nn.Parameter) and table.requires_grad):
            # sine table (buffer): only the LayerNorm's parameters receive gradient
            scratch = torch.empty(n, d, **f32)
            ops.layernorm_bwd(dsum, _f32(table)[start:start + n], _f32(lp.weight), scratch, False, n, d,
                              grads.of(lp.weight), grads.of(lp.bias))
            return
        gt = grads.of(table)
        ops.layernorm_bwd(dsum, _f32(table)[start:start + n], _f32(lp.weight), gt[start:start + n], True, n, d,
                          grads.of(lp.weight), grads.of(lp.bias))

    dpos = torch.empty(T, d, **f32)
    ops.batch_sum(dx0v, dpos, B, T, d, T, 0, False)
    if tape.ps_j == tape.ps_v:
        ops.batch_sum(dx0j, dpos, B, T, d, L, 0, True)
        pos_backward(model.temporal_pos_embed, tape.ps_v, dpos, T)
    else:
        pos_backward(model.temporal_pos_embed, tape.ps_v, dpos, T)
        dpos_j = torch.empty(T, d, **f32)
        ops.batch_sum(dx0j, dpos_j, B, T, d, L, 0, False)
        pos_backward(model.temporal_pos_embed, tape.ps_j, dpos_j, T)
    if tape.ps_t is not None:
        dpos_t = torch.empty(N, d, **f32)
        ops.batch_sum(dx0j, dpos_t, B, N, d, L, T, False)
        pos_backward(model.text_temporal_pos_embed, tape.ps_t, dpos_t, N)

    dpre_v_bf, dpre_t_bf = ops.cast_bf16(dpre_v), ops.cast_bf16(dpre_t)
    _wgrad(dpre_v_bf, tape.vb, grads.of(model.video_pre_proj.weight))
    _wgrad(dpre_t_bf, tape.tb, grads.of(model.text_pre_proj.weight))
    # gradients of the inputs that asked for one (dgrad of the bias-free pre-projections)
    tape.input_grads = []
    if tape.want_text_grad:
        d_text = torch.empty(B * N, tape.Dt, **f32)
        _dgrad(dpre_t_bf, ops.transpose_bf16(model._cache.get(model.text_pre_proj.weight)), out_f32=d_text)
        tape.input_grads.append(d_text.view(B, N, tape.Dt))
    if tape.want_video_grad:
        d_video = torch.empty(B * T, tape.Din, **f32)
        _dgrad(dpre_v_bf, ops.transpose_bf16(model._cache.get(model.video_pre_proj.weight)), out_f32=d_video)
        tape.input_grads.append(d_video.view(B, T, tape.Din))

    out = [grads.get(p) for p in params]
    if dist is not None:       # weights are replicated: sum the ranks' gradients.  The transformer blocks went out per
        # block during the backward pass (_Grads.reduce_bucket); what is left (pre-projections, input LayerNorms,
        # positional tables, head) goes in one flat all-reduce
        live = [g_ for p, g_ in zip(params, out) if g_ is not None and id(p) not in grads.reduced]
        if live:
            flat = torch.cat([g_.reshape(-1) for g_ in live])
            dist.all_reduce(flat)
            off = 0
            for g_ in live:
                g_.copy_(flat[off:off + g_.numel()].view_as(g_))
                off += g_.numel()
        grads.wait()
    return out


class _TanLossFn(torch.autograd.Function):
    """One autograd node for the whole step: forward returns the loss value the kernels already computed,
    backward runs `step_backward` and hands every parameter its gradient."""

    @staticmethod
    def forward(ctx, loss_value, holder, *inputs):          # inputs = parameters + input tensors that require grad
        ctx.holder = holder
        ctx.n = len(inputs)
        return loss_value.detach().clone()

    @staticmethod
    def backward(ctx, grad_out):
        h = ctx.holder
        if h["tape"].consumed:
            raise TanError("this forward's tape was already consumed by a backward pass (retain_graph is not supported)")
        h["tape"].consumed = True
        with torch.no_grad():
            gs = step_backward(h["tape"], h["dual"], h["joint"], grad_out, h["nce_weight"], h["dist"], h["params"],
                               h["bce_dx"])
        in_grads = [g.to(t_.dtype) for g, t_ in zip(h["tape"].input_grads, h["tape"].input_leaves)]
        h["tape"].release()
        h["dual"] = h["joint"] = None                 # the features / exp-sums of the step
        gs = [None if g is None else g.to(p.dtype) for g, p in zip(gs, h["params"])]
        return (None, None, *gs, *in_grads)


def attach_autograd(loss_value: torch.Tensor, tape: StepTape, ctx_dual: SimCtx, ctx_joint: SimCtx, nce_weight: float,
                    dist, bce_dx: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Wrap the computed loss value into a tensor whose `.backward()` runs the hand-written backward pass."""
    params = [p for p in tape.model.parameters() if p.requires_grad]
    holder = dict(tape=tape, dual=ctx_dual, joint=ctx_joint, nce_weight=float(nce_weight), dist=dist, params=params,
                  bce_dx=bce_dx)
    return _TanLossFn.apply(loss_value, holder, *params, *tape.input_leaves)


@torch.no_grad()
def clip_gradients(model, clip_grad=3, return_norms: bool = True):
    """utils/train_utils.py:3-13 (per-parameter L2 clipping, called at train/main.py:115-116) without its one
    `.item()` host synchronisation PER PARAMETER (~160 per step): multi-tensor norms, one clamped coefficient
    vector, one multi-tensor scale.  Same arithmetic as the reference (coefficients >= 1 leave the gradient
    untouched bit for bit).  Returns the list of norms like the reference (one synchronisation), or the device
    vector with return_norms=False (none)."""
    grads = [p.grad for _, p in model.named_parameters() if p.grad is not None]
    if not grads:
        return []
    norms = torch.stack(torch._foreach_norm(grads, 2))
    coef = (clip_grad / (norms + 1e-6)).clamp(max=1.0)
    torch._foreach_mul_(grads, list(coef.unbind()))
    return norms.tolist() if return_norms else norms
