"""Transformer blocks of TAN with the reference's class names, constructor arguments, parameter
names and return contracts (model/tfm_model.py), executed by the sm_100a kernels of
libtan_b200.so.

The nn.Module tree (nn.MultiheadAttention / nn.LayerNorm / nn.Linear) is used purely as a parameter
container so that `state_dict()` keys and shapes equal the reference's
(`resblocks.{i}.attn.in_proj_weight`, `...mlp.c_fc.weight`, ...); their torch `forward`s are never
called.  The standalone TemporalEncoder / TemporalDecoder modules are inference-only (outputs carry no autograd
graph); inside TemporalAligner the stacks are trained through train.py (taped forward + hand-written backward).
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict
from typing import List, Optional

import torch
from torch import nn
from torch.nn import LayerNorm

from . import ops
from ._lib import ACT_NONE, ACT_QUICKGELU, TanError


class QuickGELU(nn.Module):
    """model/tfm_model.py:11-13.  In the CUDA path the activation is fused into the c_fc GEMM
    epilogue (TAN_ACT_QUICKGELU); this module only exists for the `mlp.gelu` slot."""

    def forward(self, x: torch.Tensor):
        raise TanError("QuickGELU is fused into tan_linear_bf16; it is not callable on its own")


class _Bf16Cache:
    """bf16 shadow copies of fp32 parameters, refreshed when the parameter is modified in place."""

    def __init__(self):
        self._c = {}

    def get(self, p: torch.Tensor) -> torch.Tensor:
        key = id(p)
        ent = self._c.get(key)
        if ent is None or ent[0] != p._version or ent[1] != p.data_ptr():
            w = p.detach()
            if w.dtype != torch.float32 or not w.is_contiguous():
                w = w.float().contiguous()
            # refresh IN PLACE when possible: captured CUDA graphs keep pointing at the same shadow
            keep = ent[2] if (ent is not None and ent[2].shape == w.shape and ent[2].device == w.device) else None
            ent = (p._version, p.data_ptr(), ops.cast_bf16(w, keep), p)
            self._c[key] = ent
        return ent[2]

    def refresh(self) -> None:
        """Re-cast every shadow whose parameter changed since it was made (called before a graph replay)."""
        for ent in list(self._c.values()):
            p = ent[3]
            if ent[0] != p._version or ent[1] != p.data_ptr():
                self.get(p)


# Fusion switches (A/B aids; defaults follow scripts/ab_fuse.py on B200):
#   out-projection + residual + ln_2 (K = 512): 136 -> 111 us at 65 536 tokens, 33 -> 23 us at 16 384, neutral below
#   c_proj + residual + next LayerNorm + stage features (K = 2048): 172 -> 193 us -- the wide tile gives up the
#   overlap of MMAs and epilogue, which a K = 2048 GEMM needs -- so it stays off unless asked for.
FUSE_OUTPROJ_LN = os.environ.get("TAN_FUSE_OUTPROJ_LN", "1") != "0"
FUSE_OUTPROJ_LN_MIN_TOKENS = 12288
FUSE_CPROJ_LN = os.environ.get("TAN_FUSE_CPROJ_LN", "0") != "0"


def _f32(p: torch.Tensor) -> torch.Tensor:
    t = p.detach()
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.float().contiguous()
    return t


class ResidualAttentionBlock_Step(nn.Module):
    """model/tfm_model.py:17-38 (parameter container + kernel sequence)."""

    def __init__(self, d_model: int, n_head: int):
        super().__init__()
        if d_model != n_head * 64:
            raise TanError(f"head_dim must be 64 (d_model={d_model}, n_head={n_head})")
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([
            ("c_fc", nn.Linear(d_model, d_model * 4)),
            ("gelu", QuickGELU()),
            ("c_proj", nn.Linear(d_model * 4, d_model)),
        ]))
        self.ln_2 = LayerNorm(d_model)
        self.d_model = d_model
        self.n_head = n_head


class _StackBuffers:
    """Scratch activations of one encoder stack for M = B*L tokens (allocated once per shape)."""

    def __init__(self, M: int, d: int, device):
        bf = dict(dtype=torch.bfloat16, device=device)
        self.xn = torch.empty(M, d, **bf)
        self.qkv = torch.empty(M, 3 * d, **bf)
        self.att = torch.empty(M, d, **bf)
        self.h = torch.empty(M, 4 * d, **bf)


class StageSink:
    """Where the per-stage features of a stack go (see tan_layernorm's stage emission).
    Tensors are the FULL buffers; `stage_views(s)` returns the per-stage base views."""

    def __init__(self, S: int, l_split: int, strideA: int = 0, strideB: int = 0,
                 rawA=None, rawB=None, nrmA_bf16=None, nrmB_bf16=None, nrmA_f32=None, nrmB_f32=None,
                 offA: int = 0, offB: int = 0):
        self.S, self.l_split, self.strideA, self.strideB = S, l_split, strideA, strideB
        self.t = dict(rawA=rawA, rawB=rawB, nrmA_bf16=nrmA_bf16, nrmB_bf16=nrmB_bf16, nrmA_f32=nrmA_f32,
                      nrmB_f32=nrmB_f32)
        self.offA, self.offB = offA, offB       # rows to advance per stage for the A / B parts

    def stage_views(self, s: int) -> dict:
        out = {}
        for k, t in self.t.items():
            if t is None:
                continue
            off = self.offA if k in ("rawA", "nrmA_bf16", "nrmA_f32") else self.offB
            out[k] = t.view(-1, t.shape[-1])[s * off:]
        return out


def run_encoder_stack(blocks, x: torch.Tensor, kpm_u8: Optional[torch.Tensor], B: int, L: int,
                      cache: _Bf16Cache, buf: _StackBuffers, sink: StageSink,
                      post_ln: Optional[LayerNorm] = None, emit_final: bool = True) -> None:
    """TemporalEncoder.forward (model/tfm_model.py:48-55) on the residual stream x [B*L, d] fp32,
    updated IN PLACE (x ends as x_S).  Stage s < S-1 is emitted by layer s+1's ln_1 (the reference
    returns that x_norm, :50-53).  Stage S-1: with emit_final, `post_ln(x_S)` (model/tan_model.py:174,
    :206) -- or x_S itself when post_ln is None -- goes to the sink; without, the caller reads x."""
    S = len(blocks)
    M, d = x.shape
    sk = dict(l_split=sink.l_split, strideA=sink.strideA, strideB=sink.strideB)
    l_split = min(sink.l_split, L) if sink.l_split > 0 else L

    def fusable(emit: dict) -> bool:
        """c_proj + residual + the following LayerNorm (+ bf16 stage features) in one kernel (tan_linear_res_ln_stage):
        width 512, 32-token groups inside one clip and one part, and only the bf16 normalised features wanted."""
        if not (FUSE_CPROJ_LN and d == 512 and L % 32 == 0 and l_split % 32 == 0):
            return False
        if not set(emit) <= {"nrmA_bf16", "nrmB_bf16"}:
            return False
        return (not emit) or ("nrmA_bf16" in emit and (l_split == L or "nrmB_bf16" in emit))

    xn_ready = False          # buf.xn already holds ln_1 of the current block (written by the previous block's c_proj)
    final_done = False
    for i, blk in enumerate(blocks):
        if not xn_ready:
            emit = sink.stage_views(i - 1) if i >= 1 else {}
            ops.layernorm(x, M, d, gamma=_f32(blk.ln_1.weight), beta=_f32(blk.ln_1.bias), L_in=L, out_bf16=buf.xn,
                          **sk, **emit)
        ops.linear(buf.xn, cache.get(blk.attn.in_proj_weight), _f32(blk.attn.in_proj_bias), out_bf16=buf.qkv)
        ops.attention(buf.qkv[:, 0:d], buf.qkv[:, d:2 * d], buf.qkv[:, 2 * d:3 * d], kpm_u8, buf.att, B, blk.n_head,
                      L, L)
        if d == 512 and FUSE_OUTPROJ_LN and M >= FUSE_OUTPROJ_LN_MIN_TOKENS:
            # out-projection + residual + ln_2 in one kernel: the fp32 residual stream is read and written once
            ops.linear_res_ln(buf.att, cache.get(blk.attn.out_proj.weight), _f32(blk.attn.out_proj.bias), x,
                              _f32(blk.ln_2.weight), _f32(blk.ln_2.bias), buf.xn)
        else:
            ops.linear(buf.att, cache.get(blk.attn.out_proj.weight), _f32(blk.attn.out_proj.bias), residual=x, out_f32=x)
            ops.layernorm(x, M, d, gamma=_f32(blk.ln_2.weight), beta=_f32(blk.ln_2.bias), L_in=L, out_bf16=buf.xn)
        ops.linear(buf.xn, cache.get(blk.mlp.c_fc.weight), _f32(blk.mlp.c_fc.bias), out_bf16=buf.h, act=ACT_QUICKGELU)
        # c_proj + residual, fused with the LayerNorm that consumes it: the next block's ln_1 (which also emits
        # stage i) or the post-encoder LayerNorm (stage S-1)
        last = i + 1 == S
        nxt = blocks[i + 1].ln_1 if not last else (post_ln if emit_final else None)
        emit = sink.stage_views(i) if (not last or emit_final) else {}
        xn_ready = False
        if nxt is not None and fusable(emit) and (not last or emit):
            ops.linear_res_ln_stage(buf.h, cache.get(blk.mlp.c_proj.weight), _f32(blk.mlp.c_proj.bias), x,
                                    _f32(nxt.weight), _f32(nxt.bias), None if last else buf.xn, L, l_split,
                                    emit.get("nrmA_bf16"), sink.strideA, emit.get("nrmB_bf16"), sink.strideB)
            xn_ready = not last
            final_done = last
        else:
            ops.linear(buf.h, cache.get(blk.mlp.c_proj.weight), _f32(blk.mlp.c_proj.bias), residual=x, out_f32=x)
    if emit_final and S >= 1 and not final_done:
        emit = sink.stage_views(S - 1)
        if post_ln is not None:
            ops.layernorm(x, M, d, gamma=_f32(post_ln.weight), beta=_f32(post_ln.bias), L_in=L, **sk, **emit)
        else:
            ops.layernorm(x, M, d, L_in=L, **sk, **emit)


def _mask_u8(key_padding_mask, B: int, L: int, device) -> Optional[torch.Tensor]:
    if key_padding_mask is None:
        return None
    m = key_padding_mask.to(device=device)
    if m.shape != (B, L):
        raise TanError(f"key_padding_mask must be [B={B}, L={L}], got {tuple(m.shape)}")
    return m.to(torch.uint8).contiguous()


class TemporalEncoder(nn.Module):
    """model/tfm_model.py:41-55.  forward(x [L,B,C], key_padding_mask [B,L] bool) -> list of S
    tensors [L,B,C]: [ln_1^(2)(x_1), ..., ln_1^(S)(x_{S-1}), x_S]."""

    def __init__(self, width: int, layers: int, heads: int):
        super().__init__()
        self.width = width
        self.layers = layers
        self.resblocks = nn.ModuleList([ResidualAttentionBlock_Step(width, heads) for _ in range(layers)])
        self._cache = _Bf16Cache()
        self._bufs = {}

    def buffers(self, M: int, device) -> _StackBuffers:
        key = (M, str(device))
        if key not in self._bufs:
            self._bufs[key] = _StackBuffers(M, self.width, device)
        return self._bufs[key]

    @torch.no_grad()
    def forward(self, x: torch.Tensor, key_padding_mask: torch.Tensor = None) -> List[torch.Tensor]:
        if not x.is_cuda:
            raise TanError("TemporalEncoder runs on a CUDA (sm_100a) device only; there is no CPU path")
        L, B, Cw = x.shape
        S = self.layers
        xb = x.detach().float().permute(1, 0, 2).contiguous().view(B * L, Cw)    # batch-first residual stream
        raw = torch.empty(max(S - 1, 0), B, L, Cw, dtype=torch.float32, device=x.device)
        sink = StageSink(S, l_split=L, strideA=L, rawA=raw if S > 1 else None, offA=B * L)
        # stage S-1 is x_S itself (model/tfm_model.py:54): no emission for it
        blocks = list(self.resblocks)
        kpm = _mask_u8(key_padding_mask, B, L, x.device)
        buf = self.buffers(B * L, x.device)
        run_encoder_stack(blocks, xb, kpm, B, L, self._cache, buf, sink, emit_final=False)
        outs = [raw[s].permute(1, 0, 2) for s in range(S - 1)]
        outs.append(xb.view(B, L, Cw).permute(1, 0, 2))
        return outs


class ResidualDecoderBlock_Step(nn.Module):
    """model/tfm_model.py:59-86 (unused by TAN; kept for API completeness)."""

    def __init__(self, d_model, n_head):
        super().__init__()
        if d_model != n_head * 64:
            raise TanError(f"head_dim must be 64 (d_model={d_model}, n_head={n_head})")
        self.self_attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = LayerNorm(d_model)
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.mlp = nn.Sequential(OrderedDict([
            ("c_fc", nn.Linear(d_model, d_model * 4)),
            ("gelu", QuickGELU()),
            ("c_proj", nn.Linear(d_model * 4, d_model)),
        ]))
        self.ln_2 = LayerNorm(d_model)
        self.ln_3 = LayerNorm(d_model)
        self.d_model = d_model
        self.n_head = n_head


class TemporalDecoder(nn.Module):
    """model/tfm_model.py:89-103: pre-LN self-attention, cross-attention over UN-normalised memory
    (:84), MLP; same stage-list contract as the encoder."""

    def __init__(self, width: int, layers: int, heads: int):
        super().__init__()
        self.width = width
        self.layers = layers
        self.resblocks = nn.ModuleList([ResidualDecoderBlock_Step(width, heads) for _ in range(layers)])
        self._cache = _Bf16Cache()

    @torch.no_grad()
    def forward(self, x, memory, tgt_key_padding_mask=None, memory_key_padding_mask=None):
        if not x.is_cuda:
            raise TanError("TemporalDecoder runs on a CUDA (sm_100a) device only; there is no CPU path")
        Lq, B, d = x.shape
        Lk = memory.shape[0]
        dev = x.device
        S = self.layers
        xb = x.detach().float().permute(1, 0, 2).contiguous().view(B * Lq, d)
        mem_bf = ops.cast_bf16(memory.detach().float().permute(1, 0, 2).contiguous().view(B * Lk, d))
        tk = _mask_u8(tgt_key_padding_mask, B, Lq, dev)
        mk = _mask_u8(memory_key_padding_mask, B, Lk, dev)
        bf = dict(dtype=torch.bfloat16, device=dev)
        xn = torch.empty(B * Lq, d, **bf)
        qkv = torch.empty(B * Lq, 3 * d, **bf)
        kv = torch.empty(B * Lk, 2 * d, **bf)
        att = torch.empty(B * Lq, d, **bf)
        h = torch.empty(B * Lq, 4 * d, **bf)
        raw = torch.empty(max(S - 1, 0), B, Lq, d, dtype=torch.float32, device=dev)
        c = self._cache
        for i, blk in enumerate(self.resblocks):
            H = blk.n_head
            emit = dict(rawA=raw.view(-1, d)[(i - 1) * B * Lq:]) if i >= 1 else {}
            ops.layernorm(xb, B * Lq, d, gamma=_f32(blk.ln_1.weight), beta=_f32(blk.ln_1.bias), L_in=Lq, out_bf16=xn,
                          l_split=Lq, strideA=Lq, **emit)
            ops.linear(xn, c.get(blk.self_attn.in_proj_weight), _f32(blk.self_attn.in_proj_bias), out_bf16=qkv)
            ops.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], tk, att, B, H, Lq, Lq)
            ops.linear(att, c.get(blk.self_attn.out_proj.weight), _f32(blk.self_attn.out_proj.bias), residual=xb, out_f32=xb)
            # cross-attention: q from ln_2(x), k/v from the raw memory (model/tfm_model.py:84)
            ops.layernorm(xb, B * Lq, d, gamma=_f32(blk.ln_2.weight), beta=_f32(blk.ln_2.bias), L_in=Lq, out_bf16=xn)
            w_in = c.get(blk.attn.in_proj_weight)
            b_in = _f32(blk.attn.in_proj_bias)
            ops.linear(xn, w_in[:d], b_in[:d], out_bf16=qkv[:, :d])
            ops.linear(mem_bf, w_in[d:], b_in[d:], out_bf16=kv)
            ops.attention(qkv[:, :d], kv[:, :d], kv[:, d:], mk, att, B, H, Lq, Lk)
            ops.linear(att, c.get(blk.attn.out_proj.weight), _f32(blk.attn.out_proj.bias), residual=xb, out_f32=xb)
            ops.layernorm(xb, B * Lq, d, gamma=_f32(blk.ln_3.weight), beta=_f32(blk.ln_3.bias), L_in=Lq, out_bf16=xn)
            ops.linear(xn, c.get(blk.mlp.c_fc.weight), _f32(blk.mlp.c_fc.bias), out_bf16=h, act=ACT_QUICKGELU)
            ops.linear(h, c.get(blk.mlp.c_proj.weight), _f32(blk.mlp.c_proj.bias), residual=xb, out_f32=xb)
        outs = [raw[s].permute(1, 0, 2) for s in range(S - 1)]
        outs.append(xb.view(B, Lq, d).permute(1, 0, 2))
        return outs


class PositionEmbeddingSine(nn.Module):
    """model/tfm_model.py:106-134: host-side constant table (no kernel)."""

    def __init__(self, num_pos_feats=64, temperature=10000, normalize=True, scale=None):
        super().__init__()
        self.num_pos_feats = num_pos_feats
        self.temperature = temperature
        self.normalize = normalize
        if scale is not None and normalize is False:
            raise ValueError("normalize should be True if scale is passed")
        self.scale = 2 * math.pi if scale is None else scale

    def forward(self, mask):
        assert mask is not None
        y = (~mask).cumsum(1, dtype=torch.float32)
        if self.normalize:
            y = y / (y[:, -1:] + 1e-6) * self.scale
        dim_t = torch.arange(self.num_pos_feats, dtype=torch.float32, device=mask.device)
        dim_t = self.temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / self.num_pos_feats)
        p = y[:, :, None] / dim_t
        p = torch.stack((p[:, :, 0::2].sin(), p[:, :, 1::2].cos()), dim=3).flatten(2)
        return p.permute(0, 2, 1)


def get_position_embedding_sine(feature_dim=512, num_features=1024, temperature=10000):
    """model/tfm_model.py:137-148: [num_features, feature_dim] interleaved sin/cos table with positions
    scaled to [0, 2*pi]."""
    pos = torch.arange(num_features)
    pos = pos / (pos[-1:] + 1e-6) * (2 * math.pi)
    dim_t = torch.arange(feature_dim, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / feature_dim)
    e = pos[:, None] / dim_t
    return torch.stack((e[:, 0::2].sin(), e[:, 1::2].cos()), dim=2).flatten(1)
