"""TemporalAligner / TwinTemporalAligner with the reference's API (model/tan_model.py) on top of the
sm_100a kernels.  Drop-in surface = the SUPERSET that train/main.py actually calls:
`forward(video, text, video_padding_mask=, lang_padding_mask=, text_timestamp=, abs_text_pos=,
interpolate_from=)`, `lang_model` as well as `bert`, `get_alignability(..., abs_text_pos)`.

Differences from the reference that are part of the design:
  * bf16 tensor-core math with fp32 accumulation / residuals; inference calls carry no autograd graph, training
    goes through `enable_autograd()` (train.py: taped forward + hand-written backward behind one autograd node);
  * the video pre-projection + LayerNorm is computed ONCE per forward (the reference computes it
    twice, model/tan_model.py:155 and :187);
  * `logits_dual` / `logits_joint` are `LazyLogits` handles by default (fused mode): the
    [B,S,T,B,N] tensor is only materialised (bf16, one write) if a caller indexes it; `get_loss`
    consumes the handle directly.  `materialize_logits=True` restores eager tensors;
  * `width` / `video_dim` are constructor arguments (reference hard-codes 512 / 1024) so that
    BASELINE config 4 (width 768, 12 heads) is expressible; heads = width / 64.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn
from torch.nn import LayerNorm

from . import ops
from ._lib import TanError
from .tfm_model import (StageSink, TemporalEncoder, _Bf16Cache, _f32, _mask_u8, get_position_embedding_sine,
                        run_encoder_stack)


class _TextBackboneSlot(nn.Module):
    """Placeholder for the text backbone (`Word2VecModel` / BERT, model/tan_model.py:37-40), which is
    upstream of the hot path and needs weights that are not part of either repository.  Assign a
    real module to `model.bert` / `model.lang_model` to use `train/main.py:58-60` unchanged."""

    def forward(self, *a, **k):
        raise TanError("no text backbone attached: pass lang_module=... or assign model.bert")


class LazyLogits:
    """Handle for a `[B, S, T, B, N]` cosine-logit tensor that has not been written to HBM.

    Holds the L2-normalised bf16 stage features; `get_loss` feeds them to the fused
    similarity+NCE kernel.  Anything that needs actual numbers (`x[:, -1]`, `x / 0.07`,
    `x.float()`, `torch.diagonal(...)` via `materialize()`) triggers ONE bf16 write of the tensor.
    """

    def __init__(self, vfeat: torch.Tensor, tfeat: torch.Tensor, shared_text: bool, N: int):
        self.vfeat = vfeat            # [B, S, T, d] bf16
        self.tfeat = tfeat            # [B*N, d] (shared by all stages) or [S, B*N, d] bf16
        self.shared_text = shared_text
        self.N = N
        self._dense: Optional[torch.Tensor] = None

    @property
    def shape(self):
        B, S, T, _ = self.vfeat.shape
        C = self.tfeat.shape[-2]
        return torch.Size([B, S, T, C // self.N, self.N])

    @property
    def device(self):
        return self.vfeat.device

    @property
    def dtype(self):
        return torch.bfloat16

    def dim(self):
        return 5

    def size(self, i=None):
        return self.shape if i is None else self.shape[i]

    def materialize(self) -> torch.Tensor:
        """bf16 [B, S, T, B, N], written once by tan_sim_nce_fwd's epilogue."""
        if self._dense is None:
            B, S, T, d = self.vfeat.shape
            C = self.tfeat.shape[-2]
            dev = self.device
            g = ops.sim_geom(B, S, T, C, self.N, d, 0)
            out = torch.empty(B, S, T, C // self.N, self.N, dtype=torch.bfloat16, device=dev)
            posbits = torch.zeros(B, T, (self.N + 31) // 32, dtype=torch.int32, device=dev)
            valid = torch.ones(C, dtype=torch.uint8, device=dev)
            rs = torch.empty(2, B * S * T, dtype=torch.float32, device=dev)
            cs = torch.empty(2, S, C, dtype=torch.float32, device=dev)
            ws = torch.empty(ops.sim_workspace_bytes(g), dtype=torch.uint8, device=dev)
            ops.sim_nce_fwd(self.vfeat, self.tfeat, 0 if self.shared_text else C * d, g, posbits, valid, out,
                            rs, cs, ws)
            self._dense = out
        return self._dense

    def __getitem__(self, idx):
        return self.materialize()[idx]

    def __truediv__(self, other):
        return self.materialize() / other

    def __mul__(self, other):
        return self.materialize() * other

    def float(self):
        return self.materialize().float()

    def detach(self):
        return self

    def to(self, *a, **k):
        return self.materialize().to(*a, **k)

    def __repr__(self):
        return f"LazyLogits(shape={tuple(self.shape)}, materialised={self._dense is not None})"


class TemporalAligner(nn.Module):
    """model/tan_model.py:13-312."""

    def __init__(self, num_encoder_layers=2, num_decoder_layers=2, sim='cos', language_model='word2vec',
                 pos_enc='learned', use_text_pos_enc=0, return_dual_feature=1, random_pos_start=1,
                 use_alignability_head=0, width=512, video_dim=1024, lang_module: Optional[nn.Module] = None,
                 materialize_logits: bool = False):
        super().__init__()
        if num_encoder_layers < 1 or num_decoder_layers < 1:
            raise TanError("num_encoder_layers and num_decoder_layers must be >= 1 "
                           "(the reference itself fails at init with 0 joint layers, model/tan_model.py:85)")
        if width % 128 != 0:
            raise TanError("width must be a multiple of 128 (heads = width / 64)")
        self.num_encoder_layers = num_encoder_layers
        self.num_decoder_layers = num_decoder_layers
        self.sim = sim
        self.pos_enc = pos_enc
        self.language_model = language_model
        self.use_text_pos_enc = use_text_pos_enc
        self.return_dual_feature = return_dual_feature
        self.random_pos_start = random_pos_start
        self.use_alignability_head = use_alignability_head
        self.width = width
        self.materialize_logits = materialize_logits
        text_embed_dim = {'bert': 768, 'word2vec': 512}

        self.bert = lang_module if lang_module is not None else _TextBackboneSlot()
        heads = width // 64
        self.video_temporal_encoder = TemporalEncoder(width=width, layers=num_encoder_layers, heads=heads)
        self.joint_temporal_encoder = TemporalEncoder(width=width, layers=num_decoder_layers, heads=heads)
        self.video_pre_proj = nn.Linear(video_dim, width, bias=False)
        self.text_pre_proj = nn.Linear(text_embed_dim[language_model], width, bias=False)
        self.ln_text_init = LayerNorm(width)
        self.ln_video_init = LayerNorm(width)
        self.ln_position_init = LayerNorm(width)
        self.ln_video_post_enc = LayerNorm(width)
        self.ln_joint_post_enc = LayerNorm(width)
        if self.pos_enc == 'learned':
            self.temporal_pos_embed = nn.Parameter(torch.empty(1024, width))
            nn.init.normal_(self.temporal_pos_embed, std=0.01)
        elif self.pos_enc == 'sine':
            self.register_buffer('temporal_pos_embed', get_position_embedding_sine(width, 1024))
        else:
            raise TanError(f"unknown pos_enc {pos_enc!r}")
        self.text_temporal_pos_embed = nn.Parameter(torch.empty(1024, width))
        nn.init.normal_(self.text_temporal_pos_embed, std=0.01)
        self.mlp = nn.Linear(width, width)          # present in the reference state-dict, never used (:68)
        if self.use_alignability_head:
            self.binary_head = nn.Linear(width, 1)
            nn.init.normal_(self.binary_head.weight, std=0.01)
            nn.init.zeros_(self.binary_head.bias)
        self.initialize_parameters()
        self._cache = _Bf16Cache()
        self._scratch = {}
        self._streams = {}
        self._graphs = {}
        self._graphs_on = False
        self.two_streams = True      # run the dual and joint stacks concurrently (see forward)
        from .train import AUTOGRAD_DEFAULT
        self._autograd_on = AUTOGRAD_DEFAULT

    # the training driver calls `model.lang_model` (train/main.py:58) while the reference class
    # names it `bert` (model/tan_model.py:38-40); expose both
    @property
    def lang_model(self):
        return self.bert

    def initialize_parameters(self):
        """model/tan_model.py:76-97 (all stds derive from the JOINT encoder's width/layers)."""
        nn.init.normal_(self.video_pre_proj.weight, std=0.01)
        nn.init.normal_(self.text_pre_proj.weight, std=0.01)
        nn.init.zeros_(self.mlp.bias)
        nn.init.normal_(self.mlp.weight, std=0.01)
        w, n = self.joint_temporal_encoder.width, self.joint_temporal_encoder.layers
        proj_std = (w ** -0.5) * ((2 * n) ** -0.5)
        attn_std = w ** -0.5
        fc_std = (2 * w) ** -0.5
        for enc in (self.video_temporal_encoder, self.joint_temporal_encoder):
            for block in enc.resblocks:
                nn.init.normal_(block.attn.in_proj_weight, std=attn_std)
                nn.init.normal_(block.attn.out_proj.weight, std=proj_std)
                nn.init.normal_(block.mlp.c_fc.weight, std=fc_std)
                nn.init.normal_(block.mlp.c_proj.weight, std=proj_std)

    # ------------------------------------------------------------------------------------------
    # building blocks
    # ------------------------------------------------------------------------------------------
    def _buf(self, name, shape, dtype, device):
        key = (name, tuple(shape), dtype, str(device))
        t = self._scratch.get(key)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=device)
            self._scratch[key] = t
        return t

    def _side_stream(self, dev):
        st = self._streams.get(str(dev))
        if st is None:
            st = torch.cuda.Stream(device=dev)
            self._streams[str(dev)] = st
        return st

    def _check_device(self, t):
        if not t.is_cuda:
            raise TanError("TemporalAligner runs on a CUDA (sm_100a) device only; there is no CPU path")

    def _pos_start(self, L):
        """model/tan_model.py:162-165: one draw from the GLOBAL numpy RNG per call site."""
        return int(np.random.randint(0, int(L / 2))) if self.random_pos_start else 0

    def _pos_ln(self, table, L, pos_start, interpolate_from, name):
        """ln_position_init(pos[s:s+L]) or of the linearly interpolated table -> [L, d] fp32
        (model/tan_model.py:157-167).  Batch-independent, so computed once per forward."""
        d = self.width
        if interpolate_from:
            src = table.detach()[None, 0:int(interpolate_from), :].float()
            pos = F.interpolate(src.transpose(1, 2), size=L, mode='linear', align_corners=False).transpose(1, 2)[0]
            pos = pos.contiguous()
        else:
            if pos_start + L > table.shape[0]:
                raise TanError(f"positional table overrun: start {pos_start} + L {L} > {table.shape[0]} "
                               "(use random_pos_start=0 or interpolate_from)")
            pos = _f32(table)[pos_start:pos_start + L]
        out = torch.empty(L, d, dtype=torch.float32, device=table.device)
        ops.layernorm(pos, L, d, gamma=_f32(self.ln_position_init.weight), beta=_f32(self.ln_position_init.bias),
                      out_f32=out)
        return out

    def _video_preproj(self, video_embed):
        """video_pre_proj(x) -> [B*T, d] fp32 (GEMM part of model/tan_model.py:155 / :187), once."""
        B, T, Din = video_embed.shape
        dev = video_embed.device
        v = video_embed.detach()
        if v.dtype == torch.bfloat16:
            vb = v.contiguous().view(B * T, Din)
        else:
            vb = self._buf("video_bf16", (B * T, Din), torch.bfloat16, dev)
            ops.cast_bf16(v.float().contiguous().view(B * T, Din), vb)
        pre = self._buf("video_pre", (B * T, self.width), torch.float32, dev)
        ops.linear(vb, self._cache.get(self.video_pre_proj.weight), out_f32=pre)
        return pre

    def _text_preproj(self, lang_embed):
        B, N, Dt = lang_embed.shape
        dev = lang_embed.device
        t = lang_embed.detach()
        if t.dtype == torch.bfloat16:
            tb = t.contiguous().view(B * N, Dt)
        else:
            tb = self._buf("text_bf16", (B * N, Dt), torch.bfloat16, dev)
            ops.cast_bf16(t.float().contiguous().view(B * N, Dt), tb)
        pre = self._buf("text_pre", (B * N, self.width), torch.float32, dev)
        ops.linear(tb, self._cache.get(self.text_pre_proj.weight), out_f32=pre)
        return pre

    def _run_video_stack(self, pre, B, T, kpm_v, pos_ln, want_raw, want_nrm):
        """Video stack on pre-projected tokens -> (raw [B,E,T,d] fp32 | None, nrm [B,E,T,d] bf16 | None)."""
        d, E, dev = self.width, self.num_encoder_layers, pre.device
        x = self._buf("x_video", (B * T, d), torch.float32, dev)
        ops.layernorm(pre, B * T, d, gamma=_f32(self.ln_video_init.weight), beta=_f32(self.ln_video_init.bias),
                      add=pos_ln, add_rows=T, L_in=T, out_f32=x)
        raw = torch.empty(B, E, T, d, dtype=torch.float32, device=dev) if want_raw else None
        nrm = torch.empty(B, E, T, d, dtype=torch.bfloat16, device=dev) if want_nrm else None
        sink = StageSink(E, l_split=T, strideA=E * T, rawA=raw, nrmA_bf16=nrm, offA=T)
        enc = self.video_temporal_encoder
        run_encoder_stack(list(enc.resblocks), x, kpm_v, B, T, enc._cache, enc.buffers(B * T, dev), sink,
                          post_ln=self.ln_video_post_enc)
        return raw, nrm

    def _run_joint_stack(self, pre_v, pre_t, B, T, N, kpm_v, kpm_t, pos_ln_v, pos_ln_t, want_raw_v, want_raw_t,
                         want_nrm, nrm_t_out=None):
        """Joint stack over [video ; text] tokens (model/tan_model.py:182-209)."""
        d, D, dev = self.width, self.num_decoder_layers, pre_v.device
        L = T + N
        x = self._buf("x_joint", (B * L, d), torch.float32, dev)
        ops.layernorm(pre_v, B * T, d, gamma=_f32(self.ln_video_init.weight), beta=_f32(self.ln_video_init.bias),
                      add=pos_ln_v, add_rows=T, L_in=T, L_out=L, l_off=0, out_f32=x)
        ops.layernorm(pre_t, B * N, d, gamma=_f32(self.ln_text_init.weight), beta=_f32(self.ln_text_init.bias),
                      add=pos_ln_t, add_rows=N, L_in=N, L_out=L, l_off=T, out_f32=x)
        if kpm_v is None and kpm_t is None:
            kpm = None
        else:
            kv = kpm_v if kpm_v is not None else torch.zeros(B, T, dtype=torch.uint8, device=dev)
            kt = kpm_t if kpm_t is not None else torch.zeros(B, N, dtype=torch.uint8, device=dev)
            kpm = torch.cat((kv, kt), dim=1).contiguous()          # model/tan_model.py:203
        raw_v = torch.empty(B, D, T, d, dtype=torch.float32, device=dev) if want_raw_v else None
        raw_t = torch.empty(D, B, N, d, dtype=torch.float32, device=dev) if want_raw_t else None
        nrm_v = torch.empty(B, D, T, d, dtype=torch.bfloat16, device=dev) if want_nrm else None
        nrm_t = (nrm_t_out if nrm_t_out is not None else
                 torch.empty(D, B * N, d, dtype=torch.bfloat16, device=dev)) if want_nrm else None
        sink = StageSink(D, l_split=T, strideA=D * T, strideB=N, rawA=raw_v, rawB=raw_t, nrmA_bf16=nrm_v,
                         nrmB_bf16=nrm_t, offA=T, offB=B * N)
        enc = self.joint_temporal_encoder
        run_encoder_stack(list(enc.resblocks), x, kpm, B, L, enc._cache, enc.buffers(B * L, dev), sink,
                          post_ln=self.ln_joint_post_enc)
        return raw_v, raw_t, nrm_v, nrm_t

    def _text_features(self, pre_t, B, N, want_raw, want_nrm_bf16, want_nrm_f32, nrm_out=None):
        """ln_text_init(text_pre_proj(t)) (model/tan_model.py:231-234) + its L2-normalised copies."""
        d, dev = self.width, pre_t.device
        raw = torch.empty(B, N, d, dtype=torch.float32, device=dev) if want_raw else None
        nb = (nrm_out if nrm_out is not None else
              torch.empty(B * N, d, dtype=torch.bfloat16, device=dev)) if want_nrm_bf16 else None
        nf = torch.empty(B, N, d, dtype=torch.float32, device=dev) if want_nrm_f32 else None
        ops.layernorm(pre_t, B * N, d, gamma=_f32(self.ln_text_init.weight), beta=_f32(self.ln_text_init.bias),
                      L_in=N, l_split=N, strideA=N, rawA=raw, nrmA_bf16=nb, nrmA_f32=nf)
        return raw, nb, nf

    def _binary_head(self, feat):
        """binary_head (Linear(d, 1)) on fp32 features [..., d] -> [..., 1] (model/tan_model.py:146-148).
        A d-long dot product per sentence (B*N*(1+D) rows): host-side torch matmul on the tiny head,
        not on the measured path."""
        return feat @ self.binary_head.weight.detach().t() + self.binary_head.bias.detach()

    # ------------------------------------------------------------------------------------------
    # reference API
    # ------------------------------------------------------------------------------------------
    def enable_cuda_graphs(self, enabled: bool = True) -> None:
        """Replay `forward` from a CUDA graph captured per input shape (about 100 launches -> one).
        Eligible calls: fp32 CUDA inputs, `random_pos_start=0`, no `interpolate_from`.  The returned
        tensors are the graph's static outputs: they are overwritten by the next `forward` call with
        the same shapes, so consume them (e.g. `get_loss`) before calling `forward` again."""
        self._graphs_on = bool(enabled)
        if not enabled:
            self._graphs.clear()

    def _forward_graphed(self, video_embed, lang_embed, video_padding_mask, lang_padding_mask):
        B, T, Din = video_embed.shape
        N, Dt = lang_embed.shape[1], lang_embed.shape[2]
        dev = video_embed.device
        key = (B, T, Din, N, Dt, video_padding_mask is None, lang_padding_mask is None, str(dev))
        ent = self._graphs.get(key)
        if ent is None:
            st = {"video": torch.empty(B, T, Din, dtype=torch.float32, device=dev),
                  "text": torch.empty(B, N, Dt, dtype=torch.float32, device=dev),
                  "vpm": None if video_padding_mask is None else torch.zeros(B, T, dtype=torch.bool, device=dev),
                  "tpm": None if lang_padding_mask is None else torch.zeros(B, N, dtype=torch.bool, device=dev)}
            st["video"].copy_(video_embed)
            st["text"].copy_(lang_embed)
            if st["vpm"] is not None:
                st["vpm"].copy_(video_padding_mask.to(dev).bool())
            if st["tpm"] is not None:
                st["tpm"].copy_(lang_padding_mask.to(dev).bool())
            self._forward_impl(st["video"], st["text"], st["vpm"], st["tpm"])      # eager warm-up: scratch, shadows
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self._forward_impl(st["video"], st["text"], st["vpm"], st["tpm"])
            ent = (g, st, out)
            self._graphs[key] = ent
        g, st, out = ent
        for c in (self._cache, self.video_temporal_encoder._cache, self.joint_temporal_encoder._cache):
            c.refresh()                                   # weights changed since capture -> same shadow buffers
        st["video"].copy_(video_embed, non_blocking=True)
        st["text"].copy_(lang_embed, non_blocking=True)
        if st["vpm"] is not None:
            st["vpm"].copy_(video_padding_mask, non_blocking=True)
        if st["tpm"] is not None:
            st["tpm"].copy_(lang_padding_mask, non_blocking=True)
        g.replay()
        res = dict(out)
        for k in ("logits_dual", "logits_joint"):         # fresh handles: a cached dense copy would be stale
            if isinstance(res[k], LazyLogits):
                res[k] = LazyLogits(res[k].vfeat, res[k].tfeat, res[k].shared_text, res[k].N)
        return res

    def enable_autograd(self, enabled: bool = True) -> None:
        """Training step: while enabled, a `forward` called in train mode with grad enabled keeps the activations
        (train.py) and `get_loss(...)['loss'].backward()` fills `.grad` of the parameters through the hand-written
        backward pass.  Round-1 status: opt-in (TAN_AUTOGRAD=1 makes it the default), first correct path."""
        self._autograd_on = bool(enabled)

    def forward(self, video_embed, lang_embed, video_padding_mask=None, lang_padding_mask=None,
                text_timestamp=None, abs_text_pos=None, interpolate_from=None):
        """model/tan_model.py:100-149 (+ the `abs_text_pos` keyword train/main.py:86 passes)."""
        self._check_device(video_embed)
        if (self._autograd_on and self.training and torch.is_grad_enabled()
                and any(p.requires_grad for p in self.parameters())):
            from . import train
            with torch.no_grad():
                return train.forward_train(self, video_embed, lang_embed, video_padding_mask, lang_padding_mask,
                                           interpolate_from)
        with torch.no_grad():
            return self._forward_dispatch(video_embed, lang_embed, video_padding_mask, lang_padding_mask,
                                          interpolate_from)

    def _forward_dispatch(self, video_embed, lang_embed, video_padding_mask, lang_padding_mask, interpolate_from):
        if (self._graphs_on and not self.random_pos_start and not interpolate_from
                and video_embed.dtype == torch.float32 and lang_embed.dtype == torch.float32 and lang_embed.is_cuda
                and not torch.cuda.is_current_stream_capturing()):
            return self._forward_graphed(video_embed, lang_embed, video_padding_mask, lang_padding_mask)
        return self._forward_impl(video_embed, lang_embed, video_padding_mask, lang_padding_mask, interpolate_from)

    @torch.no_grad()
    def _forward_impl(self, video_embed, lang_embed, video_padding_mask=None, lang_padding_mask=None,
                      interpolate_from=None):
        B, T, _ = video_embed.shape
        N = lang_embed.shape[1]
        dev = video_embed.device
        kpm_v = _mask_u8(video_padding_mask, B, T, dev)
        kpm_t = _mask_u8(lang_padding_mask, B, N, dev)
        head = bool(self.use_alignability_head)

        # RNG draws in the reference's call order: video stack (:163), text-with-time (:224), joint (:195)
        ps_v = self._pos_start(T) if not interpolate_from else 0
        pos_ln_v = self._pos_ln(self.temporal_pos_embed, T, ps_v, interpolate_from, "v")
        pre_v = self._video_preproj(video_embed)
        pre_t = self._text_preproj(lang_embed)

        pos_ln_t = None
        if self.use_text_pos_enc:
            ps_t = self._pos_start(N) if not interpolate_from else 0
            pos_ln_t = self._pos_ln(self.text_temporal_pos_embed, N, ps_t, interpolate_from, "t")
        ps_j = self._pos_start(T) if not interpolate_from else 0
        pos_ln_j = pos_ln_v if (ps_j == ps_v) else self._pos_ln(self.temporal_pos_embed, T, ps_j, None, "j")

        # The video (dual) stack and the joint stack only share the pre-projections, so they run on two
        # streams: at 32 clips per GPU each kernel is short and leaves SMs idle at its head and tail
        # (one wave of tiles, prologue, drain); the other stack's kernels fill those gaps.
        main = torch.cuda.current_stream(dev)
        side = self._side_stream(dev) if self.two_streams else main
        # the L2-normalised text features of the dual model and of every joint stage live in ONE buffer
        # [1 + D, B*N, d]: the multi-GPU exchange gathers it without a packing copy (loss.pack_text_features)
        tpack = torch.empty(1 + self.num_decoder_layers, B * N, self.width, dtype=torch.bfloat16, device=dev)
        if side is not main:
            side.wait_stream(main)
            tpack.record_stream(side)
        with torch.cuda.stream(side):
            _, jt_raw, vfeat_joint, tfeat_joint = self._run_joint_stack(
                pre_v, pre_t, B, T, N, kpm_v, kpm_t, pos_ln_j, pos_ln_t, want_raw_v=False, want_raw_t=head,
                want_nrm=True, nrm_t_out=tpack[1:])
        _, vfeat_dual = self._run_video_stack(pre_v, B, T, kpm_v, pos_ln_v, want_raw=False, want_nrm=True)
        text_raw, tfeat_dual, tfeat_dual_f32 = self._text_features(
            pre_t, B, N, want_raw=head, want_nrm_bf16=True, want_nrm_f32=bool(self.return_dual_feature),
            nrm_out=tpack[0])
        if side is not main:
            main.wait_stream(side)
            for t_ in (jt_raw, vfeat_joint, tfeat_joint):
                if t_ is not None:
                    t_.record_stream(main)

        logits_dual = LazyLogits(vfeat_dual, tfeat_dual, shared_text=True, N=N)
        logits_joint = LazyLogits(vfeat_joint, tfeat_joint, shared_text=False, N=N)
        if self.materialize_logits:
            logits_dual, logits_joint = logits_dual.materialize(), logits_joint.materialize()
        out = {'logits_dual': logits_dual, 'logits_joint': logits_joint}
        if self.return_dual_feature:
            out['dual_feature_video'] = vfeat_dual                 # [B,E,T,d] L2-normalised (bf16)
            out['dual_feature_text'] = tfeat_dual_f32              # [B,N,d] L2-normalised
        if head:
            out['dual_logits_alignability'] = self._binary_head(text_raw)
            out['joint_logits_alignability'] = self._binary_head(jt_raw.permute(1, 0, 2, 3))   # [B,D,N,1]
        return out

    @torch.no_grad()
    def get_visual_feature(self, video_embed, video_padding_mask, interpolate_from=None):
        """model/tan_model.py:152-179 -> [B, S, T, C] fp32 (last stage through ln_video_post_enc)."""
        self._check_device(video_embed)
        B, T, _ = video_embed.shape
        kpm_v = _mask_u8(video_padding_mask, B, T, video_embed.device)
        ps = self._pos_start(T) if not interpolate_from else 0
        pos_ln = self._pos_ln(self.temporal_pos_embed, T, ps, interpolate_from, "v")
        raw, _ = self._run_video_stack(self._video_preproj(video_embed), B, T, kpm_v, pos_ln, want_raw=True,
                                       want_nrm=False)
        return raw

    @torch.no_grad()
    def get_textual_feature(self, lang_embed):
        """model/tan_model.py:231-234 -> [B, N, C] fp32."""
        self._check_device(lang_embed)
        B, N, _ = lang_embed.shape
        raw, _, _ = self._text_features(self._text_preproj(lang_embed), B, N, True, False, False)
        return raw

    @torch.no_grad()
    def get_textual_feature_with_time(self, lang_embed, text_timestamp=None, interpolate_from=None):
        """model/tan_model.py:212-228."""
        self._check_device(lang_embed)
        B, N, _ = lang_embed.shape
        d = self.width
        ps = self._pos_start(N) if not interpolate_from else 0
        pos_ln = self._pos_ln(self.text_temporal_pos_embed, N, ps, interpolate_from, "t")
        out = torch.empty(B, N, d, dtype=torch.float32, device=lang_embed.device)
        ops.layernorm(self._text_preproj(lang_embed), B * N, d, gamma=_f32(self.ln_text_init.weight),
                      beta=_f32(self.ln_text_init.bias), add=pos_ln, add_rows=N, L_in=N, out_f32=out)
        return out

    @torch.no_grad()
    def get_joint_feature(self, video_embed, video_padding_mask, lang_embed_with_time, lang_padding_mask,
                          interpolate_from=None):
        """model/tan_model.py:182-209.  NOTE the reference takes the already projected text features
        here; so does this method ([B,N,C] fp32) -> (video [B,S,T,C], text [B,S,N,C]) fp32."""
        self._check_device(video_embed)
        B, T, _ = video_embed.shape
        N = lang_embed_with_time.shape[1]
        dev = video_embed.device
        d, D = self.width, self.num_decoder_layers
        L = T + N
        kpm_v = _mask_u8(video_padding_mask, B, T, dev)
        kpm_t = _mask_u8(lang_padding_mask, B, N, dev)
        ps = self._pos_start(T) if not interpolate_from else 0
        pos_ln = self._pos_ln(self.temporal_pos_embed, T, ps, interpolate_from, "v")
        pre_v = self._video_preproj(video_embed)
        x = self._buf("x_joint", (B * L, d), torch.float32, dev)
        ops.layernorm(pre_v, B * T, d, gamma=_f32(self.ln_video_init.weight), beta=_f32(self.ln_video_init.bias),
                      add=pos_ln, add_rows=T, L_in=T, L_out=L, l_off=0, out_f32=x)
        tf = lang_embed_with_time.detach().float().contiguous().view(B * N, d)
        ops.layernorm(tf, B * N, d, L_in=N, L_out=L, l_off=T, out_f32=x)          # identity scatter (torch.cat, :201)
        kv = kpm_v if kpm_v is not None else torch.zeros(B, T, dtype=torch.uint8, device=dev)
        kt = kpm_t if kpm_t is not None else torch.zeros(B, N, dtype=torch.uint8, device=dev)
        kpm = torch.cat((kv, kt), dim=1).contiguous()
        raw_v = torch.empty(B, D, T, d, dtype=torch.float32, device=dev)
        raw_t = torch.empty(D, B, N, d, dtype=torch.float32, device=dev)
        sink = StageSink(D, l_split=T, strideA=D * T, strideB=N, rawA=raw_v, rawB=raw_t, offA=T, offB=B * N)
        enc = self.joint_temporal_encoder
        run_encoder_stack(list(enc.resblocks), x, kpm, B, L, enc._cache, enc.buffers(B * L, dev), sink,
                          post_ln=self.ln_joint_post_enc)
        return raw_v, raw_t.permute(1, 0, 2, 3)

    def _split_interp(self, interpolate_from):
        if isinstance(interpolate_from, (list, tuple)):
            assert len(interpolate_from) == 2
            return interpolate_from[0], interpolate_from[1]
        return interpolate_from, None

    def _eval_sim(self, vfeat, tfeat, shared_text, B, S, T, N):
        """Per-video similarity [B, S, T, N] = the diagonal blocks of the [B,S,T,B,N] matrix
        (einsum 'bstc,b(s)kc->bstk', model/tan_model.py:261-262,:280-281): tan_own_clip_sim computes only those
        blocks (fp32), instead of materialising all B^2 and indexing the diagonal."""
        return ops.own_clip_sim(vfeat, tfeat, shared_text, B, S, T, N, self.width)

    @torch.no_grad()
    def get_text_visual_sim_joint(self, video_embed, lang_embed, interpolate_from=None):
        """model/tan_model.py:237-263 -> [B, S, T, N] fp32."""
        self._check_device(video_embed)
        interpolate_from, t_if = self._split_interp(interpolate_from)
        B, T, _ = video_embed.shape
        N = lang_embed.shape[1]
        pos_ln_t = None
        if self.use_text_pos_enc:
            ps_t = self._pos_start(N) if not t_if else 0
            pos_ln_t = self._pos_ln(self.text_temporal_pos_embed, N, ps_t, t_if, "t")
        ps = self._pos_start(T) if not interpolate_from else 0
        pos_ln_v = self._pos_ln(self.temporal_pos_embed, T, ps, interpolate_from, "v")
        pre_v, pre_t = self._video_preproj(video_embed), self._text_preproj(lang_embed)
        _, _, nv, nt = self._run_joint_stack(pre_v, pre_t, B, T, N, None, None, pos_ln_v, pos_ln_t, False, False, True)
        return self._eval_sim(nv, nt, False, B, self.num_decoder_layers, T, N)

    @torch.no_grad()
    def get_text_visual_sim_dual(self, video_embed, lang_embed, interpolate_from=None):
        """model/tan_model.py:266-283 -> [B, S, T, N] fp32."""
        self._check_device(video_embed)
        B, T, _ = video_embed.shape
        N = lang_embed.shape[1]
        ps = self._pos_start(T) if not interpolate_from else 0
        pos_ln = self._pos_ln(self.temporal_pos_embed, T, ps, interpolate_from, "v")
        _, nv = self._run_video_stack(self._video_preproj(video_embed), B, T, None, pos_ln, False, True)
        _, nt, _ = self._text_features(self._text_preproj(lang_embed), B, N, False, True, False)
        return self._eval_sim(nv, nt, True, B, self.num_encoder_layers, T, N)

    @torch.no_grad()
    def get_alignability(self, video_embed, lang_embed, interpolate_from=None, abs_text_pos=None):
        """model/tan_model.py:286-312 (+ the 4th positional train/main.py:187 passes)."""
        self._check_device(video_embed)
        if not self.use_alignability_head:
            raise TanError("get_alignability needs use_alignability_head=1")
        interpolate_from, t_if = self._split_interp(interpolate_from)
        B, T, _ = video_embed.shape
        N = lang_embed.shape[1]
        pos_ln_t = None
        if self.use_text_pos_enc:
            ps_t = self._pos_start(N) if not t_if else 0
            pos_ln_t = self._pos_ln(self.text_temporal_pos_embed, N, ps_t, t_if, "t")
        ps = self._pos_start(T) if not interpolate_from else 0
        pos_ln_v = self._pos_ln(self.temporal_pos_embed, T, ps, interpolate_from, "v")
        pre_v, pre_t = self._video_preproj(video_embed), self._text_preproj(lang_embed)
        _, jt_raw, _, _ = self._run_joint_stack(pre_v, pre_t, B, T, N, None, None, pos_ln_v, pos_ln_t, False, True,
                                                False)
        text_raw, _, _ = self._text_features(pre_t, B, N, True, False, False)
        return {'alignability-dual': self._binary_head(text_raw),
                'alignability-joint': self._binary_head(jt_raw.permute(1, 0, 2, 3))}


class TwinTemporalAligner(nn.Module):
    """model/tan_model.py:315-351: online + EMA target copy."""

    def __init__(self, m=0.999, *args, **kwargs):
        super().__init__()
        self.m = m
        self.online = TemporalAligner(*args, **kwargs)
        self.target = TemporalAligner(*args, **kwargs)
        self._copy_param()
        self.bert = self.online.bert
        self.get_visual_feature = self.online.get_visual_feature
        self.get_joint_feature = self.online.get_joint_feature
        self.get_textual_feature_with_time = self.online.get_textual_feature_with_time
        self.get_textual_feature = self.online.get_textual_feature
        # the reference reads a non-existent `get_text_visual_sim` here (:328, AttributeError);
        # train/main.py:178-181 needs the joint and dual variants
        self.get_text_visual_sim_joint = self.online.get_text_visual_sim_joint
        self.get_text_visual_sim_dual = self.online.get_text_visual_sim_dual
        self.get_alignability = self.online.get_alignability
        self.target.random_pos_start = 0

    @property
    def lang_model(self):
        return self.bert

    @torch.no_grad()
    def _copy_param(self):
        """model/tan_model.py:334-338.  In-place on the Parameters themselves (not on `.data`): the copy bumps
        `_version`, which is what tells the target's bf16 weight shadows (tfm_model._Bf16Cache) to re-cast."""
        for po, pt in zip(self.online.parameters(), self.target.parameters()):
            pt.requires_grad = False
            pt.copy_(po.detach())

    @torch.no_grad()
    def _momentum_update(self):
        """model/tan_model.py:340-344 (p_t = m p_t + (1 - m) p_o) as ONE multi-tensor kernel launch
        (tan_ema_update) instead of ~160 tiny kernels.  The target Parameters are updated in place THROUGH torch's
        version counter (the reference reassigns `.data`; an update through `.data` would leave `_version` and the
        pointer unchanged and the bf16 weight shadows the forward reads would silently stay at their old values)."""
        from . import optim
        optim.ema_update(list(self.target.parameters()), list(self.online.parameters()), self.m)

    def enable_cuda_graphs(self, enabled: bool = True) -> None:
        self.online.enable_cuda_graphs(enabled)
        self.target.enable_cuda_graphs(enabled)

    def enable_autograd(self, enabled: bool = True) -> None:
        """Only the online network trains (model/tan_model.py:334-338 freezes the target)."""
        self.online.enable_autograd(enabled)

    def forward(self, *args, **kwargs):
        return self.online(*args, **kwargs)

    @torch.no_grad()
    def forward_from_ema(self, *args, **kwargs):
        return self.target(*args, **kwargs)
