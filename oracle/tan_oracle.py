"""CPU oracle: an fp32 restatement of the TAN hot path (forward + MIL-NCE loss).

TEST INFRASTRUCTURE.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` leg may import this file; the product package never does (it fails loudly if
its CUDA library is missing instead of falling back here).

Parity pinning: the reference has NO tests or golden vectors for this path (the
only known-answer is the `circulant` docstring, train/loss.py:19-20).  The oracle is therefore
pinned against OUTPUTS OF THE REFERENCE ITSELF: `oracle/make_golden.py` imports the unmodified
reference modules (oracle/ref_loader.py) in the build container and commits their fp32 CPU outputs
under `tests/golden/`; `tests/test_oracle.py` checks this restatement against those fixtures on
every run and against the live reference when /root/reference is present.

The arithmetic the reference delegates to torch (nn.MultiheadAttention, LayerNorm, F.linear,
einsum, logsumexp; torch is unpinned by the reference, here 2.11.0) is restated with explicit
matmul / softmax / mean-var formulas, batch-first.  Each function cites what it follows.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

TEMPERATURE = 0.07  # train/loss.py:65-67


def _t(x, dtype=torch.float32):
    if isinstance(x, torch.Tensor):
        return x.to(dtype)
    return torch.from_numpy(np.ascontiguousarray(x)).to(dtype)


def layer_norm(x, w, b, eps: float = 1e-5):
    """torch.nn.LayerNorm over the last dim (biased variance, eps inside the sqrt)."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def quick_gelu(x):
    """model/tfm_model.py:11-13."""
    return x * torch.sigmoid(1.702 * x)


def self_attention(xn, w_in, b_in, w_out, b_out, key_padding_mask, heads: int):
    """nn.MultiheadAttention(d, heads)(xn, xn, xn, key_padding_mask=kpm, need_weights=False)
    as called at model/tfm_model.py:30-32: packed in-projection, scale 1/sqrt(hd), additive -inf
    on ignored keys, no dropout, out-projection.  xn: [B, L, C]; kpm: [B, L] bool (True = ignore).
    """
    B, L, C = xn.shape
    hd = C // heads
    qkv = xn @ w_in.t() + b_in
    q, k, v = qkv.split(C, dim=-1)
    q = q.view(B, L, heads, hd).transpose(1, 2)
    k = k.view(B, L, heads, hd).transpose(1, 2)
    v = v.view(B, L, heads, hd).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    if key_padding_mask is not None:
        s = s.masked_fill(key_padding_mask[:, None, None, :], float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, L, C)
    return o @ w_out.t() + b_out


def cross_attention(xq, mem, w_in, b_in, w_out, b_out, key_padding_mask, heads: int):
    """nn.MultiheadAttention with query != key = value (model/tfm_model.py:78-80)."""
    B, Lq, C = xq.shape
    Lk = mem.shape[1]
    hd = C // heads
    q = xq @ w_in[:C].t() + b_in[:C]
    k = mem @ w_in[C:2 * C].t() + b_in[C:2 * C]
    v = mem @ w_in[2 * C:].t() + b_in[2 * C:]
    q = q.view(B, Lq, heads, hd).transpose(1, 2)
    k = k.view(B, Lk, heads, hd).transpose(1, 2)
    v = v.view(B, Lk, heads, hd).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    if key_padding_mask is not None:
        s = s.masked_fill(key_padding_mask[:, None, None, :], float("-inf"))
    o = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B, Lq, C)
    return o @ w_out.t() + b_out


def encoder_stack(x, key_padding_mask, sd: Dict[str, torch.Tensor], prefix: str, layers: int,
                  heads: int) -> List[torch.Tensor]:
    """TemporalEncoder.forward (model/tfm_model.py:48-55) over ResidualAttentionBlock_Step
    (:34-38).  Returns the S stage tensors [LN1_2(x_1), ..., LN1_S(x_{S-1}), x_S], each [B, L, C].
    """
    stages = []
    for i in range(layers):
        p = f"{prefix}.resblocks.{i}."
        xn = layer_norm(x, sd[p + "ln_1.weight"], sd[p + "ln_1.bias"])
        x = x + self_attention(xn, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"],
                               sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"],
                               key_padding_mask, heads)
        h = layer_norm(x, sd[p + "ln_2.weight"], sd[p + "ln_2.bias"])
        h = quick_gelu(h @ sd[p + "mlp.c_fc.weight"].t() + sd[p + "mlp.c_fc.bias"])
        x = x + (h @ sd[p + "mlp.c_proj.weight"].t() + sd[p + "mlp.c_proj.bias"])
        stages.append(xn)
    stages.pop(0)
    stages.append(x)
    return stages


def decoder_stack(x, memory, tgt_kpm, mem_kpm, sd, prefix: str, layers: int, heads: int):
    """TemporalDecoder.forward (model/tfm_model.py:96-103) over ResidualDecoderBlock_Step
    (:82-86): pre-LN self-attention, cross-attention on UN-normalised memory, MLP."""
    stages = []
    for i in range(layers):
        p = f"{prefix}.resblocks.{i}."
        xn = layer_norm(x, sd[p + "ln_1.weight"], sd[p + "ln_1.bias"])
        x = x + self_attention(xn, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"],
                               sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"],
                               tgt_kpm, heads)
        x = x + cross_attention(layer_norm(x, sd[p + "ln_2.weight"], sd[p + "ln_2.bias"]), memory,
                                sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"],
                                sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"],
                                mem_kpm, heads)
        h = layer_norm(x, sd[p + "ln_3.weight"], sd[p + "ln_3.bias"])
        h = quick_gelu(h @ sd[p + "mlp.c_fc.weight"].t() + sd[p + "mlp.c_fc.bias"])
        x = x + (h @ sd[p + "mlp.c_proj.weight"].t() + sd[p + "mlp.c_proj.bias"])
        stages.append(xn)
    stages.pop(0)
    stages.append(x)
    return stages


def interpolate_pos(table, k: int, size: int):
    """F.interpolate(table[None,:k].T, size, mode='linear', align_corners=False).T
    (model/tan_model.py:157-160): src = (i + 0.5) * k/size - 0.5 clamped at 0, lerp of neighbours."""
    scale = k / size
    idx = (torch.arange(size, dtype=torch.float32) + 0.5) * scale - 0.5
    idx = idx.clamp(min=0.0)
    i0 = idx.floor().long().clamp(max=k - 1)
    i1 = (i0 + 1).clamp(max=k - 1)
    lam = (idx - i0.float())[:, None]
    return table[i0] * (1 - lam) + table[i1] * lam


class TanOracle:
    """Functional restatement of `TemporalAligner` (model/tan_model.py:13-312)."""

    def __init__(self, state_dict, num_encoder_layers: int, num_decoder_layers: int,
                 use_text_pos_enc: int = 0, use_alignability_head: int = 0, heads: Optional[int] = None):
        self.sd = {k: _t(v) for k, v in state_dict.items()}
        self.E = num_encoder_layers
        self.D = num_decoder_layers
        self.width = self.sd["video_pre_proj.weight"].shape[0]
        self.heads = heads if heads is not None else self.width // 64
        self.use_text_pos_enc = use_text_pos_enc
        self.use_alignability_head = use_alignability_head

    def _ln(self, name, x):
        return layer_norm(x, self.sd[name + ".weight"], self.sd[name + ".bias"])

    def _video_pos(self, T, pos_start, interpolate_from, table="temporal_pos_embed"):
        tab = self.sd[table]
        if interpolate_from:
            return interpolate_pos(tab, int(interpolate_from), T)
        return tab[pos_start:pos_start + T]

    def _video_embed(self, video, pos_start, interpolate_from):
        """model/tan_model.py:155-167 (= :187-199)."""
        x = self._ln("ln_video_init", video @ self.sd["video_pre_proj.weight"].t())
        pos = self._video_pos(video.shape[1], pos_start, interpolate_from)
        return x + self._ln("ln_position_init", pos)[None]

    def get_visual_feature(self, video, video_padding_mask, interpolate_from=None, pos_start=0):
        """model/tan_model.py:152-179 -> [B, S, T, C]."""
        x = self._video_embed(_t(video), pos_start, interpolate_from)
        if self.E == 0:
            return x
        st = encoder_stack(x, video_padding_mask, self.sd, "video_temporal_encoder", self.E, self.heads)
        st[-1] = self._ln("ln_video_post_enc", st[-1])
        return torch.stack(st, dim=1)

    def get_textual_feature(self, text):
        """model/tan_model.py:231-234."""
        return self._ln("ln_text_init", _t(text) @ self.sd["text_pre_proj.weight"].t())

    def get_textual_feature_with_time(self, text, interpolate_from=None, pos_start=0):
        """model/tan_model.py:212-228."""
        t = self.get_textual_feature(text)
        pos = self._video_pos(t.shape[1], pos_start, interpolate_from, "text_temporal_pos_embed")
        return t + self._ln("ln_position_init", pos)[None]

    def get_joint_feature(self, video, video_padding_mask, text_feat, text_padding_mask,
                          interpolate_from=None, pos_start=0):
        """model/tan_model.py:182-209 -> ([B,S,T,C], [B,S,N,C])."""
        T = video.shape[1]
        x = torch.cat([self._video_embed(_t(video), pos_start, interpolate_from), text_feat], dim=1)
        kpm = torch.cat([video_padding_mask, text_padding_mask], dim=1)
        st = encoder_stack(x, kpm, self.sd, "joint_temporal_encoder", self.D, self.heads)
        st[-1] = self._ln("ln_joint_post_enc", st[-1])
        out = torch.stack(st, dim=1)
        return out[:, :, :T], out[:, :, T:]

    def forward(self, video, text, video_padding_mask, text_padding_mask, interpolate_from=None,
                pos_starts: Sequence[int] = (0, 0, 0)):
        """model/tan_model.py:100-149.  pos_starts = the three `np.random.randint` draws in call
        order (:163 video stack, :224 text-with-time, :195 joint stack)."""
        vpm = _t(video_padding_mask, torch.bool)
        tpm = _t(text_padding_mask, torch.bool)
        v = self.get_visual_feature(video, vpm, interpolate_from, pos_starts[0])
        t_raw = self.get_textual_feature(text)
        vn = v / v.norm(dim=-1, keepdim=True)
        tn = t_raw / t_raw.norm(dim=-1, keepdim=True)
        out = {"logits_dual": torch.einsum("astc,bkc->astbk", vn, tn),
               "dual_feature_video": vn, "dual_feature_text": tn}
        t_in = (self.get_textual_feature_with_time(text, interpolate_from, pos_starts[1])
                if self.use_text_pos_enc else t_raw)
        jv, jt = self.get_joint_feature(video, vpm, t_in, tpm, interpolate_from, pos_starts[2])
        jvn = jv / jv.norm(dim=-1, keepdim=True)
        jtn = jt / jt.norm(dim=-1, keepdim=True)
        out["logits_joint"] = torch.einsum("astc,bskc->astbk", jvn, jtn)
        if self.use_alignability_head:
            w, b = self.sd["binary_head.weight"], self.sd["binary_head.bias"]
            out["dual_logits_alignability"] = t_raw @ w.t() + b
            out["joint_logits_alignability"] = jt @ w.t() + b
        return out

    def get_text_visual_sim_dual(self, video, text, interpolate_from=None, pos_start=0):
        """model/tan_model.py:266-283 -> [B, S, T, N]."""
        B, T, _ = video.shape
        v = self.get_visual_feature(video, torch.zeros(B, T, dtype=torch.bool), interpolate_from, pos_start)
        t = self.get_textual_feature(text)
        vn = v / v.norm(dim=-1, keepdim=True)
        tn = t / t.norm(dim=-1, keepdim=True)
        return torch.einsum("bstc,bkc->bstk", vn, tn)

    def get_text_visual_sim_joint(self, video, text, interpolate_from=None, pos_starts=(0, 0)):
        """model/tan_model.py:237-263 -> [B, S, T, N]."""
        t_if = None
        if isinstance(interpolate_from, (list, tuple)):
            interpolate_from, t_if = interpolate_from
        t = (self.get_textual_feature_with_time(text, t_if, pos_starts[0]) if self.use_text_pos_enc
             else self.get_textual_feature(text))
        B, T, _ = video.shape
        N = t.shape[1]
        jv, jt = self.get_joint_feature(video, torch.zeros(B, T, dtype=torch.bool), t,
                                        torch.zeros(B, N, dtype=torch.bool), interpolate_from, pos_starts[1])
        jvn = jv / jv.norm(dim=-1, keepdim=True)
        jtn = jt / jt.norm(dim=-1, keepdim=True)
        return torch.einsum("bstc,bskc->bstk", jvn, jtn)


# ------------------------------------------------------------------------------------------------
# Loss (train/loss.py)
# ------------------------------------------------------------------------------------------------

def mask_from_time(start_list, end_list, T: int, N: int):
    """get_mask_from_time (train/loss.py:26-41): mask[b,n,t] = start[b,n] <= t < end[b,n];
    missing sentences are padded with start = T+100, end = -100 (all-False rows)."""
    B = len(start_list)
    start = torch.full((B, N), float(T) + 100.0)
    end = torch.full((B, N), -100.0)
    for b in range(B):
        nb = len(start_list[b])
        start[b, :nb] = torch.tensor(start_list[b], dtype=torch.float32)
        end[b, :nb] = torch.tensor(end_list[b], dtype=torch.float32)
    steps = torch.arange(T, dtype=torch.float32)[None, None, :]
    mask = (start[:, :, None] <= steps) & (steps < end[:, :, None])
    return mask, start, end


def milnce_terms(z, tgt, col_valid):
    """Closed form of train/loss.py:241-256 with masks instead of boolean-index compaction.

    z: [S, R, C] scaled logits (R = B*T rows, C = B*N columns incl. padded ones);
    tgt: [R, C] bool positives (False on padded columns); col_valid: [C] bool.
    Returns (v_loss [S, R], row_has_pos [R], t_loss [S, C], col_has_pos [C]).
    The reference's `-6e4` fill (:245) contributes exp(-6e4 - max) == 0 in fp32 whenever the row /
    column has a positive, which are the only ones kept (:250,:254), so -inf is equivalent."""
    ninf = float("-inf")
    z_all = z.masked_fill(~col_valid[None, None, :], ninf)
    z_pos = z.masked_fill(~tgt[None], ninf)
    row_has = tgt.any(dim=1)
    col_has = tgt.any(dim=0)
    v = torch.logsumexp(z_all, dim=2) - torch.logsumexp(z_pos, dim=2)
    t = torch.logsumexp(z_all, dim=1) - torch.logsumexp(z_pos, dim=1)
    return v, row_has, t, col_has


def nce_loss(logits, tgt, col_valid):
    """loss_x of train/loss.py:256 / :274 for one model.  logits: [B, S, T, B, N] (unscaled)."""
    B, S, T = logits.shape[:3]
    z = (logits / TEMPERATURE).permute(1, 0, 2, 3, 4).reshape(S, B * T, -1)
    v, rh, t, ch = milnce_terms(z, tgt, col_valid)
    return (v[:, rh].mean() + t[:, ch].mean()) / 2


def get_loss_init(logits_dual, logits_joint, start_list, end_list, text_padding_mask):
    """get_loss (train/loss.py:55-86, :231-275, :359-373) for `--model init`, loss_threshold=0,
    no agreement labelling, no alignability head.  Returns dict(loss, loss-dual, loss-joint)."""
    B, S, T, _, N = logits_dual.shape
    tpm = _t(text_padding_mask, torch.bool)
    mask, _, _ = mask_from_time(start_list, end_list, T, N)          # [B, N, T]
    tgt = torch.zeros(B, T, B, N, dtype=torch.bool)
    for b in range(B):
        tgt[b, :, b, :] = mask[b].t()
    col_valid = (~tpm).reshape(-1)
    tgt = tgt.reshape(B * T, B * N) & col_valid[None]
    ld = nce_loss(_t(logits_dual), tgt, col_valid)
    lj = nce_loss(_t(logits_joint), tgt, col_valid)
    return {"loss": (ld + lj) / 2, "loss-dual": ld.detach(), "loss-joint": lj.detach()}
