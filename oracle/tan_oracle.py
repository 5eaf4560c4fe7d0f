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


    def get_alignability(self, video, text, interpolate_from=None, pos_starts=(0, 0)):
        """model/tan_model.py:284-312 -> {'alignability-dual' [B,N,1], 'alignability-joint' [B,S,N,1]}: binary_head
        on the raw text features and on every stage of the joint stack's text part (zero padding masks)."""
        t_if = None
        if isinstance(interpolate_from, (list, tuple)):
            interpolate_from, t_if = interpolate_from
        t_raw = self.get_textual_feature(text)
        t = (self.get_textual_feature_with_time(text, t_if, pos_starts[0]) if self.use_text_pos_enc else t_raw)
        B, T, _ = video.shape
        N = t.shape[1]
        _, jt = self.get_joint_feature(video, torch.zeros(B, T, dtype=torch.bool), t,
                                       torch.zeros(B, N, dtype=torch.bool), interpolate_from, pos_starts[1])
        w, b = self.sd["binary_head.weight"], self.sd["binary_head.bias"]
        return {"alignability-dual": t_raw @ w.t() + b, "alignability-joint": jt @ w.t() + b}


# ------------------------------------------------------------------------------------------------
# Text embedder (model/word2vec_model.py:76-102)
# ------------------------------------------------------------------------------------------------

def word2vec_forward(sd, input_ids, attention_mask=None):
    """Word2VecModel.forward: Embedding lookup -> fc1 + ReLU -> max over the words with ignored words at -6e4 (a
    sentence without any kept word keeps all of them, :93) -> fc2.  Returns (pooler_output [S, 512],
    last_hidden_state [S, W, 512])."""
    x = _t(sd["word_embd.weight"])[torch.as_tensor(input_ids).long()]
    x = torch.relu(x @ _t(sd["fc1.weight"]).t() + _t(sd["fc1.bias"]))
    if attention_mask is not None:
        keep = torch.as_tensor(attention_mask).bool().clone()
        keep[keep.sum(-1) == 0, :] = True
        pooled = x.masked_fill(~keep[:, :, None], -6e4).max(dim=1).values
    else:
        pooled = x.max(dim=1).values
    w2, b2 = _t(sd["fc2.weight"]), _t(sd["fc2.bias"])
    return pooled @ w2.t() + b2, x @ w2.t() + b2


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


# ------------------------------------------------------------------------------------------------
# Loss, remaining branches: agreement self-labelling (L4), loss threshold + alignability BCE (L5)
# ------------------------------------------------------------------------------------------------

def quantile_linear(x, q: float):
    """torch.quantile(x, q) (default 'linear' interpolation) on a 1-D tensor, restated:
    sort, position q*(n-1), linear blend of the two neighbours."""
    xs, _ = torch.sort(x.float().reshape(-1))
    n = xs.numel()
    pos = q * (n - 1)
    lo = int(math.floor(pos))
    hi = min(lo + 1, n - 1)
    return xs[lo] + (xs[hi] - xs[lo]) * (pos - lo)


def own_clip_block(z, video_padding_mask=None, text_padding_mask=None):
    """z [B,S,T,B,N] -> its own-clip blocks [B,S,T,N] (torch.diagonal(dim1=0, dim2=3), train/loss.py:92-95)
    with -6e4 on padded frames / padded sentences (:96-100)."""
    B = z.shape[0]
    blk = torch.stack([z[b, :, :, b, :] for b in range(B)])                   # [B,S,T,N]
    if video_padding_mask is not None:
        blk = blk.masked_fill(video_padding_mask[:, None, :, None], -6e4)
    if text_padding_mask is not None:
        blk = blk.masked_fill(text_padding_mask[:, None, None, :], -6e4)
    return blk


def best_window(q_bn, z_bn, dur: int):
    """One sentence of the self-labelling scan (train/loss.py:111-145), without the [T,T] circulant.
    q_bn [T]: the sentence's probability over time; z_bn [T]: its last-stage logits; dur: window length.
    Window i covers frames [i, i+dur) for 0 <= i <= T-dur, never frames 0 and T-1 (:127-128); its score is
    the MEAN of q over the frames it keeps.  Returns (score, window mask [T] bool, mean logit in the window);
    the first best window wins ties (torch.max)."""
    T = q_bn.shape[0]
    best, best_i = torch.tensor(0.0), 0          # invalid windows score 0 and lose to any real one
    scores = torch.zeros(T)
    if dur >= 1:
        for i in range(0, T - dur + 1):
            lo, hi = max(i, 1), min(i + dur, T - 1)
            if hi > lo:
                scores[i] = (q_bn[lo:hi] * (1.0 / float(hi - lo))).sum()    # the reference's operation order (:133-134)
    best, best_i = scores.max(0)
    best_i = int(best_i)
    win = torch.zeros(T, dtype=torch.bool)
    mean_logit = torch.tensor(0.0)
    if dur >= 1 and best_i <= T - dur:
        lo, hi = max(best_i, 1), min(best_i + dur, T - 1)
        if hi > lo:
            win[lo:hi] = True
            mean_logit = (z_bn[lo:hi] * (1.0 / float(hi - lo))).sum()
    return best, win, mean_logit


def self_label(z_own_last, mask_bnt, text_padding_mask):
    """Self-labelled window per sentence from the last-stage own-clip logits z_own_last [B,T,N] (already
    scaled by 1/0.07 and padded with -6e4): p = softmax over sentences, q = softmax over time of p/0.07
    (:103), window length = the sentence's original duration (>= 1; 0 for padded sentences, :113-115).
    Returns (window [B,N,T] bool, max_logits [B,N])."""
    B, T, N = z_own_last.shape
    p = torch.softmax(z_own_last, dim=-1)
    q = torch.softmax(p / TEMPERATURE, dim=-2)
    dur = mask_bnt.sum(-1).clamp(min=1)
    dur = dur.masked_fill(text_padding_mask, 0)
    win = torch.zeros(B, N, T, dtype=torch.bool)
    mx = torch.zeros(B, N)
    for b in range(B):
        for n in range(N):
            _, w, ml = best_window(q[b, :, n], z_own_last[b, :, n], int(dur[b, n]))
            win[b, n] = w
            mx[b, n] = ml
    return win, mx


def agreement_targets(zd_own_last, zj_own_last, mask_bnt, text_padding_mask, kind: str):
    """train/loss.py:88-229: new per-clip targets [B,N,T] bool from the agreement of the dual and joint
    self-labels.  Returns (targets, confidence_ratio)."""
    B, T, N = zd_own_last.shape
    jw, jmax = self_label(zj_own_last, mask_bnt, text_padding_mask)
    dw, dmax = self_label(zd_own_last, mask_bnt, text_padding_mask)
    inter = (jw & dw).sum(-1).float()
    union = (jw | dw).sum(-1).float()
    iou = inter / union.clamp(min=1e-5)
    valid = ~text_padding_mask
    conf_text = (dmax >= quantile_linear(dmax[valid], 0.3)) & (jmax >= quantile_linear(jmax[valid], 0.3))
    conf_iou = iou >= 0.5
    conf = conf_text & conf_iou
    if kind == "i":
        tgt = (jw & dw) & conf[:, :, None]
    elif kind == "u":
        tgt = (jw | dw) & conf[:, :, None]
    elif kind == "keep":
        tgt = torch.where(conf_iou[:, :, None], jw | dw, mask_bnt)
    elif kind == "keep-joint":
        tgt = torch.where(conf_iou[:, :, None], jw, mask_bnt)
    else:
        raise ValueError(kind)
    # exclusion (:217-226): per frame only the first sentence keeps its 1 (sentence 0 always keeps its own
    # column); sentences left with nothing get their original timestamps back
    out = torch.zeros_like(tgt)
    for b in range(B):
        for t in range(T):
            hits = torch.nonzero(tgt[b, :, t]).reshape(-1)
            if hits.numel() > 0 and int(hits[0]) > 0:
                out[b, int(hits[0]), t] = True
        out[b, 0, :] = tgt[b, 0, :]
        for n in range(N):
            if not out[b, n].any():
                out[b, n] = mask_bnt[b, n]
    return out, conf[valid].float().mean()


def get_loss_full(logits: dict, start_list, end_list, video_padding_mask, text_padding_mask, args,
                  abs_text_pos=None):
    """get_loss (train/loss.py:55-373) with every branch: agreement self-labelling (learn_agreement,
    temporal_agreement_type), loss_threshold, alignability-head BCE, model in {init, cotrain}.
    `logits`: the forward dict (+ 'ema-logits_*' for cotrain).  Returns the reference's loss_dict keys.

    Two properties of the reference that this restatement keeps: (1) in `init` mode the -6e4 padding fill
    of the self-labelling step is applied IN PLACE to the own-clip blocks of the scaled logits the loss is
    then computed from (torch.diagonal returns a view, :92-100), so padded frames drop out of their own
    clip's columns; (2) the thresholded text loss indexes the per-text losses with a mask over ALL real
    sentences (:298), i.e. it assumes every real sentence has a positive."""
    tpm = _t(text_padding_mask, torch.bool)
    vpm = _t(video_padding_mask, torch.bool)
    zd = _t(logits["logits_dual"]) / TEMPERATURE
    zj = _t(logits["logits_joint"]) / TEMPERATURE
    B, Sd, T, _, N = zd.shape
    mask, _, _ = mask_from_time(start_list, end_list, T, N)                # [B,N,T]
    valid = ~tpm
    col_valid = valid.reshape(-1)
    out = {}
    tgt_bnt = mask
    if getattr(args, "learn_agreement", 0):
        if args.model == "cotrain":
            src_d = _t(logits["ema-logits_dual"]) / TEMPERATURE
            src_j = _t(logits["ema-logits_joint"]) / TEMPERATURE
        else:
            src_d, src_j = zd, zj
        own_j = own_clip_block(src_j, vpm, tpm)
        own_d = own_clip_block(src_d, vpm, tpm)
        if args.model != "cotrain":                                       # property (1)
            for b in range(B):
                zj[b, :, :, b, :] = own_j[b]
                zd[b, :, :, b, :] = own_d[b]
        tgt_bnt, ratio = agreement_targets(own_d[:, -1], own_j[:, -1], mask, tpm, args.temporal_agreement_type)
        out["confidence-ratio"] = ratio
        out["iou-threshold"] = torch.tensor(0.5)
    tgt = torch.zeros(B, T, B, N, dtype=torch.bool)
    for b in range(B):
        tgt[b, :, b, :] = tgt_bnt[b].t()
    tgt = tgt.reshape(B * T, B * N) & col_valid[None]

    def terms(z):
        S = z.shape[1]
        return milnce_terms(z.permute(1, 0, 2, 3, 4).reshape(S, B * T, -1), tgt, col_valid)

    vd, rh, td, ch = terms(zd)
    vj, _, tj, _ = terms(zj)
    loss_dual = (vd[:, rh].mean() + td[:, ch].mean()) / 2
    loss_joint = (vj[:, rh].mean() + tj[:, ch].mean()) / 2
    out["loss-dual"], out["loss-joint"] = loss_dual, loss_joint
    thr = float(getattr(args, "loss_threshold", 0.0))
    head = bool(getattr(args, "use_alignability_head", 0))
    if thr > 0 or head:
        md = own_clip_block(zd)[:, -1].max(dim=1).values[valid]           # [M] best frame per real sentence
        mj = own_clip_block(zj)[:, -1].max(dim=1).values[valid]
        comb = (md - md.mean()) / md.std() + (mj - mj.mean()) / mj.std()
        metric = -comb
        keep = metric <= quantile_linear(metric, thr)                     # [M]
        keep_c = torch.zeros(B * N, dtype=torch.bool)
        keep_c[col_valid] = keep
        rows_th = (tgt & keep_c[None]).any(dim=1)
        if thr > 0:
            out["loss-dual-all"], out["loss-joint-all"] = loss_dual, loss_joint
            cols_th = ch & keep_c
            loss_dual_th = (vd[:, rows_th].mean() + td[:, cols_th].mean()) / 2
            loss_joint_th = (vj[:, rows_th].mean() + tj[:, cols_th].mean()) / 2
            out["loss-dual"], out["loss-joint"] = loss_dual_th, loss_joint_th
        if head:
            label = torch.full_like(md, 2.0)
            qd, qj = quantile_linear(md, 0.5), quantile_linear(mj, 0.5)
            label[(md > qd) & (mj > qj)] = 1.0
            label[(md < qd) & (mj < qj)] = 0.0
            if abs_text_pos is not None:
                centre = _t(abs_text_pos)[valid].mean(-1)
                label[(centre < 0.2) | (centre > 0.8)] = 0.0
            has = ch[col_valid]                                            # real sentences with a positive
            xj = _t(logits["joint_logits_alignability"])[:, 2, :, 0][valid][has]
            sel = label != 2.0
            y = label[sel]
            pw = 1.0 / y.mean() - 1.0
            x = xj[sel]
            sp = lambda v: torch.clamp(v, min=0) + torch.log1p(torch.exp(-v.abs()))     # softplus
            bce = (pw * y * sp(-x) + (1 - y) * sp(x)).mean()
            out["loss-joint-bce"] = bce
            out["alignability_top1"] = ((x > 0) == (y > 0.5)).float().mean()
    nce_w = 0.0 if getattr(args, "optim_policy", "default") == "bce" else 1.0
    if thr > 0:
        out["loss-total"] = (loss_dual + loss_joint) / 2
        loss = (loss_dual_th + loss_joint_th) / 2
    else:
        loss = (loss_dual + loss_joint) / 2
    if head:
        loss = loss * nce_w + bce
    out["loss"] = loss
    return out


# ------------------------------------------------------------------------------------------------
# Sliding-window alignment, 'overlap-seq' (eval/eval_zeroshot_align.py:126-205)
# ------------------------------------------------------------------------------------------------

def overlap_seq_windows(vlen: int, seq_len: int, text_mid_ts, tgt_aligned):
    """The window loop of eval/eval_zeroshot_align.py:129-177, host part only: for every step the frame range
    [step, min(vlen, step + seq_len)) and the boolean sentence mask the reference builds (:149-167), skipping the
    steps it skips (:155-156, :176-177).  text_mid_ts = (start + end) / 2 per sentence (:134); tgt_aligned marks
    the alignable sentences, the NON-alignable ones choose the windows (:149-154)."""
    mid = np.asarray(text_mid_ts, dtype=np.float64)
    aligned = np.asarray(tgt_aligned).astype(bool)
    n_text = len(mid)
    step = np.arange(0, vlen - seq_len // 2, seq_len // 4)                            # :129
    out = []
    for idx, step_ in enumerate(step):
        na_idx = np.arange(n_text)[~aligned]                                          # :149
        na_mid = mid[~aligned]                                                        # :150
        in_win = np.logical_and(step_ - seq_len <= na_mid, na_mid <= step_ + seq_len + seq_len)   # :151-153
        active = na_idx[in_win]
        if len(active) == 0:                                                          # :155
            continue
        left, right = active.min(), active.max()                                      # :158-160
        mask = np.zeros(n_text).astype(bool)
        if idx <= 3:                                                                  # :163
            left = 0
        elif idx >= len(step) - 4:                                                    # :165
            right = vlen
        mask[left: right + 1] = True                                                  # :167
        if np.sum(mask) == 0:                                                         # :176
            continue
        out.append((int(step_), int(min(vlen, step_ + seq_len)), mask))
    return out


def overlap_seq_alignment(sim_fn, vlen: int, n_text: int, windows, use_alignability_head: bool):
    """The accumulation of eval/eval_zeroshot_align.py:136-205.  sim_fn(t0, t1, mask) plays get_text_visual_sim
    (train/main.py:171-189) for one window: {'sim', 'dual-sim'} [1, S, n_active, t1 - t0] (already / 0.07) and,
    with the head, {'alignability-dual' [1, n_active, 1], 'alignability-joint' [1, S, n_active, 1]}."""
    eps = torch.tensor(1e-5)
    logits, logits_dual = torch.zeros(n_text, vlen), torch.zeros(n_text, vlen)
    overlap = torch.zeros(n_text, vlen)
    a_dual, a_joint, text_overlap = torch.zeros(n_text), torch.zeros(n_text), torch.zeros(n_text)
    for t0, t1, mask in windows:
        m = torch.from_numpy(mask)
        o = sim_fn(t0, t1, mask)
        if use_alignability_head:                                                     # :182-187
            a_dual[m] += o["alignability-dual"][0, :, 0]
            a_joint[m] += o["alignability-joint"][0, 2, :, 0]
        else:                                                                         # :188-195
            a_dual[m] += o["dual-sim"][0, -1].max(-1).values
            a_joint[m] += o["sim"][0, -1].max(-1).values
        text_overlap[m] += 1
        logits[m, t0:t1] += o["sim"][0, -1, :]                                        # :197-199
        logits_dual[m, t0:t1] += o["dual-sim"][0, -1, :]
        overlap[m, t0:t1] += 1
    logits = logits.div(torch.maximum(overlap, eps))                                  # :200-201
    logits_dual = logits_dual.div(torch.maximum(overlap, eps))
    return {"sim-joint": logits, "sim-dual": logits_dual, "sim": (logits + logits_dual) / 2, "overlap": overlap,
            "alignability-dual": a_dual.div(torch.maximum(text_overlap, eps)),        # :203-204
            "alignability-joint": a_joint.div(torch.maximum(text_overlap, eps))}


def global_alignment(sim_fn, use_alignability_head: bool):
    """The 'global' method, eval/eval_zeroshot_align.py:207-215.  sim_fn() plays get_text_visual_sim(video, text_str,
    interpolate_from=seq_len) for the whole video (same dict as in overlap_seq_alignment)."""
    o = sim_fn()
    sim = o["sim"][0, -1, :]                                                          # :209
    if use_alignability_head:                                                         # :210-212
        a_dual, a_joint = o["alignability-dual"][0, :, 0], o["alignability-joint"][0, -1, :, 0]
    else:                                                                             # :213-215
        a_dual, a_joint = o["dual-sim"][0, -1].max(-1).values, o["sim"][0, -1].max(-1).values
    return {"sim": sim, "sim-joint": sim, "sim-dual": o["dual-sim"][0, -1, :], "alignability-dual": a_dual,
            "alignability-joint": a_joint}


def htm_align_metrics(videos, use_alignability_head: bool):
    """The per-video bookkeeping and the final metrics of eval/eval_zeroshot_align.py:217-249.  videos: iterable of
    (result dict with 'sim' [n_text, vlen] and 'alignability-joint' [n_text], tgt_aligned, start, end)."""
    from sklearn import metrics                                                       # eval_zeroshot_align.py:247
    recall, total_sim, total_tgt = [], [], []
    for res, tgt_aligned, start, end in videos:
        sim = res["sim"].clone()
        keep = torch.as_tensor(np.asarray(tgt_aligned)).bool()
        start_al, end_al = np.asarray(start)[keep.numpy()], np.asarray(end)[keep.numpy()]     # :119-120
        sim.masked_fill_(sim == 0, -6e4)                                              # :220
        prob = sim.softmax(-1)
        total_tgt.append(np.array(tgt_aligned))
        if use_alignability_head:                                                     # :217-218, :224-226
            total_sim.append(res["alignability-joint"].numpy())
        else:
            total_sim.append(sim.max(-1)[0].numpy())
        prob = prob[keep, :]                                                          # :229
        for i in range(prob.size(0)):                                                 # :231-234
            s, e = math.floor(start_al[i]), math.ceil(end_al[i])
            recall.append(s <= prob[i].argmax(-1).item() <= e)
    total_sim, total_tgt = np.concatenate(total_sim, 0), np.concatenate(total_tgt, 0)
    return {"Recall": np.mean(recall), "AUC": metrics.roc_auc_score(total_tgt, total_sim)}
