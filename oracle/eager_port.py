"""Torch-eager GPU port of the reference's hot path -- the "existing GPU implementation" bar.

TEST / BENCH INFRASTRUCTURE (never imported by the product).  The reference is Python on torch: on a GPU its path is
`nn.MultiheadAttention` / `nn.LayerNorm` / `nn.Linear` / `torch.einsum` / boolean indexing / `torch.logsumexp` under
`torch.cuda.amp.autocast()` (train/main.py:81).  /root/reference does not exist on the GPU box, so this module
re-states those SAME torch library calls, module for module, with the reference's parameter names (it loads the same
state dict), so that bench.py can time them on the same B200:
  * `EagerBlock` / `EagerEncoder`        model/tfm_model.py:17-55 (incl. the seq-first layout and the stage list)
  * `EagerTAN.forward`                   model/tan_model.py:100-149 (pre-projection computed twice, as the reference)
  * `eager_get_loss_init`                train/loss.py:55-86,:231-275,:359-373 (`--model init`, no flags), operation
                                         for operation: `/ 0.07`, the materialised [B,T,B,N] target, boolean-index
                                         compaction, clone + `-6e4` fill, four logsumexp per model
Checked against the CPU oracle in tests/test_oracle.py (CPU, fp32).
"""
from __future__ import annotations

from collections import OrderedDict

import torch
from torch import nn


class QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class EagerBlock(nn.Module):
    """model/tfm_model.py:17-38."""

    def __init__(self, d_model: int, n_head: int):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = nn.LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d_model, d_model * 4)), ("gelu", QuickGELU()),
                                              ("c_proj", nn.Linear(d_model * 4, d_model))]))
        self.ln_2 = nn.LayerNorm(d_model)

    def forward(self, x, key_padding_mask):
        xn = self.ln_1(x)
        x = x + self.attn(xn, xn, xn, need_weights=False, key_padding_mask=key_padding_mask)[0]
        x = x + self.mlp(self.ln_2(x))
        return x, xn


class EagerEncoder(nn.Module):
    """model/tfm_model.py:41-55."""

    def __init__(self, width: int, layers: int, heads: int):
        super().__init__()
        self.resblocks = nn.ModuleList([EagerBlock(width, heads) for _ in range(layers)])

    def forward(self, x, key_padding_mask):
        outs = []
        for blk in self.resblocks:
            x, xn = blk(x, key_padding_mask)
            outs.append(xn)
        outs.pop(0)
        outs.append(x)
        return outs


class EagerTAN(nn.Module):
    """model/tan_model.py:13-149 (`random_pos_start=0`, learned positions, no text positions, no head)."""

    def __init__(self, E: int, D: int, width: int = 512, video_dim: int = 1024, text_dim: int = 512):
        super().__init__()
        h = width // 64
        self.video_temporal_encoder = EagerEncoder(width, E, h)
        self.joint_temporal_encoder = EagerEncoder(width, D, h)
        self.video_pre_proj = nn.Linear(video_dim, width, bias=False)
        self.text_pre_proj = nn.Linear(text_dim, width, bias=False)
        for n in ("ln_text_init", "ln_video_init", "ln_position_init", "ln_video_post_enc", "ln_joint_post_enc"):
            setattr(self, n, nn.LayerNorm(width))
        self.temporal_pos_embed = nn.Parameter(torch.zeros(1024, width))
        self.text_temporal_pos_embed = nn.Parameter(torch.zeros(1024, width))
        self.mlp = nn.Linear(width, width)

    def _video_in(self, video):
        T = video.shape[1]
        x = self.ln_video_init(self.video_pre_proj(video))
        return x + self.ln_position_init(self.temporal_pos_embed[None, 0:T, :])

    def forward(self, video, text, video_padding_mask, text_padding_mask):
        T = video.shape[1]
        # dual (model/tan_model.py:152-179, :231-234)
        st = self.video_temporal_encoder(self._video_in(video).permute(1, 0, 2), video_padding_mask)
        st[-1] = self.ln_video_post_enc(st[-1])
        v = torch.stack(st, dim=1).permute(2, 1, 0, 3)
        t = self.ln_text_init(self.text_pre_proj(text))
        vn = v / v.norm(dim=-1, keepdim=True)
        tn = t / t.norm(dim=-1, keepdim=True)
        logits_dual = torch.einsum("astc,bkc->astbk", vn, tn)
        # joint (model/tan_model.py:182-209): the pre-projection is computed a second time, as the reference does
        x = torch.cat((self._video_in(video), t), dim=1).permute(1, 0, 2)
        kpm = torch.cat((video_padding_mask, text_padding_mask), dim=-1)
        st = self.joint_temporal_encoder(x, kpm)
        st[-1] = self.ln_joint_post_enc(st[-1])
        j = torch.stack(st, dim=1).permute(2, 1, 0, 3)
        jv, jt = j[:, :, :T], j[:, :, T:]
        jvn = jv / jv.norm(dim=-1, keepdim=True)
        jtn = jt / jt.norm(dim=-1, keepdim=True)
        logits_joint = torch.einsum("astc,bskc->astbk", jvn, jtn)
        return {"logits_dual": logits_dual, "logits_joint": logits_joint}


def eager_targets(start_list, end_list, T: int, N: int, device):
    """train/loss.py:26-41 + :80-85: the materialised [B, T, B, N] float target."""
    from torch.nn.utils.rnn import pad_sequence
    B = len(start_list)
    start = pad_sequence([torch.as_tensor(i, dtype=torch.float32) for i in start_list], batch_first=True,
                         padding_value=T + 1e2).to(device, non_blocking=True)
    end = pad_sequence([torch.as_tensor(i, dtype=torch.float32) for i in end_list], batch_first=True,
                       padding_value=-1e2).to(device, non_blocking=True)
    if start.shape[1] < N:
        start = torch.nn.functional.pad(start, (0, N - start.shape[1]), value=T + 1e2)
        end = torch.nn.functional.pad(end, (0, N - end.shape[1]), value=-1e2)
    steps = torch.arange(T, device=device)[None, None, :]
    tgt_raw = (start[:, :, None] <= steps) & (steps < end[:, :, None])                    # [B, N, T]
    return tgt_raw.permute(0, 2, 1).unsqueeze(2).repeat(1, 1, B, 1) * torch.eye(B, device=device)[:, None, :, None]


def eager_get_loss_init(logits, start_list, end_list, text_padding_mask, T: int, N: int):
    """train/loss.py:55-86,:231-275,:359-373 for `--model init` without flags, the reference's own operations."""
    ld, lj = logits["logits_dual"] / 0.07, logits["logits_joint"] / 0.07
    B = ld.shape[0]
    keep = ~text_padding_mask.bool()
    tgt = eager_targets(start_list, end_list, T, N, ld.device)[:, :, keep].view(B * T, -1)
    v_has, t_has = tgt.sum(-1) > 0, tgt.sum(-2) > 0
    losses = []
    for z in (ld, lj):
        S = z.shape[1]
        z = z[:, :, :, keep].permute(1, 0, 2, 3).reshape(S, B * T, -1)
        zp = z.clone()
        zp[:, ~tgt.bool()] = -6e4
        v = (torch.logsumexp(z, dim=-1) - torch.logsumexp(zp, dim=-1))[:, v_has]
        t = (torch.logsumexp(z, dim=-2) - torch.logsumexp(zp, dim=-2))[:, t_has]
        losses.append((v.mean() + t.mean()) / 2)
    return {"loss": (losses[0] + losses[1]) / 2, "loss-dual": losses[0].detach(), "loss-joint": losses[1].detach()}
