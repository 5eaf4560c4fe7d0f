"""Sliding-window text-to-video alignment of one long video: the 'overlap-seq' method of
`eval/eval_zeroshot_align.py:125-205` -- overlapped temporal windows, per-window text subsets, stitch by
averaging -- with ALL windows of the video batched into one forward of the dual and the joint stack.

The reference runs every window as its own batch-1 call and, inside it, the joint model and the dual model
as two more passes (`train/main.py:171-189`).  Here the windows become the clips of one batch: a short last
window is padded with masked frames and every window's sentences are padded to the longest subset with masked
sentences (masked keys never reach the softmax, so each window's result equals its stand-alone, unmasked
computation); the per-window similarities are the own-clip blocks (tan_own_clip_sim) of the last stage, stitched by
tan_align_stitch (overlap average, one launch per batch of windows) and decided by tan_align_argmax.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from ._lib import TanError
from .tan_model import LazyLogits, TemporalAligner

Window = Tuple[int, int, int, int]          # frames [t0, t1), sentences [n0, n1)


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise TanError(f"{what} runs on a CUDA (sm_100a) device only; there is no CPU path")


def plan_windows(vlen: int, seq_len: int, text_mid_ts: Sequence[float], anchor_mask: Sequence[bool]) -> List[Window]:
    """The window plan of eval/eval_zeroshot_align.py:129-177: windows of `seq_len` frames every seq_len/4
    frames; a window's sentences are the index range spanned by the ANCHOR sentences (the reference uses the
    non-alignable ones, whose ASR timestamps do not leak ground truth) whose mid timestamp lies within
    [step - seq_len, step + 2*seq_len]; the first / last four windows extend to the first / last sentence."""
    mid = np.asarray(text_mid_ts, dtype=np.float64)
    anchors = np.arange(len(mid))[np.asarray(anchor_mask, dtype=bool)]
    steps = np.arange(0, vlen - seq_len // 2, max(seq_len // 4, 1))
    out: List[Window] = []
    for idx, step in enumerate(steps):
        act = anchors[(step - seq_len <= mid[anchors]) & (mid[anchors] <= step + 2 * seq_len)]
        if len(act) == 0:
            continue
        left, right = int(act.min()), int(act.max())
        if idx <= 3:
            left = 0
        elif idx >= len(steps) - 4:
            right = min(vlen, len(mid) - 1)            # the reference's slice [left : vlen + 1] over the sentences
        if right < left:                               # an empty sentence slice: the reference skips the window (:176)
            continue
        out.append((int(step), int(min(vlen, step + seq_len)), left, right + 1))
    return out


@torch.no_grad()
def sliding_window_alignment(model: TemporalAligner, video: torch.Tensor, text_embed: torch.Tensor,
                             windows: Sequence[Window], max_windows_per_batch: int = 256) -> dict:
    """video [vlen, D_in] fp32 (CUDA), text_embed [n_text, D_text] fp32 (CUDA), windows from `plan_windows`.

    Returns fp32 tensors: 'sim-joint' / 'sim-dual' [n_text, vlen] = last-stage logits / 0.07 averaged over the
    windows that cover (sentence, frame) (eval_zeroshot_align.py:197-201), 'sim' = their mean (:205),
    'overlap' [n_text, vlen] coverage counts, and 'alignability-dual' / 'alignability-joint' [n_text] averaged over
    the windows that hold the sentence (:203-204): the alignability head's outputs (joint stage index 2, :182-187)
    or, for a model without the head, the sentence's largest similarity inside each window (:188-195)."""
    _require_cuda(video, "sliding_window_alignment")
    if model.random_pos_start:
        raise TanError("sliding_window_alignment needs random_pos_start=0 (deterministic positional offsets)")
    dev = video.device
    vlen, n_text = video.shape[0], text_embed.shape[0]
    head = bool(model.use_alignability_head)
    sim_j = torch.empty(n_text, vlen, dtype=torch.float32, device=dev)
    sim_d = torch.empty_like(sim_j)
    cover = torch.empty_like(sim_j)
    a_d = torch.zeros(n_text, dtype=torch.float32, device=dev)
    a_j = torch.zeros_like(a_d)
    a_n = torch.zeros_like(a_d)
    windows = list(windows)
    if not windows:
        for t_ in (sim_j, sim_d, cover):
            t_.zero_()
    for w0 in range(0, len(windows), max_windows_per_batch):
        chunk = windows[w0:w0 + max_windows_per_batch]
        W = len(chunk)
        win = np.asarray(chunk, dtype=np.int64).reshape(W, 4)
        T = int((win[:, 1] - win[:, 0]).max())
        N = int((win[:, 3] - win[:, 2]).max())
        # the windows become the clips of one batch: frame / sentence indices and padding masks built on the host in
        # one shot, gathered on the device by two index_select calls (no per-window copies)
        fi = win[:, 0:1] + np.arange(T)[None, :]
        ni = win[:, 2:3] + np.arange(N)[None, :]
        vpad, tpad = fi >= win[:, 1:2], ni >= win[:, 3:4]
        fi_d = torch.from_numpy(np.where(vpad, 0, fi).reshape(-1)).to(dev, non_blocking=True)
        ni_d = torch.from_numpy(np.where(tpad, 0, ni).reshape(-1)).to(dev, non_blocking=True)
        vpm = torch.from_numpy(vpad).to(dev, non_blocking=True)
        tpm = torch.from_numpy(tpad).to(dev, non_blocking=True)
        vb = video.index_select(0, fi_d).view(W, T, -1).masked_fill(vpm[..., None], 0.0)
        tb = text_embed.index_select(0, ni_d).view(W, N, -1).masked_fill(tpm[..., None], 0.0)
        out = model._forward_impl(vb.float(), tb.float(), vpm, tpm)
        blocks = {}
        for key in ("logits_dual", "logits_joint"):
            lg: LazyLogits = out[key]
            Bv, S, Tv, d = lg.vfeat.shape
            blocks[key] = ops.own_clip_sim(lg.vfeat, lg.tfeat, lg.shared_text, Bv, S, Tv, N, d, s_first=S - 1,
                                           s_count=1).view(W, Tv, N)                   # [W, T, N] cosines
        # stitching (eval_zeroshot_align.py:197-201): one kernel per batch of windows, sums in window order
        win_d = torch.from_numpy(win.astype(np.int32)).to(dev, non_blocking=True)
        last = w0 + max_windows_per_batch >= len(windows)
        ops.align_stitch(blocks["logits_joint"], blocks["logits_dual"], win_d, sim_j, sim_d, cover, accumulate=w0 > 0,
                         finalize=last)
        if head:
            # sentence n of window i -> row n0_i + n of the per-sentence accumulators (padded sentences -> weight 0)
            keep = torch.from_numpy((~tpad).reshape(-1).astype(np.float32)).to(dev, non_blocking=True)
            a_d.index_add_(0, ni_d, out["dual_logits_alignability"][:, :N, 0].reshape(-1).float() * keep)
            a_j.index_add_(0, ni_d, out["joint_logits_alignability"][:, 2, :N, 0].reshape(-1).float() * keep)
            a_n.index_add_(0, ni_d, keep)
        else:
            # no alignability head (eval_zeroshot_align.py:188-195): a sentence's score in a window is the largest
            # similarity (logit / 0.07) over the window's real frames
            keep = torch.from_numpy((~tpad).reshape(-1)).to(dev, non_blocking=True)
            for key, acc in (("logits_dual", a_d), ("logits_joint", a_j)):
                top = blocks[key].float().masked_fill(vpm[..., None], float("-inf")).amax(dim=1) * (1.0 / 0.07)
                acc.index_add_(0, ni_d, torch.where(keep, top.reshape(-1), torch.zeros((), device=dev)))
            a_n.index_add_(0, ni_d, keep.float())
    eps = 1e-5
    res = {"sim-joint": sim_j, "sim-dual": sim_d, "overlap": cover}
    res["sim"] = (res["sim-joint"] + res["sim-dual"]) / 2
    res["alignability-dual"] = a_d / a_n.clamp(min=eps)
    res["alignability-joint"] = a_j / a_n.clamp(min=eps)
    return res


def predicted_frames(sim: torch.Tensor) -> torch.Tensor:
    """Per sentence, the frame the alignment picks (eval_zeroshot_align.py:222-238: uncovered entries count as
    -6e4, softmax over time, argmax): tan_align_argmax, a warp per sentence."""
    _require_cuda(sim, "predicted_frames")
    return ops.align_argmax(sim.float().contiguous())


@torch.no_grad()
def global_alignment(model: TemporalAligner, video: torch.Tensor, text_embed: torch.Tensor, seq_len: int) -> dict:
    """The 'global' method of eval/eval_zeroshot_align.py:207-215: the whole video in ONE pass, the positional table
    of the first `seq_len` positions interpolated to `vlen` frames (get_text_visual_sim of train/main.py:171-189 with
    `interpolate_from=seq_len`).  video [vlen, D_in], text_embed [n_text, D_text] (CUDA).  Returns fp32 tensors
    'sim-joint' / 'sim-dual' [n_text, vlen] = last-stage logits / 0.07, 'sim' = the JOINT one (:209; this method does
    not average the two), and 'alignability-dual' / 'alignability-joint' [n_text]: the head's outputs (dual, and the
    LAST joint stage, :211-212) or without a head the sentence's largest similarity over the video (:214-215)."""
    _require_cuda(video, "global_alignment")
    v, t = video[None].float(), text_embed[None].float()
    inv_tau = 1.0 / 0.07
    sim_j = (model.get_text_visual_sim_joint(v, t, seq_len)[0, -1].t() * inv_tau).contiguous()
    sim_d = (model.get_text_visual_sim_dual(v, t, seq_len)[0, -1].t() * inv_tau).contiguous()
    res = {"sim-joint": sim_j, "sim-dual": sim_d, "sim": sim_j}
    if model.use_alignability_head:
        a = model.get_alignability(v, t, seq_len, None)
        res["alignability-dual"] = a["alignability-dual"][0, :, 0].float()
        res["alignability-joint"] = a["alignability-joint"][0, -1, :, 0].float()
    else:
        res["alignability-dual"] = sim_d.amax(dim=-1)
        res["alignability-joint"] = sim_j.amax(dim=-1)
    return res


def roc_auc(targets, scores) -> float:
    """Area under the ROC curve of binary `targets` against `scores` -- what sklearn.metrics.roc_auc_score returns at
    eval/eval_zeroshot_align.py:247 -- as the Mann-Whitney statistic with tied scores counted half."""
    y = np.asarray(targets).astype(bool).reshape(-1)
    s = np.asarray(scores, dtype=np.float64).reshape(-1)
    if y.shape != s.shape:
        raise ValueError(f"roc_auc: {y.shape[0]} targets for {s.shape[0]} scores")
    n_pos, n_neg = int(y.sum()), int((~y).sum())
    if n_pos == 0 or n_neg == 0:
        raise ValueError("roc_auc: only one class present in targets")
    _, inv, cnt = np.unique(s, return_inverse=True, return_counts=True)
    last = np.cumsum(cnt)                                   # 1-based rank of the last member of every tie group
    rank = (last - (cnt - 1) / 2.0)[inv]                    # average rank inside the group
    return float((rank[y].sum() - n_pos * (n_pos + 1) / 2.0) / (n_pos * n_neg))


class AlignmentMeter:
    """The HTM-Align bookkeeping of eval/eval_zeroshot_align.py:217-249 over the videos of a test set: Recall = the
    share of alignable sentences whose predicted frame lies within [floor(start), ceil(end)] (:233-236), AUC = the
    ROC-AUC of the per-sentence alignability score against the alignable flags (:222-226, :247)."""

    def __init__(self, use_alignability_head: bool):
        self.use_alignability_head = bool(use_alignability_head)
        self.hits: List[bool] = []
        self.scores: List[np.ndarray] = []
        self.targets: List[np.ndarray] = []

    def update(self, result: dict, tgt_aligned: Sequence[bool], start: Sequence[float], end: Sequence[float]) -> None:
        """`result` from sliding_window_alignment / global_alignment; tgt_aligned / start / end per sentence (frames)."""
        sim = result["sim"].float()
        aligned = np.asarray(tgt_aligned).astype(bool)
        if aligned.shape[0] != sim.shape[0]:
            raise TanError(f"AlignmentMeter: {aligned.shape[0]} sentence flags for a [{sim.shape[0]}, .] similarity")
        frames = predicted_frames(sim).cpu().numpy()
        if self.use_alignability_head:
            score = result["alignability-joint"]                                           # :217-218
        else:
            score = torch.where(sim != 0, sim, torch.full_like(sim, -6e4)).amax(dim=-1)    # :220, :226
        self.scores.append(score.float().cpu().numpy())
        self.targets.append(aligned.astype(np.int64))
        s, e = np.floor(np.asarray(start, dtype=np.float64)), np.ceil(np.asarray(end, dtype=np.float64))
        for n in np.flatnonzero(aligned):
            self.hits.append(bool(s[n] <= frames[n] <= e[n]))

    def compute(self) -> dict:
        return {"Recall": float(np.mean(self.hits)),
                "AUC": roc_auc(np.concatenate(self.targets, 0), np.concatenate(self.scores, 0))}


@torch.no_grad()
def evaluate_alignment(model: TemporalAligner, samples, seq_len: int, method: str = "overlap-seq", embed_text=None,
                       max_windows_per_batch: int = 256) -> dict:
    """The loop of `test_alignment_htm` (eval/eval_zeroshot_align.py:97-252) over an iterable of HTM-Align samples
    as its loader yields them (:63-78): 'video' [vlen, D_in] (or [1, vlen, D_in]), 'start' / 'end' [n_text] in
    frames, 'aligned' [n_text] flags and either 'text_embed' [n_text, D_text] or 'str' (sentences) together with
    `embed_text(list of str) -> [n_text, D_text]` (the tokenizer + `model.lang_model` of train/main.py:172-175).
    method 'overlap-seq' (:127-205; the windows follow the NON-alignable sentences, :145-154) or 'global'
    (:207-215).  Returns {'Recall', 'AUC'} (:249)."""
    if method not in ("overlap-seq", "global"):
        raise TanError(f"evaluate_alignment: unknown method {method!r} (overlap-seq | global)")
    dev = next(model.parameters()).device
    meter = AlignmentMeter(model.use_alignability_head)
    for sample in samples:
        video = torch.as_tensor(sample["video"])
        video = (video[0] if video.dim() == 3 else video).to(dev).float()
        if "text_embed" in sample:
            text = torch.as_tensor(sample["text_embed"])
        elif embed_text is not None:
            text = embed_text(list(sample["str"]))
        else:
            raise TanError("evaluate_alignment: a sample needs 'text_embed', or 'str' plus an embed_text callable")
        text = text.to(dev).float()
        start = torch.as_tensor(sample["start"]).reshape(-1).double().cpu().numpy()
        end = torch.as_tensor(sample["end"]).reshape(-1).double().cpu().numpy()
        aligned = torch.as_tensor(sample["aligned"]).reshape(-1).cpu().numpy().astype(bool)
        if not (len(start) == len(end) == len(aligned) == text.shape[0]):
            raise TanError(f"evaluate_alignment: {text.shape[0]} sentences, {len(start)} / {len(end)} timestamps, "
                           f"{len(aligned)} flags")
        if method == "overlap-seq":
            windows = plan_windows(video.shape[0], seq_len, (start + end) / 2, ~aligned)
            res = sliding_window_alignment(model, video, text, windows, max_windows_per_batch)
        else:
            res = global_alignment(model, video, text, seq_len)
        meter.update(res, aligned, start, end)
    return meter.compute()
