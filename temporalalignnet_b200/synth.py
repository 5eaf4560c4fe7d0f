"""Deterministic synthetic weights and batches for the TAN hot path.

Everything is drawn from numpy's PCG64 `default_rng`, one independent stream per tensor (seeded by
`crc32(name) ^ seed`), so a tensor's values do not depend on construction order, torch version or
device.  Shapes/statistics follow SURVEY.md section 8(d):

  * weights: the reference's init statistics (`model/tan_model.py:76-97`: attn std d^-1/2,
    proj std d^-1/2 (2D)^-1/2, fc std (2d)^-1/2 -- all derived from the JOINT encoder; pre-proj,
    positional tables std 0.01), optionally perturbed LayerNorm affine / linear biases so parity
    tests exercise every parameter (the reference leaves them at 1/0).
  * batch: `video_embed = randn(B,T,D_in)`, `N = max(T/8,1)`, `n_b = randint(N/2, N+1)` real
    sentences per clip, text rows >= n_b replicate row n_b-1 (mimics `pad_sequence_by_last`,
    `data/loader_htm.py:13-23`), `start = sorted(randint(0,T-1,n_b))`,
    `end = min(start + randint(1,9), T)` as python lists (the reference's input type,
    `train/loss.py:32-39`).
"""
from __future__ import annotations

import zlib
from typing import Dict, List, Optional

import numpy as np


def _rng(name: str, seed: int) -> np.random.Generator:
    return np.random.default_rng([zlib.crc32(name.encode()), seed])


def _normal(name, seed, shape, std=1.0, mean=0.0):
    return (mean + std * _rng(name, seed).standard_normal(shape)).astype(np.float32)


def make_state_dict(num_encoder_layers: int, num_decoder_layers: int, width: int = 512,
                    d_in: int = 1024, text_dim: int = 512, use_alignability_head: bool = False,
                    seed: int = 888, perturb: bool = True, n_pos: int = 1024) -> Dict[str, np.ndarray]:
    """float32 numpy state-dict with the reference's key names."""
    d = width
    sd: Dict[str, np.ndarray] = {}
    proj_std = (d ** -0.5) * ((2 * max(num_decoder_layers, 1)) ** -0.5)
    attn_std = d ** -0.5
    fc_std = (2 * d) ** -0.5
    aff = 0.1 if perturb else 0.0
    bias_std = 0.02 if perturb else 0.0

    def ln(prefix):
        sd[prefix + ".weight"] = _normal(prefix + ".weight", seed, (d,), aff, 1.0)
        sd[prefix + ".bias"] = _normal(prefix + ".bias", seed, (d,), aff, 0.0)

    sd["temporal_pos_embed"] = _normal("temporal_pos_embed", seed, (n_pos, d), 0.01)
    sd["text_temporal_pos_embed"] = _normal("text_temporal_pos_embed", seed, (n_pos, d), 0.01)
    for stack, n_layers in (("video_temporal_encoder", num_encoder_layers),
                            ("joint_temporal_encoder", num_decoder_layers)):
        for i in range(n_layers):
            p = f"{stack}.resblocks.{i}."
            sd[p + "attn.in_proj_weight"] = _normal(p + "attn.in_proj_weight", seed, (3 * d, d), attn_std)
            sd[p + "attn.in_proj_bias"] = _normal(p + "attn.in_proj_bias", seed, (3 * d,), bias_std)
            sd[p + "attn.out_proj.weight"] = _normal(p + "attn.out_proj.weight", seed, (d, d), proj_std)
            sd[p + "attn.out_proj.bias"] = _normal(p + "attn.out_proj.bias", seed, (d,), bias_std)
            ln(p + "ln_1")
            sd[p + "mlp.c_fc.weight"] = _normal(p + "mlp.c_fc.weight", seed, (4 * d, d), fc_std)
            sd[p + "mlp.c_fc.bias"] = _normal(p + "mlp.c_fc.bias", seed, (4 * d,), bias_std)
            sd[p + "mlp.c_proj.weight"] = _normal(p + "mlp.c_proj.weight", seed, (d, 4 * d), proj_std)
            sd[p + "mlp.c_proj.bias"] = _normal(p + "mlp.c_proj.bias", seed, (d,), bias_std)
            ln(p + "ln_2")
    sd["video_pre_proj.weight"] = _normal("video_pre_proj.weight", seed, (d, d_in), 0.01)
    sd["text_pre_proj.weight"] = _normal("text_pre_proj.weight", seed, (d, text_dim), 0.01)
    for name in ("ln_text_init", "ln_video_init", "ln_position_init", "ln_video_post_enc",
                 "ln_joint_post_enc"):
        ln(name)
    sd["mlp.weight"] = _normal("mlp.weight", seed, (d, d), 0.01)
    sd["mlp.bias"] = np.zeros((d,), np.float32)
    if use_alignability_head:
        sd["binary_head.weight"] = _normal("binary_head.weight", seed, (1, d), 0.01 if not perturb else 0.05)
        sd["binary_head.bias"] = _normal("binary_head.bias", seed, (1,), bias_std)
    return sd


def make_batch(B: int, T: int, N: Optional[int] = None, d_in: int = 1024, text_dim: int = 512,
               seed: int = 888, pad_video_every: int = 0, tag: str = "", force_full: bool = False) -> dict:
    """One synthetic batch.  Returns numpy arrays + python lists (the reference's input types).

    pad_video_every=k>0 pads a suffix (T/4 frames) of every k-th clip (coverage variant of 8(d)).
    `tag` decorrelates batches drawn with the same seed (e.g. per rank / per step).
    force_full gives clip 0 all N sentences: the reference's get_mask_from_time only broadcasts when the
    longest clip has exactly N sentences (train/loss.py:31-40).
    """
    if N is None:
        N = max(T // 8, 1)
    r = _rng(f"batch{tag}", seed)
    video = r.standard_normal((B, T, d_in)).astype(np.float32)
    text = r.standard_normal((B, N, text_dim)).astype(np.float32)
    n_b = r.integers(max(N // 2, 1), N + 1, size=B)
    if force_full:
        n_b[0] = N
    text_padding_mask = np.zeros((B, N), bool)
    video_padding_mask = np.zeros((B, T), bool)
    start: List[List[float]] = []
    end: List[List[float]] = []
    sentences: List[List[str]] = []
    for b in range(B):
        nb = int(n_b[b])
        text[b, nb:] = text[b, nb - 1]
        text_padding_mask[b, nb:] = True
        s = np.sort(r.integers(0, max(T - 1, 1), size=nb))
        e = np.minimum(s + r.integers(1, 9, size=nb), T)
        start.append([float(v) for v in s])
        end.append([float(v) for v in e])
        sentences.append([f"clip{b}-sent{i}" for i in range(nb)])
        if pad_video_every and b % pad_video_every == pad_video_every - 1:
            video_padding_mask[b, T - max(T // 4, 1):] = True
    return {
        "video": video, "text": text,
        "video_padding_mask": video_padding_mask, "text_padding_mask": text_padding_mask,
        "start": start, "end": end, "text_str": sentences, "n_b": n_b.astype(np.int64),
    }


def make_logit_case(B: int, S: int, T: int, N: int, seed: int = 888, tag: str = "", pad_video_every: int = 0) -> dict:
    """Synthetic cosine logits [B,S,T,B,N] for loss-only parity cases (no model): uniform noise in
    [-0.1, 0.1] plus, in every own-clip block, one sloped bump per sentence (peak 0.45, +-10 frames wide, wider
    than any sentence; usually at the same place for the dual and the joint model), so that every candidate
    window of the self-labelling scan (train/loss.py:133-136) trades frames of significant probability and
    its argmax is decisive -- on a flat profile near-ties flip with the summation order.
    Also EMA logits, alignability-head logits, absolute text positions and a batch whose clip 0 has all N
    sentences."""
    batch = make_batch(B, T, N, seed=seed, tag=f"lc{tag}", force_full=True, pad_video_every=pad_video_every)
    r = _rng(f"logits{tag}", seed)
    out = {"batch": batch}
    # bump centres stay inside the frames that are never padded (a sentence whose evidence lies in padded
    # frames sees constant probabilities there and its argmax is decided by rounding, in the reference too)
    c_hi = max(T - max(T // 4, 1) - 11, 3)
    base = [[int(r.integers(2, c_hi)) for _ in range(N)] for _ in range(B)]
    dt = np.arange(-10, 11)
    shape = (0.45 - np.where(dt < 0, 0.04, 0.03) * np.abs(dt)).astype(np.float32)     # asymmetric: no mirrored ties
    for k in ("logits_dual", "logits_joint", "ema-logits_dual", "ema-logits_joint"):
        x = (r.random((B, S, T, B, N), dtype=np.float32) * 2 - 1) * 0.1
        for b in range(B):
            for n in range(N):
                c = base[b][n]                         # the two models mostly agree (IoU >= 0.5 is exercised)
                if r.random() < 0.3:
                    c = int(r.integers(2, c_hi))
                lo, hi = max(c - 10, 0), min(c + 11, T)
                x[b, :, lo:hi, b, n] += shape[lo - (c - 10):hi - (c - 10)] * (1.0 + 0.05 * r.random(hi - lo, dtype=np.float32))
        out[k] = x
    out["dual_logits_alignability"] = r.standard_normal((B, N, 1)).astype(np.float32)
    out["joint_logits_alignability"] = r.standard_normal((B, S, N, 1)).astype(np.float32)
    out["abs_text_pos"] = r.random((B, N, 2), dtype=np.float32)
    return out
