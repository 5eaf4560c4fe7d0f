"""Input pipeline of the hot path: the reference's ragged collate (data/loader_htm.py:111-129)
with its outputs in PINNED host memory and, next to the reference's own keys, the padded per-sentence tensors the
loss needs -- so that a step starts with a handful of asynchronous H2D copies instead of python lists ->
`pad_sequence` -> H2D inside `get_loss` (train/loss.py:32-39) and `pad_sequence(torch.split(...))` on the device
(train/main.py:52-65).

    loader = DataLoader(dataset, batch_size=B, collate_fn=temporalalignnet_b200.data.collate_fn, ...)
    batch  = to_device(next(iter(loader)), 'cuda')            # non-blocking copies of the pinned tensors

Keys of the reference's collate_fn are kept with the same types ('video' [B,T,D] padded by the last frame,
'padding_mask' [B,T] padded with 1, 'text' / 'start' / 'end' / 'vid' / 'token' python lists, optional 'cut_*' /
'abs_text_*'), so train/main.py runs unchanged; the ADDED keys are
    'n_sentences'        [B] int64                     sentences per clip (= num_sentence_per_sample, main.py:52)
    'start_pad','end_pad' [B, N] float32               padded with T+100 / -100 (train/loss.py:35-38)
    'text_padding_mask'  [B, N] float32 (1 = padding)  what main.py:62-65 builds on the device
    'token_flat'         [sum n, 32] int64             torch.concat(token_list) of main.py:53-54
`get_loss` uses 'start_pad' / 'end_pad' when they are present (no per-step list processing).
"""
from __future__ import annotations

from typing import Dict, List

import torch
from torch.nn.utils.rnn import pad_sequence


def pad_sequence_by_last(sequences: List[torch.Tensor]) -> torch.Tensor:
    """data/loader_htm.py:13-23: pad a list of [L_i, ...] tensors to the longest by repeating each one's LAST row."""
    trailing = tuple(sequences[0].shape[1:])
    max_len = max(int(s.shape[0]) for s in sequences)
    out = sequences[0].new_zeros((len(sequences), max_len) + trailing)
    for i, t in enumerate(sequences):
        n = int(t.shape[0])
        out[i, :n] = t
        out[i, n:] = t[-1]
    return out


def _pin(t: torch.Tensor, pin: bool) -> torch.Tensor:
    return t.pin_memory() if (pin and torch.cuda.is_available() and not t.is_pinned()) else t


def collate_fn(batch: List[dict], pin: bool = True) -> Dict[str, object]:
    """data/loader_htm.py:111-129 + the padded tensors of train/loss.py:32-39 and train/main.py:52-65 (module doc)."""
    out: Dict[str, object] = {}
    out['video'] = _pin(pad_sequence_by_last([s['video'] for s in batch]), pin)
    out['padding_mask'] = _pin(pad_sequence([s['padding_mask'] for s in batch], batch_first=True, padding_value=1.0), pin)
    for k in ('text', 'start', 'end', 'vid', 'token'):
        out[k] = [s[k] for s in batch]
    for k in ('cut_start', 'cut_end', 'abs_text_start', 'abs_text_end'):
        if k in batch[0]:
            out[k] = [s[k] for s in batch]
    B = len(batch)
    T = int(out['video'].shape[1])
    n = [len(s['start']) for s in batch]
    N = max(n)
    start = torch.full((B, N), float(T) + 1e2)
    end = torch.full((B, N), -1e2)
    tpm = torch.ones(B, N)
    for b, s in enumerate(batch):
        if n[b]:
            start[b, :n[b]] = torch.as_tensor(s['start'], dtype=torch.float32)
            end[b, :n[b]] = torch.as_tensor(s['end'], dtype=torch.float32)
            tpm[b, :n[b]] = 0.0
    out['n_sentences'] = _pin(torch.tensor(n, dtype=torch.int64), pin)
    out['start_pad'], out['end_pad'], out['text_padding_mask'] = _pin(start, pin), _pin(end, pin), _pin(tpm, pin)
    toks = [torch.as_tensor(s['token']) for s in batch]
    if all(t.dim() == 2 for t in toks):
        out['token_flat'] = _pin(torch.cat(toks, 0).long(), pin)
    return out


def to_device(batch: Dict[str, object], device) -> Dict[str, object]:
    """Non-blocking H2D copies of every tensor of a collated batch (lists stay on the host)."""
    return {k: (v.to(device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in batch.items()}


def pad_text_embed(text_embed_flat: torch.Tensor, n_sentences: List[int]) -> torch.Tensor:
    """train/main.py:61: pad_sequence_by_last(torch.split(text_embed, num_sentence_per_sample)) -> [B, N, C]."""
    return pad_sequence_by_last(list(torch.split(text_embed_flat, list(n_sentences), dim=0)))
