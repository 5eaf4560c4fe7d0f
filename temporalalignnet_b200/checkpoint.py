"""Checkpoint compatibility with the reference: the key remapping train/main.py:462-470 applies
when a stage-1 (`init`) checkpoint initialises the co-training model, and loading of reference-shaped checkpoints
(`{'state_dict': ...}`, optional `module.` prefixes of DataParallel, `lang_model.*` / `bert.*` text-backbone keys)
into the B200 classes, whose parameter names equal the reference's.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from ._lib import TanError


def strip_module_prefix(state_dict: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Checkpoints saved from nn.DataParallel carry a `module.` prefix (train/main.py:403-407 wraps the model)."""
    if state_dict and all(k.startswith("module.") for k in state_dict):
        return {k[len("module."):]: v for k, v in state_dict.items()}
    return dict(state_dict)


def remap_for_cotrain(state_dict: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """train/main.py:464-469: a single-model state dict becomes `target.*` + `online.*` copies; text-backbone keys
    (`lang_model.`) are also kept under their own names."""
    out = {f"target.{k}": v for k, v in state_dict.items()}
    out.update({f"online.{k}": v for k, v in state_dict.items()})
    out.update({k: v for k, v in state_dict.items() if 'lang_model.' in k})
    return out


def load_reference_checkpoint(model, checkpoint, cotrain_from_init: bool = False, strict: bool = True
                              ) -> Tuple[list, list]:
    """Load a reference checkpoint (the dict train/main.py saves: {'state_dict': ..., 'epoch': ..., ...}, or a bare
    state dict) into a TemporalAligner / TwinTemporalAligner of this package.

    * `cotrain_from_init`: apply train/main.py:464-469 (stage-1 weights -> online + target of the twin model) and,
      like :483, copy the online parameters into the target.
    * the reference registers its text backbone as `bert.*` (model/tan_model.py:38-40) and the training driver reads
      it as `lang_model`; keys under either name are routed to the attached backbone when one is attached and are
      dropped (reported as unexpected only with strict=False) when the slot is empty -- the published checkpoints
      carry the word2vec weights there (readme.md:45).
    Returns (missing_keys, unexpected_keys); raises TanError on a strict mismatch."""
    sd = checkpoint.get("state_dict", checkpoint) if isinstance(checkpoint, dict) else checkpoint
    sd = strip_module_prefix(sd)
    if cotrain_from_init:
        sd = remap_for_cotrain(sd)
    own = model.state_dict()
    routed = {}
    dropped = []
    for k, v in sd.items():
        k2 = k
        for a, b in (("lang_model.", "bert."), ("online.lang_model.", "online.bert."), ("target.lang_model.", "target.bert.")):
            if k.startswith(a):
                k2 = b + k[len(a):]
        if k2 in own:
            routed[k2] = v
        elif ".bert." in "." + k2 or k2.startswith("bert."):
            dropped.append(k)                     # text-backbone weights without an attached backbone
        else:
            routed[k2] = v
    missing, unexpected = model.load_state_dict(routed, strict=False)
    missing = [k for k in missing if not (k.startswith("bert.") or ".bert." in k)] if not _has_backbone(model) else list(missing)
    if strict and (missing or unexpected):
        raise TanError(f"checkpoint does not match the model: missing {missing[:8]}, unexpected {list(unexpected)[:8]}")
    if cotrain_from_init and hasattr(model, "_copy_param"):
        model._copy_param()
    return missing, list(unexpected) + dropped


def _has_backbone(model) -> bool:
    bert = getattr(model, "bert", None)
    return bert is not None and len(list(bert.parameters())) > 0
