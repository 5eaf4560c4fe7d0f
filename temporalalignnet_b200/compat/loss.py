"""Shim for `from loss import get_loss, get_mask_from_time, get_text_pos` (train/main.py:16)."""
from temporalalignnet_b200.loss import get_loss, get_mask_from_time, get_text_pos, circulant  # noqa: F401
