"""B200-native TAN hot path: the reference's `TemporalAligner` / `TemporalEncoder` / `get_loss`
surface (TengdaHan/TemporalAlignNet, model/tan_model.py, model/tfm_model.py, train/loss.py) on
hand-written sm_100a kernels reached through the C ABI in include/tan_b200.h.

Importing this package does not load the CUDA library; the first kernel call does, and raises
`TanError` if libtan_b200.so is missing -- there is no CPU or eager fallback.
"""
from ._lib import TanError  # noqa: F401

__all__ = ["TanError", "TemporalAligner", "TwinTemporalAligner", "TemporalEncoder", "TemporalDecoder",
           "get_loss", "get_mask_from_time", "get_text_pos", "LazyLogits"]


def __getattr__(name):
    if name in ("TemporalAligner", "TwinTemporalAligner", "LazyLogits"):
        from . import tan_model
        return getattr(tan_model, name)
    if name in ("TemporalEncoder", "TemporalDecoder", "get_position_embedding_sine"):
        from . import tfm_model
        return getattr(tfm_model, name)
    if name in ("get_loss", "get_mask_from_time", "get_text_pos", "circulant"):
        from . import loss
        return getattr(loss, name)
    raise AttributeError(name)
