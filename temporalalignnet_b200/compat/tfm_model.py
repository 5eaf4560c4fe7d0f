"""Shim for `from tfm_model import TemporalEncoder, get_position_embedding_sine` (model/tan_model.py:9)."""
from temporalalignnet_b200.tfm_model import (TemporalEncoder, TemporalDecoder, ResidualAttentionBlock_Step,  # noqa: F401
                                             ResidualDecoderBlock_Step, QuickGELU, PositionEmbeddingSine,
                                             get_position_embedding_sine)
