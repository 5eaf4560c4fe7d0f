"""Shim so that the reference's `from tan_model import TemporalAligner, TwinTemporalAligner`
(train/main.py:20-21) resolves to the B200 implementation: put this directory on sys.path instead
of the reference's `../model/`."""
from temporalalignnet_b200.tan_model import TemporalAligner, TwinTemporalAligner, LazyLogits  # noqa: F401
