"""Shim so that the reference's `from tan_model import TemporalAligner, TwinTemporalAligner`
(train/main.py:20-21) resolves to the B200 implementation: put this directory on sys.path instead
of the reference's `../model/`.

train/main.py trains: it calls `model.train()`, `loss.backward()` (:112) and an optimizer.  The classes exported
here therefore have the training step switched on at construction (`enable_autograd(True)`), so the driver runs
unchanged and without environment variables; under `torch.no_grad()` / `model.eval()` (its evaluation code)
they take the inference path like the base classes.  Same class names, constructor arguments and state-dict keys."""
from temporalalignnet_b200 import tan_model as _tm
from temporalalignnet_b200.tan_model import LazyLogits  # noqa: F401


class TemporalAligner(_tm.TemporalAligner):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.enable_autograd(True)


class TwinTemporalAligner(_tm.TwinTemporalAligner):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.enable_autograd(True)          # the online network; the EMA target never trains (model/tan_model.py:334-338)
