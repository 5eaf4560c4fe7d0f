"""Import the UNMODIFIED reference modules from /root/reference (test infrastructure only).

TEST INFRASTRUCTURE -- never imported by the product package (`temporalalignnet_b200`).
Only `tests/`, `oracle/make_golden.py` and the dev-container validation of `oracle/tan_oracle.py`
use it.  `/root/reference` exists only in the build container, never on the GPU box, so every
caller must gate on `reference_available()`.

Stubs needed to import the reference here:
  * `matplotlib`, `matplotlib.pyplot`, `ffmpeg`  -> empty modules (train/loss.py:4,:12 import them
    at module top; they are only used by the dead `visualize` path, train/loss.py:376,:426).
  * `tan_model.Word2VecModel` -> empty nn.Module (model/tan_model.py:40 constructs it; its
    weights `s3d_dict.npy` are not in the repo, model/readme.md:11-14; text embeddings are
    inputs to the path anyway).
"""
import os
import sys
import types

REF_ROOT = os.environ.get("TAN_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "model", "tan_model.py"))


_loaded = {}


def load_reference():
    """Returns (tfm_model, tan_model, loss) reference modules."""
    if _loaded:
        return _loaded["tfm"], _loaded["tan"], _loaded["loss"]
    if not reference_available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    import torch.nn as nn

    for name in ("matplotlib", "matplotlib.pyplot", "ffmpeg"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            if name == "matplotlib.pyplot":
                m.switch_backend = lambda *a, **k: None
            sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    for p in (REF_ROOT, os.path.join(REF_ROOT, "model"), os.path.join(REF_ROOT, "train")):
        if p not in sys.path:
            sys.path.insert(0, p)
    if "word2vec_model" not in sys.modules:
        w2v = types.ModuleType("word2vec_model")

        class Word2VecModel(nn.Module):  # stub: the text backbone is upstream of the path
            def __init__(self, *a, **k):
                super().__init__()

        w2v.Word2VecModel = Word2VecModel
        sys.modules["word2vec_model"] = w2v
    import tfm_model  # noqa: E402
    import tan_model  # noqa: E402
    import loss as ref_loss  # noqa: E402

    _loaded.update(tfm=tfm_model, tan=tan_model, loss=ref_loss)
    return tfm_model, tan_model, ref_loss
