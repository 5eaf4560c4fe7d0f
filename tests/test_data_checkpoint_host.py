"""Host logic of SURVEY.md 8(f) f4 (CPU): the pinned ragged collate against the reference's collate_fn
(data/loader_htm.py:111-129) / train/loss.py:32-39 / train/main.py:52-65, and checkpoint compatibility
(train/main.py:462-470: `online.` / `target.` / `lang_model.` prefixes)."""
import os

import numpy as np
import pytest
import torch

from oracle.ref_loader import REF_ROOT, reference_available
from temporalalignnet_b200 import synth


def _samples(B=4, T=20, D=16, seed=3):
    """Dataset items with the reference's keys (data/loader_htm.py:156-165): ragged video lengths and sentence counts."""
    r = np.random.default_rng(seed)
    out = []
    for b in range(B):
        t = T if b != 1 else T - 5                                   # one shorter clip: padded by its last frame
        n = int(r.integers(1, 5))
        s = np.sort(r.integers(0, t - 1, size=n))
        e = np.minimum(s + r.integers(1, 6, size=n), t)
        out.append({'video': torch.from_numpy(r.standard_normal((t, D)).astype(np.float32)),
                    'padding_mask': torch.zeros(t).long(), 'vid': f'v{b}', 'text': [f's{b}_{i}' for i in range(n)],
                    'start': [int(x) for x in s], 'end': [int(x) for x in e],
                    'token': torch.from_numpy(r.integers(0, 50, size=(n, 32))),
                    'abs_text_start': s.astype(np.float32) / 100, 'abs_text_end': e.astype(np.float32) / 100})
    return out


def test_collate_keeps_reference_keys_and_adds_padded_tensors():
    from temporalalignnet_b200 import data, loss
    items = _samples()
    out = data.collate_fn(items, pin=False)
    B, T = 4, 20
    # the reference's own keys and types
    assert tuple(out['video'].shape) == (B, T, 16) and tuple(out['padding_mask'].shape) == (B, T)
    assert torch.equal(out['video'][1, 15:], items[1]['video'][-1].expand(5, -1))        # padded by the LAST frame
    assert out['padding_mask'][1, 15:].eq(1).all() and out['padding_mask'][1, :15].eq(0).all()
    for k in ('text', 'start', 'end', 'vid', 'token', 'abs_text_start', 'abs_text_end'):
        assert isinstance(out[k], list) and len(out[k]) == B
    # added tensors == what train/loss.py:32-39 / train/main.py:52-65 build per step
    n = [len(s['start']) for s in items]
    N = max(n)
    assert out['n_sentences'].tolist() == n
    m_ref, s_ref, e_ref = loss.get_mask_from_time(out['start'], out['end'], T, N, device='cpu')
    assert torch.equal(out['start_pad'], s_ref) and torch.equal(out['end_pad'], e_ref)
    tpm_ref = torch.nn.utils.rnn.pad_sequence(torch.split(torch.zeros(sum(n)), n, dim=0), batch_first=True, padding_value=1)
    assert torch.equal(out['text_padding_mask'], tpm_ref)
    assert torch.equal(out['token_flat'], torch.cat([s['token'] for s in items], 0).long())
    emb = torch.randn(sum(n), 8)
    padded = data.pad_text_embed(emb, n)
    assert tuple(padded.shape) == (B, N, 8) and torch.equal(padded[0, n[0] - 1], padded[0, -1])


@pytest.mark.skipif(not reference_available(), reason="/root/reference only exists in the build container")
def test_collate_matches_the_reference_collate_fn():
    """The shared keys are bit-identical to data/loader_htm.py's collate_fn (loaded with stubs for its dataset-only
    imports; the static method itself is pure torch)."""
    import importlib.util
    import sys
    import types
    from oracle.ref_loader import load_reference
    load_reference()
    for name in ("simplejson", "tqdm", "pandas"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                m.tqdm = lambda x, *a, **k: x
                sys.modules[name] = m
    w2v = sys.modules["word2vec_model"]
    if not hasattr(w2v, "Word2VecTokenizer"):
        w2v.Word2VecTokenizer = type("Word2VecTokenizer", (), {})
    spec = importlib.util.spec_from_file_location("ref_loader_htm", os.path.join(REF_ROOT, "data", "loader_htm.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    cls = [v for v in vars(ref).values() if isinstance(v, type) and hasattr(v, "collate_fn")][0]
    from temporalalignnet_b200 import data
    items = _samples(seed=9)
    a, b = cls.collate_fn(items), data.collate_fn(items, pin=False)
    for k, v in a.items():
        if torch.is_tensor(v):
            assert torch.equal(v, b[k]), k
        else:
            assert len(v) == len(b[k]) and all((x == y) if not torch.is_tensor(x) and not isinstance(x, np.ndarray)
                                               else np.array_equal(np.asarray(x), np.asarray(y)) for x, y in zip(v, b[k])), k
    assert torch.equal(ref.pad_sequence_by_last([s['video'] for s in items]), data.pad_sequence_by_last([s['video'] for s in items]))


def test_reference_shaped_cotrain_checkpoint_loads_strict_clean():
    """train/main.py:462-470: a stage-1 state dict (+ `lang_model.` text-backbone keys) initialises the twin model."""
    from temporalalignnet_b200 import TwinTemporalAligner, TemporalAligner, checkpoint
    E = D = 1
    sd = {k: torch.from_numpy(v) for k, v in synth.make_state_dict(E, D).items()}
    ref_ckpt = {"epoch": 3, "state_dict": dict(sd), "optimizer": {}}
    ref_ckpt["state_dict"].update({"lang_model.fc1.weight": torch.zeros(2048, 300), "lang_model.fc1.bias": torch.zeros(2048)})
    single = TemporalAligner(E, D, random_pos_start=0)
    missing, unexpected = checkpoint.load_reference_checkpoint(single, ref_ckpt)
    assert missing == [] and sorted(unexpected) == ["lang_model.fc1.bias", "lang_model.fc1.weight"]
    assert torch.equal(single.video_pre_proj.weight, sd["video_pre_proj.weight"])
    twin = TwinTemporalAligner(m=0.999, num_encoder_layers=E, num_decoder_layers=D)
    remapped = checkpoint.remap_for_cotrain(ref_ckpt["state_dict"])
    assert set(k.split(".")[0] for k in remapped) == {"online", "target", "lang_model"}
    missing, unexpected = checkpoint.load_reference_checkpoint(twin, ref_ckpt, cotrain_from_init=True)
    assert missing == []
    for (k, po), pt in zip(twin.online.named_parameters(), twin.target.parameters()):
        assert torch.equal(po, pt) and not pt.requires_grad
        if k in sd:
            assert torch.equal(po.detach(), sd[k]), k
    # a `_cotrain_` checkpoint (already prefixed, DataParallel `module.` wrapper) loads as it is
    ck2 = {"state_dict": {"module." + k: v for k, v in twin.state_dict().items()}}
    twin2 = TwinTemporalAligner(m=0.999, num_encoder_layers=E, num_decoder_layers=D)
    missing, unexpected = checkpoint.load_reference_checkpoint(twin2, ck2)
    assert missing == [] and unexpected == []
    assert all(torch.equal(a, b) for a, b in zip(twin.state_dict().values(), twin2.state_dict().values()))
    # a wrong checkpoint is refused
    from temporalalignnet_b200 import TanError
    with pytest.raises(TanError):
        checkpoint.load_reference_checkpoint(TemporalAligner(2, 2), ref_ckpt)


def test_text_backbone_keys_route_to_an_attached_backbone():
    from temporalalignnet_b200 import TemporalAligner, checkpoint
    from temporalalignnet_b200.word2vec_model import Word2VecModel
    lang = Word2VecModel(num_embeddings=50)
    m = TemporalAligner(1, 1, random_pos_start=0, lang_module=lang)
    sd = {k: torch.from_numpy(v) for k, v in synth.make_state_dict(1, 1).items()}
    w = torch.randn(2048, 300)
    sd.update({"lang_model." + k: v.clone() for k, v in lang.state_dict().items()})
    sd["lang_model.fc1.weight"] = w
    missing, unexpected = checkpoint.load_reference_checkpoint(m, {"state_dict": sd})
    assert missing == [] and unexpected == []
    assert torch.equal(m.lang_model.fc1.weight.detach(), w)
