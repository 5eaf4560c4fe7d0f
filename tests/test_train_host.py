"""Host logic of the training step on CPU (`-m "not gpu"`): `train.forward_train` / `train.step_backward` /
`get_loss` run with the C-ABI wrappers replaced by torch stand-ins (tests/cpu_ops.py), so what is tested is the
product's own orchestration -- kernel sequence, saved-activation bookkeeping, stage-gradient injection, row maps of
the video|text concatenation, positional-table slices, the single autograd node -- against torch autograd over the
oracle.  The kernels themselves are tested on the GPU (tests/test_backward_kernels_gpu.py, tests/test_train_gpu.py)."""
import types

import numpy as np
import pytest
import torch

from tests import cpu_ops
from tests.helpers import case_inputs, compare_param_grads, oracle_param_grads


def _args(**kw):
    d = dict(model="init", sim="cos", learn_agreement=0, temporal_agreement_type="keep", loss_threshold=0.0,
             use_alignability_head=0, optim_policy="default")
    d.update(kw)
    return types.SimpleNamespace(**d)


def _model(cfg, sd, head=0, **kw):
    from temporalalignnet_b200 import TemporalAligner
    m = TemporalAligner(num_encoder_layers=cfg["E"], num_decoder_layers=cfg["D"], use_text_pos_enc=cfg["use_text_pos_enc"],
                        use_alignability_head=head, **kw)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items() if head or not k.startswith("binary_head")})
    m.train()
    m.enable_autograd(True)
    return m


def _step(m, batch, args):
    from temporalalignnet_b200 import get_loss
    video, text = torch.from_numpy(batch["video"]), torch.from_numpy(batch["text"])
    vpm, tpm = torch.from_numpy(batch["video_padding_mask"]), torch.from_numpy(batch["text_padding_mask"])
    out = m(video, text, video_padding_mask=vpm, lang_padding_mask=tpm)
    res = get_loss({"start": batch["start"], "end": batch["end"], "text": batch["text_str"]}, video, text,
                   vpm.float(), tpm.float(), out, args, None, shard_batch=False)
    res["loss"].backward()
    return out, res


@pytest.mark.parametrize("name,kw,head", [("g1_e1d1_T32_B4", {}, 0), ("g2_e2d3_T24_B3", {}, 0),
                                          ("g2_e2d3_T24_B3", dict(loss_threshold=0.5, use_alignability_head=1), 1)])
def test_training_step_host_logic_vs_oracle_autograd(monkeypatch, name, kw, head):
    cpu_ops.install(monkeypatch)
    cfg, sd, batch, _ = case_inputs(name)
    args = _args(**kw)
    ref_loss, ref_grads = oracle_param_grads(dict(cfg, head=head), sd, batch, args)
    m = _model(cfg, sd, head=head, random_pos_start=0)
    _, res = _step(m, batch, args)
    assert abs(res["loss"].item() - ref_loss) < 2e-3 * abs(ref_loss), (res["loss"].item(), ref_loss)
    compare_param_grads(m, ref_grads, loose=bool(kw))
    assert m.mlp.weight.grad is None


def test_random_pos_start_offsets_reach_the_right_table_rows(monkeypatch):
    """random_pos_start=1 (the reference's training default, model/tan_model.py:162-165,:195,:224): three draws from
    the global numpy RNG pick the table slices of the video stack, the text positions and the joint stack; the
    gradient of each slice must land in exactly those rows of temporal_pos_embed / text_temporal_pos_embed."""
    cpu_ops.install(monkeypatch)
    cfg, sd, batch, _ = case_inputs("g2_e2d3_T24_B3")
    m = _model(cfg, sd, random_pos_start=1)
    np.random.seed(7)
    out, res = _step(m, batch, _args())
    tape = out["logits_dual"].tape
    starts = (tape.ps_v, tape.ps_t, tape.ps_j)
    assert tape.ps_v != tape.ps_j, "pick a seed whose two video draws differ (exercises the two-slice path)"
    ref_loss, ref_grads = oracle_param_grads(dict(cfg, head=0), sd, batch, _args(), pos_starts=starts)
    assert abs(res["loss"].item() - ref_loss) < 2e-3 * abs(ref_loss)
    compare_param_grads(m, ref_grads)
    T = cfg["T"]
    g = m.temporal_pos_embed.grad
    used = torch.zeros(g.shape[0], dtype=torch.bool)
    used[tape.ps_v:tape.ps_v + T] = True
    used[tape.ps_j:tape.ps_j + T] = True
    assert float(g[~used].abs().max()) == 0.0 and float(g[used].abs().min(dim=1).values.min()) >= 0.0
    assert float(g[used].abs().sum(dim=1).min()) > 0.0


def test_sine_positions_are_constants(monkeypatch):
    """pos_enc='sine': the table is a buffer (model/tan_model.py:60-62); only ln_position_init learns."""
    from temporalalignnet_b200.tfm_model import get_position_embedding_sine
    cpu_ops.install(monkeypatch)
    cfg, sd, batch, _ = case_inputs("g1_e1d1_T32_B4")
    sd = dict(sd)
    sd["temporal_pos_embed"] = get_position_embedding_sine(512, 1024).numpy().astype(np.float32)
    ref_loss, ref_grads = oracle_param_grads(dict(cfg, head=0), sd, batch, _args())
    ref_grads.pop("temporal_pos_embed")
    m = _model(cfg, sd, random_pos_start=0, pos_enc="sine")
    assert "temporal_pos_embed" not in dict(m.named_parameters())
    _, res = _step(m, batch, _args())
    assert abs(res["loss"].item() - ref_loss) < 2e-3 * abs(ref_loss)
    compare_param_grads(m, ref_grads)
    assert m.ln_position_init.weight.grad is not None


def test_inference_paths_carry_no_tape(monkeypatch):
    cpu_ops.install(monkeypatch)
    cfg, sd, batch, _ = case_inputs("g1_e1d1_T32_B4")
    m = _model(cfg, sd, random_pos_start=0)
    video, text = torch.from_numpy(batch["video"]), torch.from_numpy(batch["text"])
    m.two_streams = False
    monkeypatch.setattr(torch.cuda, "current_stream", lambda dev=None: None)
    with torch.no_grad():
        out = m(video, text)
    assert getattr(out["logits_dual"], "tape", None) is None
    m.eval()
    assert getattr(m(video, text)["logits_dual"], "tape", None) is None
    m.train()
    assert getattr(m(video, text)["logits_dual"], "tape", None) is not None


def test_second_backward_on_a_consumed_tape_raises(monkeypatch):
    from temporalalignnet_b200 import TanError
    cpu_ops.install(monkeypatch)
    cfg, sd, batch, _ = case_inputs("g1_e1d1_T32_B4")
    m = _model(cfg, sd, random_pos_start=0)
    _, res = _step(m, batch, _args())
    with pytest.raises((TanError, RuntimeError)):
        res["loss"].backward()


def test_clip_gradients_matches_reference():
    """train.clip_gradients == utils/train_utils.py:3-13 (the reference's own function when it is present)."""
    import os
    import sys

    from temporalalignnet_b200.train import clip_gradients

    def ref_clip(model, clip_grad=3):                       # restatement of utils/train_utils.py:3-13
        norms = []
        for _, p in model.named_parameters():
            if p.grad is not None:
                param_norm = p.grad.data.norm(2)
                norms.append(param_norm.item())
                clip_coef = clip_grad / (param_norm + 1e-6)
                if clip_coef < 1:
                    p.grad.data.mul_(clip_coef)
        return norms

    if os.path.isfile("/root/reference/utils/train_utils.py"):
        sys.path.insert(0, "/root/reference")
        try:
            import importlib.util
            spec = importlib.util.spec_from_file_location("ref_train_utils", "/root/reference/utils/train_utils.py")
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            ref_clip = mod.clip_gradients
        finally:
            sys.path.pop(0)
    torch.manual_seed(0)
    a = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.LayerNorm(16), torch.nn.Linear(16, 4))
    b = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.LayerNorm(16), torch.nn.Linear(16, 4))
    b.load_state_dict(a.state_dict())
    for m in (a, b):
        for i, p in enumerate(m.parameters()):
            g = torch.Generator().manual_seed(i)
            p.grad = torch.randn(p.shape, generator=g) * (10.0 if i % 2 == 0 else 0.01)   # some clipped, some not
        list(m.parameters())[3].grad = None                                              # and one without gradient
    n_ref = ref_clip(a, 3)
    n_got = clip_gradients(b, 3)
    assert len(n_ref) == len(n_got) and all(abs(x - y) <= 1e-6 * max(1.0, abs(x)) for x, y in zip(n_ref, n_got))
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert (pa.grad is None) == (pb.grad is None)
        if pa.grad is not None:
            assert torch.equal(pa.grad, pb.grad)


def test_gradients_reach_the_text_and_video_inputs(monkeypatch):
    """The text backbone trains THROUGH `lang_embed` in the reference (train/main.py:58-60: fc1 / fc2 of the word2vec
    module are ordinary parameters): inputs that require grad get their gradient from the step's autograd node."""
    from oracle import tan_oracle as O
    from temporalalignnet_b200 import get_loss
    cpu_ops.install(monkeypatch)
    cfg, sd, batch, _ = case_inputs("g2_e2d3_T24_B3")
    sd = {k: v for k, v in sd.items() if not k.startswith("binary_head")}
    m = _model(cfg, sd, random_pos_start=0)
    # upstream of the path: a toy "text backbone" whose parameter must receive a gradient through lang_embed
    scale = torch.nn.Parameter(torch.ones(512))
    text0 = torch.from_numpy(batch["text"])
    video = torch.from_numpy(batch["video"]).clone().requires_grad_(True)
    vpm, tpm = torch.from_numpy(batch["video_padding_mask"]), torch.from_numpy(batch["text_padding_mask"])
    text = text0 * scale
    out = m(video, text, video_padding_mask=vpm, lang_padding_mask=tpm)
    res = get_loss({"start": batch["start"], "end": batch["end"], "text": batch["text_str"]}, video, text,
                   vpm.float(), tpm.float(), out, _args(), None, shard_batch=False)
    res["loss"].backward()
    assert scale.grad is not None and video.grad is not None
    # oracle
    sd_t = {k: torch.from_numpy(v) for k, v in sd.items()}
    orc = O.TanOracle(sd_t, cfg["E"], cfg["D"], use_text_pos_enc=cfg["use_text_pos_enc"])
    scale_r = torch.ones(512, requires_grad=True)
    video_r = torch.from_numpy(batch["video"]).clone().requires_grad_(True)
    ro = orc.forward(video_r, text0 * scale_r, batch["video_padding_mask"], batch["text_padding_mask"])
    O.get_loss_init(ro["logits_dual"], ro["logits_joint"], batch["start"], batch["end"],
                    batch["text_padding_mask"])["loss"].backward()
    for got, ref in ((scale.grad, scale_r.grad), (video.grad, video_r.grad)):
        g, r = got.double().reshape(-1), ref.double().reshape(-1)
        cos = float((g @ r) / (g.norm() * r.norm()))
        rel = float((g - r).norm() / r.norm())
        assert cos > 0.999 and rel < 3e-2, (cos, rel)


def test_more_than_64_sentences_takes_the_unfused_similarity_gradient(monkeypatch):
    """N > 64 (BASELINE config 4 has N = 128): tan_sim_grad_gemm keeps two target words per row, so train.py falls
    back to tan_linear_bf16 + tan_sim_grad_tiles; same gradients."""
    from temporalalignnet_b200 import synth
    cpu_ops.install(monkeypatch)
    calls = {"gemm": 0, "tiles": 0}
    from temporalalignnet_b200 import ops
    g0, t0 = ops.sim_grad_gemm, ops.sim_grad_tiles
    monkeypatch.setattr(ops, "sim_grad_gemm", lambda *a, **k: (calls.__setitem__("gemm", calls["gemm"] + 1), g0(*a, **k))[1])
    monkeypatch.setattr(ops, "sim_grad_tiles", lambda *a, **k: (calls.__setitem__("tiles", calls["tiles"] + 1), t0(*a, **k))[1])
    cfg = dict(E=1, D=1, use_text_pos_enc=0)
    sd = synth.make_state_dict(1, 1)
    batch = synth.make_batch(2, 16, 70, seed=11)
    ref_loss, ref_grads = oracle_param_grads(dict(cfg, head=0), sd, batch, _args())
    m = _model(cfg, sd, random_pos_start=0)
    _, res = _step(m, batch, _args())
    assert calls["tiles"] > 0 and calls["gemm"] == 0
    assert abs(res["loss"].item() - ref_loss) < 2e-3 * abs(ref_loss)
    compare_param_grads(m, ref_grads)


def test_compat_shim_classes_train_out_of_the_box(monkeypatch):
    """`from tan_model import TemporalAligner` through temporalalignnet_b200/compat (what train/main.py:20-21 does):
    the exported classes have the training step on, same names / state-dict keys as the base classes."""
    import importlib
    import os
    import sys

    import temporalalignnet_b200
    from temporalalignnet_b200 import get_loss
    cpu_ops.install(monkeypatch)
    compat = os.path.join(os.path.dirname(temporalalignnet_b200.__file__), "compat")
    monkeypatch.syspath_prepend(compat)
    sys.modules.pop("tan_model", None)
    tan_model = importlib.import_module("tan_model")
    try:
        assert tan_model.TemporalAligner.__name__ == "TemporalAligner"
        cfg, sd, batch, _ = case_inputs("g1_e1d1_T32_B4")
        m = tan_model.TemporalAligner(num_encoder_layers=1, num_decoder_layers=1, random_pos_start=0)
        assert set(m.state_dict()) == set(sd)
        m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        video, text = torch.from_numpy(batch["video"]), torch.from_numpy(batch["text"])
        vpm, tpm = torch.from_numpy(batch["video_padding_mask"]), torch.from_numpy(batch["text_padding_mask"])
        out = m(video, text, video_padding_mask=vpm, lang_padding_mask=tpm, text_timestamp=None, abs_text_pos=None)
        loss = get_loss({"start": batch["start"], "end": batch["end"], "text": batch["text_str"]}, video, text,
                        vpm.float(), tpm.float(), out, _args(), None, shard_batch=False)["loss"]
        assert loss.requires_grad                               # train/main.py:112 can call .backward()
        tw = tan_model.TwinTemporalAligner(m=0.99, num_encoder_layers=1, num_decoder_layers=1)
        assert tw.online._autograd_on and not tw.target._autograd_on
        assert all(k.startswith(("online.", "target.")) for k in tw.state_dict())
    finally:
        sys.modules.pop("tan_model", None)


def test_reference_training_loop_body_runs_against_the_drop_in_api(monkeypatch):
    """train/main.py:46-127 restated line by line on synthetic data (ragged token lists -> lang_model -> padded text
    embeddings -> forward -> get_loss with the reference's KEYWORD arguments -> backward -> clip -> AdamW step),
    with the compat classes and a toy text backbone: the loop trains, and the text backbone learns through
    `lang_embed` as in the reference."""
    import importlib
    import os
    import sys
    import types

    from torch.nn.utils.rnn import pad_sequence

    import temporalalignnet_b200
    from temporalalignnet_b200 import synth
    from temporalalignnet_b200.loss import get_loss, get_mask_from_time
    from temporalalignnet_b200.train import clip_gradients
    cpu_ops.install(monkeypatch)
    monkeypatch.syspath_prepend(os.path.join(os.path.dirname(temporalalignnet_b200.__file__), "compat"))
    sys.modules.pop("tan_model", None)
    tan_model = importlib.import_module("tan_model")

    class ToyLang(torch.nn.Module):                       # stands in for Word2VecModel (model/word2vec_model.py:76-102)
        def __init__(self):
            super().__init__()
            self.word_embd = torch.nn.Embedding(50, 64)
            self.fc2 = torch.nn.Linear(64, 512)

        def forward(self, input_ids, attention_mask=None):
            x = self.word_embd(input_ids) * attention_mask[..., None].float()
            return {"pooler_output": self.fc2(x.sum(1) / attention_mask.sum(1, keepdim=True).clamp(min=1).float())}

    def pad_sequence_by_last(sequences):                  # data/loader_htm.py:13-23
        out = sequences[0].new_zeros((len(sequences), max(s.size(0) for s in sequences)) + sequences[0].shape[1:])
        for i, t in enumerate(sequences):
            out[i, :t.size(0)] = t
            out[i, t.size(0):] = t[-1]
        return out

    try:
        torch.manual_seed(0)
        device = "cpu"
        args = types.SimpleNamespace(model="init", sim="cos", learn_agreement=0, temporal_agreement_type="keep",
                                     loss_threshold=0.0, use_alignability_head=0, optim_policy="default", clip_grad=3.0)
        model = tan_model.TemporalAligner(num_encoder_layers=1, num_decoder_layers=1, sim="cos", language_model="word2vec",
                                          pos_enc="learned", use_text_pos_enc=0, use_alignability_head=0,
                                          random_pos_start=0, lang_module=ToyLang())
        model.train()
        optimizer = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-3)
        b = synth.make_batch(4, 32, 4, seed=3)
        g = torch.Generator().manual_seed(1)
        input_data = {"video": torch.from_numpy(b["video"]), "padding_mask": torch.from_numpy(b["video_padding_mask"]).float(),
                      "start": b["start"], "end": b["end"], "text": b["text_str"],
                      "token": [torch.randint(1, 50, (len(s), 6), generator=g) for s in b["start"]]}
        losses = []
        for idx in range(4):
            video_seq = input_data["video"].to(device)
            video_padding_mask = input_data["padding_mask"].to(device)
            num_sentence_per_sample = [i.shape[0] for i in input_data["token"]]                      # :52-55
            flatten_sentence_token = torch.concat([i.to(device) for i in input_data["token"]], 0).long()
            text_embed = model.lang_model(input_ids=flatten_sentence_token,
                                          attention_mask=flatten_sentence_token != 0)["pooler_output"]  # :58-60
            text_embed = pad_sequence_by_last(torch.split(text_embed, num_sentence_per_sample, dim=0))
            text_padding_mask = pad_sequence(torch.split(torch.zeros(flatten_sentence_token.shape[0], device=device),
                                                         num_sentence_per_sample, dim=0), batch_first=True, padding_value=1)
            B, T, _ = video_seq.shape
            N = text_embed.shape[1]
            binary_sentence_timestamp, _, _ = get_mask_from_time(input_data["start"], input_data["end"],
                                                                 num_timestamp=T, num_text=N, device=device)
            logits = model(video_seq, text_embed, video_padding_mask=video_padding_mask.bool(),
                           lang_padding_mask=text_padding_mask.bool(), text_timestamp=binary_sentence_timestamp,
                           abs_text_pos=None)                                                        # :81-87
            loss_dict = get_loss(input_data=input_data, video_seq=video_seq, text_embed=text_embed,
                                 video_padding_mask=video_padding_mask, text_padding_mask=text_padding_mask,
                                 logits=logits, args=args, abs_text_pos=None)                        # :98-105
            loss = loss_dict["loss"]
            assert not torch.isinf(loss) and not torch.isnan(loss)
            loss.backward()                                                                          # :112 (scale 1)
            _ = clip_gradients(model, clip_grad=args.clip_grad)                                      # :115-116
            if idx == 0:
                assert model.bert.fc2.weight.grad is not None and float(model.bert.fc2.weight.grad.abs().sum()) > 0
                assert model.bert.word_embd.weight.grad is not None
            optimizer.step()
            optimizer.zero_grad()
            losses.append({k: v.item() for k, v in loss_dict.items()}["loss"])                        # :127
        assert losses[-1] < losses[0], losses
    finally:
        sys.modules.pop("tan_model", None)
