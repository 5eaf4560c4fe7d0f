"""Training step (GPU): `loss.backward()` through the hand-written backward pass against torch autograd over the
fp32 CPU oracle (itself pinned to the reference's parameter gradients, tests/test_oracle.py).
Tolerances: per parameter cosine similarity >= 0.999 and rel-Frobenius <= 2e-2 for the matrices;
small vectors (biases, LayerNorm affine) whose gradient is a sum of bf16-rounded terms: rel-Frobenius <= 5e-2."""
import types

import numpy as np
import pytest
import torch

from tests.helpers import case_inputs, compare_param_grads, oracle_param_grads

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _args(**kw):
    d = dict(model="init", sim="cos", learn_agreement=0, temporal_agreement_type="keep", loss_threshold=0.0,
             use_alignability_head=0, optim_policy="default")
    d.update(kw)
    return types.SimpleNamespace(**d)


def _build(cfg, sd, head=0):
    from temporalalignnet_b200 import TemporalAligner
    m = TemporalAligner(num_encoder_layers=cfg["E"], num_decoder_layers=cfg["D"], sim="cos", language_model="word2vec",
                        pos_enc="learned", use_text_pos_enc=cfg["use_text_pos_enc"], return_dual_feature=1,
                        random_pos_start=0, use_alignability_head=head)
    sd = {k: torch.from_numpy(v) for k, v in sd.items() if head or not k.startswith("binary_head")}
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return m.to(DEV)


def _oracle_grads(cfg, sd, batch, args):
    return oracle_param_grads(cfg, sd, batch, args)


def _compare(model, ref_grads, loose=False):
    compare_param_grads(model, ref_grads, loose)


@pytest.mark.parametrize("name,head_off", [("g1_e1d1_T32_B4", False), ("g2_e2d3_T24_B3", True), ("g3_e6d6_T64_B2", False)])
def test_backward_vs_oracle_autograd(name, head_off):
    from temporalalignnet_b200 import get_loss
    cfg, sd, batch, _ = case_inputs(name)
    cfg = dict(cfg, head=0)
    args = _args()
    ref_loss, ref_grads = _oracle_grads(cfg, sd, batch, args)
    m = _build(cfg, sd)
    m.train()
    m.enable_autograd(True)
    video, text = torch.from_numpy(batch["video"]).to(DEV), torch.from_numpy(batch["text"]).to(DEV)
    vpm = torch.from_numpy(batch["video_padding_mask"]).to(DEV)
    tpm = torch.from_numpy(batch["text_padding_mask"]).to(DEV)
    out = m(video, text, video_padding_mask=vpm, lang_padding_mask=tpm, text_timestamp=None, abs_text_pos=None)
    input_data = {"start": batch["start"], "end": batch["end"], "text": batch["text_str"]}
    res = get_loss(input_data, video, text, vpm.float(), tpm.float(), out, args, None)
    loss = res["loss"]
    assert loss.requires_grad
    assert abs(loss.item() - ref_loss) < 1e-3 * abs(ref_loss), (loss.item(), ref_loss)
    (loss * 4.0).backward()                      # a GradScaler-style scale flows through grad_out
    for p in m.parameters():
        if p.grad is not None:
            p.grad.div_(4.0)
    torch.cuda.synchronize()
    _compare(m, ref_grads)
    # the unused `mlp` Linear gets no gradient, as in the reference
    assert m.mlp.weight.grad is None


def test_training_forward_equals_inference_forward():
    """The taped forward (unfused kernel sequence) and the inference forward agree on the loss (1e-3)."""
    from temporalalignnet_b200 import get_loss
    cfg, sd, batch, _ = case_inputs("g3_e6d6_T64_B2")
    m = _build(cfg, sd)
    video, text = torch.from_numpy(batch["video"]).to(DEV), torch.from_numpy(batch["text"]).to(DEV)
    vpm = torch.from_numpy(batch["video_padding_mask"]).to(DEV)
    tpm = torch.from_numpy(batch["text_padding_mask"]).to(DEV)
    input_data = {"start": batch["start"], "end": batch["end"], "text": batch["text_str"]}
    l0 = get_loss(input_data, video, text, vpm.float(), tpm.float(), m(video, text, video_padding_mask=vpm,
                  lang_padding_mask=tpm), _args(), None)["loss"]
    assert not l0.requires_grad
    m.enable_autograd(True)
    l1 = get_loss(input_data, video, text, vpm.float(), tpm.float(), m(video, text, video_padding_mask=vpm,
                  lang_padding_mask=tpm), _args(), None)["loss"]
    assert l1.requires_grad
    assert abs(l0.item() - l1.item()) < 1e-3 * abs(l0.item())
    with torch.no_grad():                        # no_grad / eval: the inference path, no tape
        out = m(video, text, video_padding_mask=vpm, lang_padding_mask=tpm)
    assert getattr(out["logits_dual"], "tape", None) is None


def test_backward_with_threshold_vs_oracle_autograd():
    """loss_threshold keeps a subset of sentences / frames (train/loss.py:277-304): the selections reach the
    gradient through the row / column coefficients."""
    from temporalalignnet_b200 import get_loss
    cfg, sd, batch, _ = case_inputs("g2_e2d3_T24_B3")
    cfg = dict(cfg, head=0)
    args = _args(loss_threshold=0.5)
    ref_loss, ref_grads = _oracle_grads(cfg, sd, batch, args)
    m = _build(cfg, sd)
    m.enable_autograd(True)
    video, text = torch.from_numpy(batch["video"]).to(DEV), torch.from_numpy(batch["text"]).to(DEV)
    vpm = torch.from_numpy(batch["video_padding_mask"]).to(DEV)
    tpm = torch.from_numpy(batch["text_padding_mask"]).to(DEV)
    out = m(video, text, video_padding_mask=vpm, lang_padding_mask=tpm)
    input_data = {"start": batch["start"], "end": batch["end"], "text": batch["text_str"]}
    res = get_loss(input_data, video, text, vpm.float(), tpm.float(), out, args, None)
    assert abs(res["loss"].item() - ref_loss) < 2e-3 * abs(ref_loss), (res["loss"].item(), ref_loss)
    res["loss"].backward()
    torch.cuda.synchronize()
    _compare(m, ref_grads, loose=True)


def test_sgd_step_reduces_loss():
    """A few optimizer steps on one batch lower the loss: gradients point downhill end to end."""
    from temporalalignnet_b200 import get_loss
    cfg, sd, batch, _ = case_inputs("g1_e1d1_T32_B4")
    m = _build(cfg, sd)
    m.enable_autograd(True)
    opt = torch.optim.SGD(m.parameters(), lr=0.05)
    video, text = torch.from_numpy(batch["video"]).to(DEV), torch.from_numpy(batch["text"]).to(DEV)
    vpm = torch.from_numpy(batch["video_padding_mask"]).to(DEV)
    tpm = torch.from_numpy(batch["text_padding_mask"]).to(DEV)
    input_data = {"start": batch["start"], "end": batch["end"], "text": batch["text_str"]}
    losses = []
    for _ in range(5):
        out = m(video, text, video_padding_mask=vpm, lang_padding_mask=tpm)
        loss = get_loss(input_data, video, text, vpm.float(), tpm.float(), out, _args(), None)["loss"]
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(np.isfinite(losses)), losses
    assert losses[-1] < losses[0], losses


def test_wgrad_split_contraction_deterministic():
    """Weight gradient with the contraction split over the CTA pairs inside one launch (train._wgrad ->
    tan_gemm_tn_bf16) == dy^T @ x, bias gradient == column sums, bit-identical on a re-run."""
    from temporalalignnet_b200 import train
    g = torch.Generator().manual_seed(3)
    M, N, K = 12288 + 100, 512, 1024
    dy = (torch.randn(M, N, generator=g) * 0.1).to(DEV).to(torch.bfloat16)
    x = torch.randn(M, K, generator=g).to(DEV).to(torch.bfloat16)
    gw = torch.ones(N, K, dtype=torch.float32, device=DEV)
    gb = torch.zeros(N, dtype=torch.float32, device=DEV)
    train._wgrad(dy, x, gw, gb)
    torch.cuda.synchronize()
    ref = 1.0 + dy.float().t() @ x.float()
    assert ((gw - ref).norm() / ref.norm()).item() < 1e-4
    assert (gb - dy.float().sum(0)).abs().max().item() < 1e-2
    gw2 = torch.ones(N, K, dtype=torch.float32, device=DEV)
    train._wgrad(dy, x, gw2)
    torch.cuda.synchronize()
    assert torch.equal(gw, gw2)                               # fixed-order partial sums: deterministic
