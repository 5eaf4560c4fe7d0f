"""Fused optimizer step (tan_optim_adamw_step / tan_ema_update) against torch.optim.AdamW, the reference's
clip_gradients (utils/train_utils.py:3-13) and its EMA update (model/tan_model.py:340-344)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _params(seed=0):
    g = torch.Generator().manual_seed(seed)
    shapes = [(512, 1024), (1536, 512), (1536,), (512,), (1, 512), (1,), (2048, 512), (100, 3), (1024, 512)]
    return [torch.randn(*s, generator=g).to(DEV).requires_grad_(True) for s in shapes]


def _set_grads(ps, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    for p in ps:
        p.grad = (torch.randn(*p.shape, generator=g) * scale).to(DEV)


def _reference_clip(ps, clip):
    """utils/train_utils.py:3-13, literally."""
    norms = []
    for p in ps:
        if p.grad is not None:
            n = p.grad.data.norm(2)
            norms.append(n.item())
            c = clip / (n + 1e-6)
            if c < 1:
                p.grad.data.mul_(c)
    return norms


@pytest.mark.parametrize("steps", [1, 4])
def test_fused_adamw_bit_identical_to_torch_adamw(steps):
    from temporalalignnet_b200.optim import FusedAdamW
    a, b = _params(1), _params(1)
    groups = lambda ps: [{"params": ps[2:6], "weight_decay": 0.0}, {"params": ps[:2] + ps[6:], "weight_decay": 1e-5}]
    ref = torch.optim.AdamW(groups(a), lr=1e-4, weight_decay=1e-5, foreach=True)
    ours = FusedAdamW(groups(b), lr=1e-4, weight_decay=1e-5)
    for s in range(steps):
        _set_grads(a, 10 + s, 0.01)
        _set_grads(b, 10 + s, 0.01)
        for grp in (ref.param_groups, ours.param_groups):        # a moving learning rate, as LambdaLR does
            for g in grp:
                g["lr"] = 1e-4 * (1.0 + 0.5 * s)
        ref.step()
        ours.step()
    torch.cuda.synchronize()
    for x, y in zip(a, b):
        assert torch.equal(x.detach(), y.detach())
    for x, y in zip(a, b):
        assert torch.equal(ref.state[x]["exp_avg"], ours.state[y]["exp_avg"])
        assert torch.equal(ref.state[x]["exp_avg_sq"], ours.state[y]["exp_avg_sq"])


def test_fused_clip_matches_reference_clip_gradients():
    from temporalalignnet_b200.optim import FusedAdamW
    a, b = _params(2), _params(2)
    _set_grads(a, 5, 1.0)
    _set_grads(b, 5, 1.0)
    a[3].grad = None                                              # a parameter without gradient is left alone
    b[3].grad = None
    a3 = a[3].detach().clone()
    norms = _reference_clip(a, 3.0)
    ref = torch.optim.AdamW(a, lr=1e-3, weight_decay=0.01, foreach=True)
    ref.step()
    ours = FusedAdamW(b, lr=1e-3, weight_decay=0.01, clip_grad=3.0)
    ours.step()
    torch.cuda.synchronize()
    got = ours.last_norms.tolist()
    got = got[:3] + got[4:]
    assert max(abs(x - y) / y for x, y in zip(got, norms)) < 1e-5
    assert torch.equal(b[3].detach(), a3)
    for x, y in zip(a, b):                                        # the norm's summation order differs in the last bits
        assert (x.detach() - y.detach()).abs().max().item() <= 2e-6 * max(1.0, x.detach().abs().max().item())


def test_fused_ema_and_momentum_update():
    from temporalalignnet_b200.optim import FusedAdamW, ema_update
    on, tg = _params(3), [p.detach().clone() for p in _params(4)]
    ref = [t * 0.999 + o.detach() * (1.0 - 0.999) for t, o in zip(tg, on)]        # model/tan_model.py:343
    v0 = [t._version for t in tg]
    ema_update(tg, on, 0.999)
    torch.cuda.synchronize()
    for r, t in zip(ref, tg):
        assert torch.equal(r, t)
    assert all(t._version > v for t, v in zip(tg, v0))
    # EMA folded into the optimizer step: target follows the UPDATED online parameters
    _set_grads(on, 6, 0.01)
    tg2 = [t.clone() for t in tg]
    opt = FusedAdamW(on, lr=1e-3, clip_grad=3.0, ema=(tg, on, 0.99))
    opt.step()
    torch.cuda.synchronize()
    for t_old, t_new, o in zip(tg2, tg, on):
        assert torch.equal(t_new, t_old * 0.99 + o.detach() * (1.0 - 0.99))


def test_ema_target_forward_uses_updated_weights():
    """the target's bf16 weight shadows must follow `_momentum_update` / `_copy_param`.
    forward_from_ema after an update equals a fresh model loaded with the EMA weights."""
    from temporalalignnet_b200 import TemporalAligner, TwinTemporalAligner, synth
    sd = synth.make_state_dict(1, 1)
    tw = TwinTemporalAligner(m=0.5, num_encoder_layers=1, num_decoder_layers=1, random_pos_start=0).to(DEV)
    tw.online.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    tw._copy_param()
    batch = synth.make_batch(2, 32, 4)
    video, text = torch.from_numpy(batch["video"]).to(DEV), torch.from_numpy(batch["text"]).to(DEV)
    out0 = tw.forward_from_ema(video, text)["logits_dual"].vfeat.float().clone()      # builds the shadows
    with torch.no_grad():
        for p in tw.online.parameters():
            p.add_(0.05 * torch.randn_like(p))
    tw._momentum_update()
    out1 = tw.forward_from_ema(video, text)["logits_dual"].vfeat.float().clone()
    fresh = TemporalAligner(1, 1, random_pos_start=0).to(DEV)
    fresh.load_state_dict(tw.target.state_dict())
    out2 = fresh(video, text)["logits_dual"].vfeat.float()
    assert not torch.equal(out0, out1)
    assert torch.equal(out1, out2)
