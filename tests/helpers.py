"""Shared helpers for the parity tests (test infrastructure)."""
import os

import numpy as np
import torch

from temporalalignnet_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# must mirror oracle/make_golden.py:CASES
CASES = {
    "g1_e1d1_T32_B4": dict(E=1, D=1, B=4, T=32, N=4, pad_video_every=0, use_text_pos_enc=0, head=0),
    "g2_e2d3_T24_B3": dict(E=2, D=3, B=3, T=24, N=5, pad_video_every=2, use_text_pos_enc=1, head=1),
    "g3_e6d6_T64_B2": dict(E=6, D=6, B=2, T=64, N=8, pad_video_every=0, use_text_pos_enc=0, head=0),
}


def checksum(a) -> float:
    a = np.asarray(a, dtype=np.float64).ravel()
    return float((a * (1.0 + (np.arange(a.size) % 7))).sum())


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def case_inputs(name):
    """(cfg, state_dict numpy, batch) regenerated from seeds + drift check against the fixture."""
    cfg = CASES[name]
    g = load_golden(name)
    sd = synth.make_state_dict(cfg["E"], cfg["D"], use_alignability_head=bool(cfg["head"]))
    batch = synth.make_batch(cfg["B"], cfg["T"], cfg["N"], pad_video_every=cfg["pad_video_every"])
    assert abs(checksum(batch["video"]) - float(g["in_checksum_video"])) < 1e-6, "synthetic RNG drifted"
    assert abs(checksum(batch["text"]) - float(g["in_checksum_text"])) < 1e-6
    assert abs(sum(checksum(v) for v in sd.values()) - float(g["in_checksum_weights"])) < 1e-6
    return cfg, sd, batch, g


def rel_fro(a, b):
    a = torch.as_tensor(np.asarray(a), dtype=torch.float64)
    b = torch.as_tensor(np.asarray(b), dtype=torch.float64)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_abs(a, b):
    a = torch.as_tensor(np.asarray(a), dtype=torch.float64)
    b = torch.as_tensor(np.asarray(b), dtype=torch.float64)
    return float((a - b).abs().max())


def pack_posbits(mask_bnt: torch.Tensor) -> torch.Tensor:
    """[B, N, T] bool -> packed target bits [B, T, ceil(N/32)] int32 (include/tan_b200.h), torch checker."""
    B, N, T = mask_bnt.shape
    W = (N + 31) // 32
    m = torch.zeros(B, T, W * 32, dtype=torch.int64)
    m[:, :, :N] = mask_bnt.permute(0, 2, 1).to(torch.int64)
    words = (m.view(B, T, W, 32) << torch.arange(32, dtype=torch.int64)).sum(-1)
    words = torch.where(words >= 2 ** 31, words - 2 ** 32, words)
    return words.to(torch.int32)


def unpack_posbits(posbits: torch.Tensor, N: int) -> torch.Tensor:
    """packed [B, T, W] int32 -> [B, N, T] bool."""
    B, T, W = posbits.shape
    w = posbits.to(torch.int64) & 0xFFFFFFFF
    bits = (w[..., None] >> torch.arange(32, dtype=torch.int64)) & 1
    return bits.view(B, T, W * 32)[:, :, :N].permute(0, 2, 1).bool()


def cpu_pos_from_time(start, end, valid, B, T, N):
    """torch restatement of tan_pos_from_time (checker for CPU tests)."""
    tt = torch.arange(T, dtype=torch.float32)[None, None, :]
    m = (start.view(B, N, 1) <= tt) & (tt < end.view(B, N, 1))
    if valid is not None:
        m = m & valid.view(B, N, 1).bool()
    return pack_posbits(m)


def oracle_param_grads(cfg, sd, batch, args, pos_starts=(0, 0, 0)):
    """(loss, {name: d loss / d parameter}) by torch autograd over the fp32 CPU oracle.  The alignability head is
    part of the model only when `args.use_alignability_head` is set.  pos_starts: the three positional-table
    offsets of a `random_pos_start=1` forward (video stack, text, joint stack)."""
    from oracle import tan_oracle as O
    head = int(getattr(args, "use_alignability_head", 0))
    sd_t = {k: torch.from_numpy(v).clone().requires_grad_(True) for k, v in sd.items()
            if head or not k.startswith("binary_head")}
    orc = O.TanOracle(sd_t, cfg["E"], cfg["D"], use_text_pos_enc=cfg["use_text_pos_enc"], use_alignability_head=head)
    orc.sd = sd_t                                                   # keep the leaves (the constructor re-wraps)
    out = orc.forward(batch["video"], batch["text"], batch["video_padding_mask"], batch["text_padding_mask"],
                      pos_starts=pos_starts)
    if head or getattr(args, "learn_agreement", 0) or getattr(args, "loss_threshold", 0.0) > 0:
        res = O.get_loss_full(out, batch["start"], batch["end"], torch.from_numpy(batch["video_padding_mask"]),
                              torch.from_numpy(batch["text_padding_mask"]), args)
    else:
        res = O.get_loss_init(out["logits_dual"], out["logits_joint"], batch["start"], batch["end"],
                              batch["text_padding_mask"])
    res["loss"].backward()
    return float(res["loss"].detach()), {k: v.grad for k, v in sd_t.items()}


def compare_param_grads(model, ref_grads, loose=False):
    """Per parameter: cosine >= 0.999 and rel-Frobenius <= 2e-2 (matrices) / 5e-2 (vectors) against `ref_grads`
    (SURVEY.md 8(c) tolerances of the bf16 path); `loose` doubles them (thresholded recipes)."""
    bad = []
    for name, p in model.named_parameters():
        ref = ref_grads.get(name)
        if ref is None or float(ref.norm()) == 0.0:
            assert p.grad is None or float(p.grad.norm()) == 0.0, name
            continue
        assert p.grad is not None, f"no gradient for {name}"
        g = p.grad.detach().float().cpu().double().reshape(-1)
        r = ref.double().reshape(-1)
        if r.numel() == 1:            # a scalar (binary_head.bias) is a cancelling sum: absolute tolerance
            if abs(float(g) - float(r)) > 2e-3:
                bad.append((name, float(g), float(r)))
            continue
        cos = float((g @ r) / (g.norm() * r.norm()).clamp_min(1e-300))
        rel = float((g - r).norm() / r.norm())
        small = p.dim() == 1
        tol_rel = (5e-2 if small else 2e-2) * (2.0 if loose else 1.0)
        tol_cos = 0.998 if (small or loose) else 0.999
        if not (cos >= tol_cos and rel <= tol_rel):
            bad.append((name, round(cos, 5), round(rel, 4)))
    assert not bad, bad


# must mirror oracle/make_golden.py:BENCH_CASES (reference-generated fixtures at the benchmarked shapes)
BENCH_CASES = {
    "c3": dict(E=6, D=6, B=4, T=256, N=32, pad_video_every=0, use_text_pos_enc=0, head=0, seed=31, kw={}),
    "c5": dict(E=2, D=3, B=4, T=512, N=64, pad_video_every=0, use_text_pos_enc=0, head=1, seed=32,
               kw=dict(learn_agreement=1, loss_threshold=0.5, use_alignability_head=1)),
}


def bench_case_inputs(tag):
    """(cfg, state_dict, batch, loss-args namespace, fixture) of a g_bench.npz case, regenerated from seeds."""
    import types
    c = BENCH_CASES[tag]
    g = load_golden("g_bench")
    sd = synth.make_state_dict(c["E"], c["D"], use_alignability_head=bool(c["head"]), seed=c["seed"])
    batch = synth.make_batch(c["B"], c["T"], c["N"], pad_video_every=c["pad_video_every"], seed=c["seed"],
                             force_full=True)
    chk = checksum(batch["video"]) + checksum(batch["text"]) + sum(checksum(v) for v in sd.values())
    assert abs(chk - float(g[f"{tag}/in_checksum"])) < 1e-6 * abs(chk), "synthetic RNG drifted"
    a = dict(model="init", sim="cos", learn_agreement=0, temporal_agreement_type="keep", loss_threshold=0.0,
             use_alignability_head=0, optim_policy="default")
    a.update(c["kw"])
    return c, sd, batch, types.SimpleNamespace(**a), g


def compare_grads_to_fixture(named_grads, g, tag, tol_norm, tol_cos):
    """Gradients against the reference's per-parameter norm + every-997th-element subsample (g_bench / g_param_grads
    layout).  named_grads: {name: tensor or None}."""
    names = [k[len(tag) + 6:] for k in g if k.startswith(f"{tag}/norm/")]
    assert len(names) >= 37
    bad = []
    for name in names:
        gr = named_grads.get(name)
        assert gr is not None, f"no gradient for {name}"
        got = gr.detach().double().cpu().reshape(-1)
        ref_norm = float(g[f"{tag}/norm/{name}"])
        sub = torch.from_numpy(g[f"{tag}/sub/{name}"]).double()
        if ref_norm == 0.0:
            assert float(got.norm()) == 0.0, name
            continue
        rel = abs(float(got.norm()) - ref_norm) / ref_norm
        gs = got[::997]
        cos = float((gs @ sub) / (gs.norm() * sub.norm()).clamp_min(1e-300)) if sub.numel() >= 64 else 1.0
        if rel > tol_norm or cos < tol_cos:
            bad.append((name, round(rel, 4), round(cos, 5)))
    assert not bad, bad
