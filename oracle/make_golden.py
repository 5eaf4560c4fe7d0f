"""Generate tests/golden/*.npz by running the UNMODIFIED reference (fp32, CPU) in the build container.

TEST INFRASTRUCTURE.  Run as `python -m oracle.make_golden` from the repo root; needs
/root/reference (oracle/ref_loader.py).  Inputs and weights are NOT stored: they are regenerated
from seeds by `temporalalignnet_b200.synth` (numpy PCG64, order-independent); each fixture
records a float64 checksum of every input so RNG drift is detected instead of silently
mis-comparing.  Outputs are the reference's own tensors.
"""
from __future__ import annotations

import argparse
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_loader import load_reference  # noqa: E402
from temporalalignnet_b200 import synth  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# name -> (model kwargs, batch kwargs)
CASES = {
    # BASELINE.json configs[0]: E1D1 len=32 d=512 batch=4 (the correctness gate)
    "g1_e1d1_T32_B4": dict(E=1, D=1, B=4, T=32, N=4, pad_video_every=0, use_text_pos_enc=0, head=0),
    # ragged text, padded video suffix, text pos-enc, alignability head, deeper stacks
    "g2_e2d3_T24_B3": dict(E=2, D=3, B=3, T=24, N=5, pad_video_every=2, use_text_pos_enc=1, head=1),
    # the paper's stage count at toy size (E6D6), joint length T+N = 72 like config 2
    "g3_e6d6_T64_B2": dict(E=6, D=6, B=2, T=64, N=8, pad_video_every=0, use_text_pos_enc=0, head=0),
}


def checksum(a) -> float:
    a = np.asarray(a, dtype=np.float64).ravel()
    return float((a * (1.0 + (np.arange(a.size) % 7))).sum())


def loss_args(**kw):
    d = dict(model="init", sim="cos", learn_agreement=0, temporal_agreement_type="keep",
             loss_threshold=0.0, use_alignability_head=0, optim_policy="default")
    d.update(kw)
    return types.SimpleNamespace(**d)


def build_reference_model(tan, E, D, use_text_pos_enc, head, seed=888):
    m = tan.TemporalAligner(num_encoder_layers=E, num_decoder_layers=D, sim="cos",
                            language_model="word2vec", pos_enc="learned",
                            use_text_pos_enc=use_text_pos_enc, return_dual_feature=1,
                            random_pos_start=0, use_alignability_head=head)
    sd = synth.make_state_dict(E, D, use_alignability_head=bool(head), seed=seed)
    missing, unexpected = m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    assert not unexpected and all(k.startswith("bert.") for k in missing), (missing, unexpected)
    m.eval()
    return m, sd


def run_case(name, cfg):
    tfm, tan, ref_loss = load_reference()
    m, sd = build_reference_model(tan, cfg["E"], cfg["D"], cfg["use_text_pos_enc"], cfg["head"])
    batch = synth.make_batch(cfg["B"], cfg["T"], cfg["N"], pad_video_every=cfg["pad_video_every"])
    video = torch.from_numpy(batch["video"])
    text = torch.from_numpy(batch["text"])
    vpm = torch.from_numpy(batch["video_padding_mask"])
    tpm = torch.from_numpy(batch["text_padding_mask"])
    out = {"in_checksum_video": checksum(batch["video"]), "in_checksum_text": checksum(batch["text"]),
           "in_checksum_weights": sum(checksum(v) for v in sd.values()),
           "in_start": np.array([x for s in batch["start"] for x in s]),
           "in_end": np.array([x for s in batch["end"] for x in s])}
    sub = 4 if cfg["T"] > 32 else 1              # keep fixtures small: every 4th frame of long clips
    with torch.no_grad():
        res = m(video, text, vpm, tpm, None)
        for k, v in res.items():
            out["fwd_" + k] = v.numpy()[:, :, ::sub] if k == "dual_feature_video" else v.numpy()
        out["feat_subsample"] = np.array(sub)
        out["visual_feature"] = m.get_visual_feature(video, vpm).numpy()[:, :, ::sub]
        t_in = (m.get_textual_feature_with_time(text, None) if cfg["use_text_pos_enc"]
                else m.get_textual_feature(text))
        jv, jt = m.get_joint_feature(video, vpm, t_in, tpm)
        out["joint_video"] = jv.numpy()[:, :, ::sub]
        out["joint_text"] = jt.numpy()
        out["sim_dual_eval"] = m.get_text_visual_sim_dual(video, text).numpy()
        out["sim_joint_eval"] = m.get_text_visual_sim_joint(video, text).numpy()
        # positional-table interpolation path (eval 'global' method, eval_zeroshot_align.py:208-209)
        k_from = max(cfg["T"] // 2, 2)
        out["interp_from"] = np.array(k_from)
        out["sim_dual_eval_interp"] = m.get_text_visual_sim_dual(video, text, k_from).numpy()
        out["sim_joint_eval_interp"] = m.get_text_visual_sim_joint(video, text, k_from).numpy()
    # loss + gradient w.r.t. logits (fp32)
    input_data = {"start": batch["start"], "end": batch["end"], "text": batch["text_str"]}
    ld = res["logits_dual"].clone().requires_grad_(True)
    lj = res["logits_joint"].clone().requires_grad_(True)
    logits = dict(res)
    logits["logits_dual"], logits["logits_joint"] = ld, lj
    loss = ref_loss.get_loss(input_data, video, text, vpm.float(), tpm.float(), logits, loss_args(), None)
    loss["loss"].backward()
    for k, v in loss.items():
        out["loss_" + k] = np.array(float(v))
    out["grad_logits_dual"] = ld.grad.numpy()
    out["grad_logits_joint"] = lj.grad.numpy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **{k: np.asarray(v) for k, v in out.items()})
    print(name, {k: float(v) for k, v in loss.items()})


def block_weights(module, tag, seed=888):
    """Seeded (numpy, order-independent) weights for a bare encoder/decoder stack: matrices
    N(0, 0.05), vectors N(0.5, 0.2).  tests/ regenerate them with the same rule."""
    return {k: torch.from_numpy(synth._normal(f"{tag}.{k}", seed, tuple(v.shape),
                                              0.05 if v.dim() > 1 else 0.2, 0.0 if v.dim() > 1 else 0.5))
            for k, v in module.state_dict().items()}


def run_blocks():
    """Golden vectors for the building blocks at shapes TAN itself never uses: width 768 / 12 heads
    (BASELINE config 4) and the unused TemporalDecoder (model/tfm_model.py:59-103)."""
    tfm, tan, ref_loss = load_reference()
    out = {}
    g = torch.Generator().manual_seed(888)
    for tag, width, heads, layers, L, B in (("enc768", 768, 12, 2, 20, 2), ("enc128", 128, 2, 3, 37, 3)):
        enc = tfm.TemporalEncoder(width, layers, heads).eval()
        enc.load_state_dict(block_weights(enc, tag))
        x = torch.randn(L, B, width, generator=g)
        kpm = torch.zeros(B, L, dtype=torch.bool)
        kpm[0, L - 3:] = True
        with torch.no_grad():
            st = enc(x, kpm)
        out[tag + "_x"] = x.numpy()
        out[tag + "_kpm"] = kpm.numpy()
        out[tag + "_out"] = torch.stack(st).numpy()           # [S, L, B, C]
    dec = tfm.TemporalDecoder(128, 2, 2).eval()
    dec.load_state_dict(block_weights(dec, "dec"))
    x = torch.randn(10, 2, 128, generator=g)
    mem = torch.randn(12, 2, 128, generator=g)
    tk = torch.zeros(2, 10, dtype=torch.bool); tk[1, 8:] = True
    mk = torch.zeros(2, 12, dtype=torch.bool); mk[0, 9:] = True
    with torch.no_grad():
        st = dec(x, mem, tk, mk)
    out.update(dec_x=x.numpy(), dec_mem=mem.numpy(), dec_tk=tk.numpy(), dec_mk=mk.numpy(),
               dec_out=torch.stack(st).numpy())
    # the reference's only executable known-answer (train/loss.py:19-20)
    out["circulant_012"] = ref_loss.circulant(torch.tensor([0, 1, 2]), dim=0).numpy()
    out["sine_pos_16x8"] = tfm.get_position_embedding_sine(8, 16).numpy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "g_blocks.npz"), **out)
    print("g_blocks", out["circulant_012"].tolist())


# loss-only cases (every get_loss branch): name -> (B, S, T, N, pad_video_every, loss-arg overrides, use abs_text_pos)
LOSS_CASES = {
    "init": (4, 3, 32, 4, 0, dict(), False),
    "init_pad": (4, 3, 32, 4, 2, dict(), False),
    "agree_keep": (4, 3, 32, 4, 0, dict(learn_agreement=1), False),
    "agree_i": (3, 3, 24, 5, 0, dict(learn_agreement=1, temporal_agreement_type="i"), False),
    "agree_u": (3, 3, 24, 5, 0, dict(learn_agreement=1, temporal_agreement_type="u"), False),
    "agree_keep_joint": (5, 4, 40, 6, 0, dict(learn_agreement=1, temporal_agreement_type="keep-joint"), False),
    "cotrain_keep_pad": (4, 3, 32, 4, 2, dict(model="cotrain", learn_agreement=1), False),
    "threshold": (6, 3, 64, 8, 0, dict(loss_threshold=0.5), False),
    "head": (4, 3, 32, 4, 0, dict(use_alignability_head=1), True),
    "all_init": (6, 3, 64, 8, 0, dict(learn_agreement=1, loss_threshold=0.5, use_alignability_head=1), True),
    "all_cotrain_bce": (5, 4, 40, 6, 2, dict(model="cotrain", learn_agreement=1, loss_threshold=0.3,
                                            use_alignability_head=1, optim_policy="bce"), True),
    "n40": (3, 3, 48, 40, 0, dict(learn_agreement=1, loss_threshold=0.4, use_alignability_head=1), False),
}


def run_loss_cases():
    """Reference get_loss (train/loss.py:55-422) on synthetic logits for every branch: agreement
    self-labelling, threshold, alignability BCE, init / cotrain."""
    tfm, tan, ref_loss = load_reference()
    out = {}
    for name, (B, S, T, N, pad, kw, use_pos) in LOSS_CASES.items():
        case = synth.make_logit_case(B, S, T, N, tag=name, pad_video_every=pad)
        batch = case["batch"]
        logits = {k: torch.from_numpy(v.copy()) for k, v in case.items() if k not in ("batch", "abs_text_pos")}
        vpm = torch.from_numpy(batch["video_padding_mask"]).float()
        tpm = torch.from_numpy(batch["text_padding_mask"]).float()
        atp = torch.from_numpy(case["abs_text_pos"]) if use_pos else None
        res = ref_loss.get_loss({"start": batch["start"], "end": batch["end"], "text": batch["text_str"]},
                                torch.zeros(B, T, 1), torch.zeros(B, N, 1), vpm, tpm, logits, loss_args(**kw), atp)
        for k, v in res.items():
            out[f"{name}/{k}"] = np.array(float(v))
        out[f"{name}/in_checksum"] = np.array(checksum(case["logits_dual"]) + checksum(case["logits_joint"]))
        print(name, {k: round(float(v), 6) for k, v in res.items()})
    np.savez_compressed(os.path.join(GOLDEN_DIR, "g_loss_full.npz"), **out)


# parameter-gradient fixtures (training step): name -> (model case, loss-arg overrides).  The model carries the
# alignability head only when the loss uses it (g2_head).
GRAD_CASES = {
    "g1": ("g1_e1d1_T32_B4", {}),
    "g2": ("g2_e2d3_T24_B3", {}),
    "g2_thr": ("g2_e2d3_T24_B3", dict(loss_threshold=0.5)),
    "g3": ("g3_e6d6_T64_B2", {}),
    "g2_head": ("g2_e2d3_T24_B3", dict(loss_threshold=0.5, use_alignability_head=1)),
}
GRAD_STRIDE = 997        # every 997th element of each parameter gradient is stored (+ its norm)


def run_param_grads():
    """d loss / d parameter of the UNMODIFIED reference (forward + get_loss + autograd, fp32 CPU): per parameter
    the gradient norm and a strided subsample -> tests/golden/g_param_grads.npz.  Pins the oracle's autograd
    (tests/test_oracle.py), which in turn is what the GPU backward pass is checked against."""
    tfm, tan, ref_loss = load_reference()
    out = {}
    for tag, (case, kw) in GRAD_CASES.items():
        cfg = CASES[case]
        m, sd = build_reference_model(tan, cfg["E"], cfg["D"], cfg["use_text_pos_enc"], int(kw.get("use_alignability_head", 0)))
        m.train()
        batch = synth.make_batch(cfg["B"], cfg["T"], cfg["N"], pad_video_every=cfg["pad_video_every"])
        video, text = torch.from_numpy(batch["video"]), torch.from_numpy(batch["text"])
        vpm, tpm = torch.from_numpy(batch["video_padding_mask"]), torch.from_numpy(batch["text_padding_mask"])
        res = m(video, text, vpm, tpm, None)
        input_data = {"start": batch["start"], "end": batch["end"], "text": batch["text_str"]}
        loss = ref_loss.get_loss(input_data, video, text, vpm.float(), tpm.float(), res, loss_args(**kw), None)
        loss["loss"].backward()
        out[f"{tag}/loss"] = np.array(float(loss["loss"]))
        n = 0
        for name, p in m.named_parameters():
            if name.startswith("bert.") or p.grad is None:
                continue
            g = p.grad.detach().double().reshape(-1)
            out[f"{tag}/norm/{name}"] = np.array(float(g.norm()))
            out[f"{tag}/sub/{name}"] = g[::GRAD_STRIDE].float().numpy()
            n += 1
        print("param grads", tag, float(loss["loss"]), n, "parameters")
    np.savez_compressed(os.path.join(GOLDEN_DIR, "g_param_grads.npz"), **out)


# Reference-generated fixtures at the BENCHMARKED shapes (parity must not stop at T = 160):
#   c3: BASELINE configs[2]'s per-clip shape (E6D6, T=256, N=32) at 4 clips, `--model init` loss
#   c5: BASELINE configs[4]'s loss recipe (learn_agreement + loss_threshold + alignability head, train/loss.py:88-357)
#       at T=512, N=64 with a joint stack deep enough for the head (D=3)
BENCH_CASES = {
    "c3": dict(E=6, D=6, B=4, T=256, N=32, pad_video_every=0, use_text_pos_enc=0, head=0, seed=31, kw={}),
    "c5": dict(E=2, D=3, B=4, T=512, N=64, pad_video_every=0, use_text_pos_enc=0, head=1, seed=32,
               kw=dict(learn_agreement=1, loss_threshold=0.5, use_alignability_head=1)),
}


def run_bench_cases():
    """Forward + get_loss + autograd of the UNMODIFIED reference at the benchmarked shapes -> g_bench.npz: the loss
    dict, a strided subsample of the logits, and per parameter the gradient norm + a strided subsample."""
    tfm, tan, ref_loss = load_reference()
    out = {}
    for tag, c in BENCH_CASES.items():
        m, sd = build_reference_model(tan, c["E"], c["D"], c["use_text_pos_enc"], c["head"], seed=c["seed"])
        m.train()
        batch = synth.make_batch(c["B"], c["T"], c["N"], pad_video_every=c["pad_video_every"], seed=c["seed"],
                                 force_full=True)
        video, text = torch.from_numpy(batch["video"]), torch.from_numpy(batch["text"])
        vpm, tpm = torch.from_numpy(batch["video_padding_mask"]), torch.from_numpy(batch["text_padding_mask"])
        res = m(video, text, vpm, tpm, None)
        input_data = {"start": batch["start"], "end": batch["end"], "text": batch["text_str"]}
        # get_loss fills the logits in place under learn_agreement (train/loss.py:96-100): keep pristine copies
        for k in ("logits_dual", "logits_joint"):
            out[f"{tag}/{k}_sub"] = res[k].detach()[:, :, ::16].numpy().copy()
        loss = ref_loss.get_loss(input_data, video, text, vpm.float(), tpm.float(), res, loss_args(**c["kw"]), None)
        loss["loss"].backward()
        for k, v in loss.items():
            out[f"{tag}/loss/{k}"] = np.array(float(v))
        out[f"{tag}/in_checksum"] = np.array(checksum(batch["video"]) + checksum(batch["text"]) +
                                             sum(checksum(v) for v in sd.values()))
        for name, p in m.named_parameters():
            if name.startswith("bert.") or p.grad is None:
                continue
            g = p.grad.detach().double().reshape(-1)
            out[f"{tag}/norm/{name}"] = np.array(float(g.norm()))
            out[f"{tag}/sub/{name}"] = g[::GRAD_STRIDE].float().numpy()
        print("bench case", tag, {k: round(float(v), 6) for k, v in loss.items()})
    np.savez_compressed(os.path.join(GOLDEN_DIR, "g_bench.npz"), **out)


def run_align():
    """get_alignability (model/tan_model.py:284-312) of the reference for the head-carrying case g2, without and with
    positional interpolation (tuple form: video table, text table) -> g_align.npz."""
    tfm, tan, ref_loss = load_reference()
    cfg = CASES["g2_e2d3_T24_B3"]
    m, sd = build_reference_model(tan, cfg["E"], cfg["D"], cfg["use_text_pos_enc"], 1)
    batch = synth.make_batch(cfg["B"], cfg["T"], cfg["N"], pad_video_every=cfg["pad_video_every"])
    video, text = torch.from_numpy(batch["video"]), torch.from_numpy(batch["text"])
    out = {}
    with torch.no_grad():
        a = m.get_alignability(video, text)
        b = m.get_alignability(video, text, (12, 3))
    for k in a:
        out[k] = a[k].numpy()
        out[k + "/interp_12_3"] = b[k].numpy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "g_align.npz"), **out)
    print("g_align", {k: v.shape for k, v in out.items()})


def make_word2vec_case(V=500, S=7, seed=888):
    """Seeded weights / tokens of a small Word2VecModel (vocabulary V instead of 66250; same layer sizes): sentence 2
    consists of stop words only (all ids 0 -> attention mask all zero, model/word2vec_model.py:93), sentence 3 repeats
    a word (tie of the max-pool)."""
    sd = {"word_embd.weight": synth._normal("w2v.word_embd", seed, (V, 300), 0.3),
          "fc1.weight": synth._normal("w2v.fc1.weight", seed, (2048, 300), 0.05),
          "fc1.bias": synth._normal("w2v.fc1.bias", seed, (2048,), 0.1),
          "fc2.weight": synth._normal("w2v.fc2.weight", seed, (512, 2048), 0.03),
          "fc2.bias": synth._normal("w2v.fc2.bias", seed, (512,), 0.1)}
    r = synth._rng("w2v.tokens", seed)
    ids = np.zeros((S, 32), np.int64)
    for s_ in range(S):
        n = int(r.integers(1, 33))
        ids[s_, :n] = r.integers(1, V, size=n)
    ids[2, :] = 0
    ids[3, :4] = ids[3, 0]
    return sd, ids


def run_word2vec():
    """Word2VecModel.forward of the reference (model/word2vec_model.py:83-101) on synthetic weights: the class needs
    the S3D checkpoint to CONSTRUCT (:79), so an instance is made without __init__ and given its three sub-modules."""
    import importlib.util
    # the real module file (load_reference() registers a STUB under the name `word2vec_model`); its only missing
    # import is the S3D class of the MIL-NCE checkpoint code (model/word2vec_model.py:8), unused by forward()
    sys.modules.setdefault("s3d_milnce", types.ModuleType("s3d_milnce"))
    s3dg = types.ModuleType("s3d_milnce.s3dg")
    s3dg.S3D = object
    sys.modules.setdefault("s3d_milnce.s3dg", s3dg)
    from oracle.ref_loader import REF_ROOT
    spec = importlib.util.spec_from_file_location("ref_word2vec_model", os.path.join(REF_ROOT, "model", "word2vec_model.py"))
    w2v = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(w2v)
    sd, ids = make_word2vec_case()
    m = w2v.Word2VecModel.__new__(w2v.Word2VecModel)
    torch.nn.Module.__init__(m)
    m.word_embd = torch.nn.Embedding(sd["word_embd.weight"].shape[0], 300)
    m.fc1 = torch.nn.Linear(300, 2048)
    m.fc2 = torch.nn.Linear(2048, 512)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    tok = torch.from_numpy(ids)
    out = m(input_ids=tok, attention_mask=(tok != 0))
    pooled = out["pooler_output"]
    g_out = torch.from_numpy(synth._normal("w2v.gout", 888, tuple(pooled.shape), 1.0))
    (pooled * g_out).sum().backward()
    res = {"pooler_output": pooled.detach().numpy(), "last_hidden_state": out["last_hidden_state"].detach().numpy()[:, ::8],
           "in_checksum": np.array(sum(checksum(v) for v in sd.values()) + checksum(ids))}
    for k, p in m.named_parameters():          # norm + every 97th element (word_embd is frozen: no gradient, :84-85)
        if p.grad is None:
            res["grad_none/" + k] = np.array(1)
            continue
        gflat = p.grad.detach().double().reshape(-1)
        res["grad_norm/" + k] = np.array(float(gflat.norm()))
        res["grad_sub/" + k] = gflat[::97].float().numpy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "g_word2vec.npz"), **res)
    print("g_word2vec", pooled.shape, {k: float(v) for k, v in res.items() if k.startswith("grad_norm/")},
          [k for k in res if k.startswith("grad_none/")])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.manual_seed(888)
    np.random.seed(888)
    torch.set_grad_enabled(True)
    for name, cfg in CASES.items():
        if a.only and a.only != name:
            continue
        run_case(name, cfg)
    if not a.only or a.only == "g_blocks":
        run_blocks()
    if not a.only or a.only == "g_loss_full":
        run_loss_cases()
    if not a.only or a.only == "g_param_grads":
        run_param_grads()
    if not a.only or a.only == "g_bench":
        run_bench_cases()
    if not a.only or a.only == "g_align":
        run_align()
    if not a.only or a.only == "g_word2vec":
        run_word2vec()


if __name__ == "__main__":
    main()
