"""Pins the CPU oracle (oracle/tan_oracle.py) against outputs of the UNMODIFIED reference:
the committed fixtures (tests/golden, made by oracle/make_golden.py) always, and the live
reference modules when /root/reference exists (build container only)."""
import numpy as np
import pytest
import torch

from oracle import tan_oracle as O
from oracle.ref_loader import load_reference, reference_available
from temporalalignnet_b200 import synth
from tests.helpers import CASES, case_inputs, checksum, load_golden, max_abs

FP32_TOL = 2e-5   # fp32 reassociation noise between two CPU evaluations of the same math


def _run_oracle(cfg, sd, batch):
    orc = O.TanOracle(sd, cfg["E"], cfg["D"], use_text_pos_enc=cfg["use_text_pos_enc"],
                      use_alignability_head=cfg["head"])
    video = torch.from_numpy(batch["video"])
    text = torch.from_numpy(batch["text"])
    out = orc.forward(video, text, batch["video_padding_mask"], batch["text_padding_mask"])
    return orc, out


@pytest.mark.parametrize("name", list(CASES))
def test_forward_matches_reference_fixture(name):
    cfg, sd, batch, g = case_inputs(name)
    orc, out = _run_oracle(cfg, sd, batch)
    sub = int(g["feat_subsample"])
    assert max_abs(out["logits_dual"], g["fwd_logits_dual"]) < FP32_TOL
    assert max_abs(out["logits_joint"], g["fwd_logits_joint"]) < FP32_TOL
    assert max_abs(out["dual_feature_video"][:, :, ::sub], g["fwd_dual_feature_video"]) < FP32_TOL
    assert max_abs(out["dual_feature_text"], g["fwd_dual_feature_text"]) < FP32_TOL
    if cfg["head"]:
        assert max_abs(out["dual_logits_alignability"], g["fwd_dual_logits_alignability"]) < FP32_TOL
        assert max_abs(out["joint_logits_alignability"], g["fwd_joint_logits_alignability"]) < 1e-4
    video = torch.from_numpy(batch["video"])
    text = torch.from_numpy(batch["text"])
    vpm = torch.from_numpy(batch["video_padding_mask"])
    vf = orc.get_visual_feature(video, vpm)
    assert max_abs(vf[:, :, ::sub], g["visual_feature"]) < 2e-4   # un-normalised features, |x| ~ 1-10


@pytest.mark.parametrize("name", list(CASES))
def test_eval_sims_match_reference_fixture(name):
    cfg, sd, batch, g = case_inputs(name)
    orc = O.TanOracle(sd, cfg["E"], cfg["D"], use_text_pos_enc=cfg["use_text_pos_enc"],
                      use_alignability_head=cfg["head"])
    video = torch.from_numpy(batch["video"])
    text = torch.from_numpy(batch["text"])
    k = int(g["interp_from"])
    assert max_abs(orc.get_text_visual_sim_dual(video, text), g["sim_dual_eval"]) < FP32_TOL
    assert max_abs(orc.get_text_visual_sim_joint(video, text), g["sim_joint_eval"]) < FP32_TOL
    assert max_abs(orc.get_text_visual_sim_dual(video, text, k), g["sim_dual_eval_interp"]) < FP32_TOL
    assert max_abs(orc.get_text_visual_sim_joint(video, text, k), g["sim_joint_eval_interp"]) < FP32_TOL


@pytest.mark.parametrize("name", list(CASES))
def test_loss_matches_reference_fixture(name):
    cfg, sd, batch, g = case_inputs(name)
    ld = torch.from_numpy(g["fwd_logits_dual"]).requires_grad_(True)
    lj = torch.from_numpy(g["fwd_logits_joint"]).requires_grad_(True)
    loss = O.get_loss_init(ld, lj, batch["start"], batch["end"], batch["text_padding_mask"])
    for k in ("loss", "loss-dual", "loss-joint"):
        assert abs(float(loss[k]) - float(g["loss_" + k])) < 1e-5 * abs(float(g["loss_" + k])), k
    loss["loss"].backward()
    assert max_abs(ld.grad, g["grad_logits_dual"]) < 1e-6
    assert max_abs(lj.grad, g["grad_logits_joint"]) < 1e-6


def _block_sd(prefix_tag, keys_shapes, seed=888):
    return {k: synth._normal(f"{prefix_tag}.{k}", seed, shape, 0.05 if len(shape) > 1 else 0.2,
                             0.0 if len(shape) > 1 else 0.5) for k, shape in keys_shapes.items()}


def _enc_shapes(width, layers, dec=False):
    d = {}
    for i in range(layers):
        p = f"resblocks.{i}."
        names = ["attn"] + (["self_attn"] if dec else [])
        for a in names:
            d[p + a + ".in_proj_weight"] = (3 * width, width)
            d[p + a + ".in_proj_bias"] = (3 * width,)
            d[p + a + ".out_proj.weight"] = (width, width)
            d[p + a + ".out_proj.bias"] = (width,)
        for ln in ["ln_1", "ln_2"] + (["ln_3"] if dec else []):
            d[p + ln + ".weight"] = (width,)
            d[p + ln + ".bias"] = (width,)
        d[p + "mlp.c_fc.weight"] = (4 * width, width)
        d[p + "mlp.c_fc.bias"] = (4 * width,)
        d[p + "mlp.c_proj.weight"] = (width, 4 * width)
        d[p + "mlp.c_proj.bias"] = (width,)
    return d


@pytest.mark.parametrize("tag,width,heads,layers", [("enc768", 768, 12, 2), ("enc128", 128, 2, 3)])
def test_encoder_blocks_match_reference_fixture(tag, width, heads, layers):
    g = load_golden("g_blocks")
    sd = {"e." + k: torch.from_numpy(v) for k, v in _block_sd(tag, _enc_shapes(width, layers)).items()}
    x = torch.from_numpy(g[tag + "_x"]).transpose(0, 1)          # [L,B,C] -> [B,L,C]
    kpm = torch.from_numpy(g[tag + "_kpm"])
    st = O.encoder_stack(x, kpm, sd, "e", layers, heads)
    got = torch.stack(st).permute(0, 2, 1, 3)                     # [S,B,L,C] -> [S,L,B,C]
    assert max_abs(got, g[tag + "_out"]) < 5e-4 * float(np.abs(g[tag + "_out"]).max())


def test_decoder_blocks_match_reference_fixture():
    g = load_golden("g_blocks")
    sd = {"d." + k: torch.from_numpy(v) for k, v in _block_sd("dec", _enc_shapes(128, 2, dec=True)).items()}
    x = torch.from_numpy(g["dec_x"]).transpose(0, 1)
    mem = torch.from_numpy(g["dec_mem"]).transpose(0, 1)
    st = O.decoder_stack(x, mem, torch.from_numpy(g["dec_tk"]), torch.from_numpy(g["dec_mk"]), sd, "d", 2, 2)
    got = torch.stack(st).permute(0, 2, 1, 3)
    assert max_abs(got, g["dec_out"]) < 5e-4 * float(np.abs(g["dec_out"]).max())


def test_reference_known_answer_circulant():
    """train/loss.py:19-20 -- the reference's only executable golden vector."""
    g = load_golden("g_blocks")
    assert g["circulant_012"].tolist() == [[0, 1, 2], [2, 0, 1], [1, 2, 0]]


@pytest.mark.skipif(not reference_available(), reason="/root/reference only exists in the build container")
def test_oracle_vs_live_reference_random_pos_start():
    """Replays the reference's three np.random.randint draws (model/tan_model.py:163,:224,:195)."""
    from oracle.make_golden import build_reference_model
    from oracle.ref_loader import load_reference
    _, tan, _ = load_reference()
    cfg = dict(E=2, D=2, use_text_pos_enc=1, head=0)
    m, sd = build_reference_model(tan, cfg["E"], cfg["D"], cfg["use_text_pos_enc"], cfg["head"])
    m.random_pos_start = 1
    batch = synth.make_batch(2, 16, 4, seed=5)
    video, text = torch.from_numpy(batch["video"]), torch.from_numpy(batch["text"])
    vpm, tpm = torch.from_numpy(batch["video_padding_mask"]), torch.from_numpy(batch["text_padding_mask"])
    np.random.seed(123)
    with torch.no_grad():
        ref = m(video, text, vpm, tpm, None)
    np.random.seed(123)
    draws = (np.random.randint(0, 8), np.random.randint(0, 2), np.random.randint(0, 8))
    orc = O.TanOracle(sd, 2, 2, use_text_pos_enc=1)
    out = orc.forward(video, text, vpm, tpm, pos_starts=draws)
    assert max_abs(out["logits_dual"], ref["logits_dual"]) < FP32_TOL
    assert max_abs(out["logits_joint"], ref["logits_joint"]) < FP32_TOL


# ------------------------------------------------------------------------------------------------
# every get_loss branch (agreement self-labelling, threshold, alignability BCE; train/loss.py:88-373)
# ------------------------------------------------------------------------------------------------
from oracle.make_golden import LOSS_CASES, loss_args  # noqa: E402


def _loss_case(name):
    B, S, T, N, pad, kw, use_pos = LOSS_CASES[name]
    case = synth.make_logit_case(B, S, T, N, tag=name, pad_video_every=pad)
    batch = case["batch"]
    logits = {k: torch.from_numpy(v.copy()) for k, v in case.items() if k not in ("batch", "abs_text_pos")}
    atp = torch.from_numpy(case["abs_text_pos"]) if use_pos else None
    return case, batch, logits, atp, loss_args(**kw)


@pytest.mark.parametrize("name", list(LOSS_CASES))
def test_full_loss_oracle_matches_reference_fixture(name):
    gold = load_golden("g_loss_full")
    case, batch, logits, atp, args = _loss_case(name)
    assert abs(checksum(case["logits_dual"]) + checksum(case["logits_joint"]) - float(gold[f"{name}/in_checksum"])) < 1e-6
    res = O.get_loss_full(logits, batch["start"], batch["end"], batch["video_padding_mask"],
                          batch["text_padding_mask"], args, atp)
    keys = {k.split("/", 1)[1] for k in gold if k.startswith(name + "/") and not k.endswith("in_checksum")}
    assert set(res) == keys
    for k in keys:
        ref, got = float(gold[f"{name}/{k}"]), float(res[k])
        assert abs(got - ref) <= 1e-5 * max(abs(ref), 1e-3), (name, k, got, ref)


@pytest.mark.skipif(not reference_available(), reason="needs /root/reference")
@pytest.mark.parametrize("name", ["agree_keep", "all_init", "all_cotrain_bce"])
def test_full_loss_oracle_matches_live_reference(name):
    _, _, ref_loss = load_reference()
    case, batch, logits, atp, args = _loss_case(name)
    B, S, T, N = LOSS_CASES[name][:4]
    ref = ref_loss.get_loss({"start": batch["start"], "end": batch["end"], "text": batch["text_str"]},
                            torch.zeros(B, T, 1), torch.zeros(B, N, 1),
                            torch.from_numpy(batch["video_padding_mask"]).float(),
                            torch.from_numpy(batch["text_padding_mask"]).float(),
                            {k: v.clone() for k, v in logits.items()}, args, atp)
    res = O.get_loss_full(logits, batch["start"], batch["end"], batch["video_padding_mask"],
                          batch["text_padding_mask"], args, atp)
    assert set(res) == set(ref)
    for k in ref:
        assert abs(float(res[k]) - float(ref[k])) <= 1e-5 * max(abs(float(ref[k])), 1e-3), (k, float(res[k]), float(ref[k]))


# ------------------------------------------------------------------------------------------------
# training step: the oracle's parameter gradients (torch autograd over the restatement) against the reference's
# ------------------------------------------------------------------------------------------------
GRAD_CASES = {"g1": ("g1_e1d1_T32_B4", {}), "g2": ("g2_e2d3_T24_B3", {}),
              "g2_thr": ("g2_e2d3_T24_B3", dict(loss_threshold=0.5)), "g3": ("g3_e6d6_T64_B2", {}),
              "g2_head": ("g2_e2d3_T24_B3", dict(loss_threshold=0.5, use_alignability_head=1))}


@pytest.mark.parametrize("tag", list(GRAD_CASES))
def test_oracle_param_grads_vs_reference_fixture(tag):
    """tests/golden/g_param_grads.npz holds, per parameter, the norm and every 997th element of the UNMODIFIED
    reference's d loss / d parameter (oracle/make_golden.py:run_param_grads)."""
    import types
    from tests.helpers import case_inputs, oracle_param_grads
    g = load_golden("g_param_grads")
    case, kw = GRAD_CASES[tag]
    cfg, sd, batch, _ = case_inputs(case)
    a = dict(model="init", sim="cos", learn_agreement=0, temporal_agreement_type="keep", loss_threshold=0.0,
             use_alignability_head=0, optim_policy="default")
    a.update(kw)
    loss, grads = oracle_param_grads(cfg, sd, batch, types.SimpleNamespace(**a))
    assert abs(loss - float(g[f"{tag}/loss"])) < 1e-5 * abs(loss)
    names = [k[len(tag) + 6:] for k in g if k.startswith(f"{tag}/norm/")]
    assert len(names) >= 37
    for name in names:
        got = grads[name].double().reshape(-1)
        ref_norm = float(g[f"{tag}/norm/{name}"])
        assert abs(float(got.norm()) - ref_norm) <= 2e-4 * ref_norm + 5e-7, name      # (floor: cancelling fp32 sums)
        sub = torch.from_numpy(g[f"{tag}/sub/{name}"]).double()
        assert float((got[::997] - sub).abs().max()) <= 2e-4 * max(float(sub.abs().max()), ref_norm * 1e-2) + 5e-7, name
    # parameters the reference leaves without gradient get none / zero here as well
    for name, gr in grads.items():
        if name not in names:
            assert gr is None or float(gr.abs().max()) == 0.0, name


@pytest.mark.parametrize("tag", ["c3", "c5"])
def test_oracle_at_benchmarked_shapes_vs_reference_fixture(tag):
    """g_bench.npz: forward + get_loss + autograd of the UNMODIFIED reference at BASELINE configs[2]'s per-clip shape
    (E6D6, T=256, N=32) and with configs[4]'s loss recipe at T=512, N=64 (oracle/make_golden.py:run_bench_cases)."""
    from tests.helpers import bench_case_inputs, compare_grads_to_fixture, oracle_param_grads
    c, sd, batch, args, g = bench_case_inputs(tag)
    orc = O.TanOracle(sd, c["E"], c["D"], use_alignability_head=c["head"])
    with torch.no_grad():
        out = orc.forward(torch.from_numpy(batch["video"]), torch.from_numpy(batch["text"]),
                          batch["video_padding_mask"], batch["text_padding_mask"])
    for k in ("logits_dual", "logits_joint"):
        assert max_abs(out[k][:, :, ::16], g[f"{tag}/{k}_sub"]) < FP32_TOL
    loss, grads = oracle_param_grads(c, sd, batch, args)
    assert abs(loss - float(g[f"{tag}/loss/loss"])) < 2e-5 * abs(loss)
    compare_grads_to_fixture(grads, g, tag, tol_norm=5e-4, tol_cos=0.99999)


def test_oracle_get_alignability_vs_reference_fixture():
    g = load_golden("g_align")
    cfg, sd, batch, _ = case_inputs("g2_e2d3_T24_B3")
    orc = O.TanOracle(sd, cfg["E"], cfg["D"], use_text_pos_enc=1, use_alignability_head=1)
    video, text = torch.from_numpy(batch["video"]), torch.from_numpy(batch["text"])
    a = orc.get_alignability(video, text)
    b = orc.get_alignability(video, text, (12, 3))
    for k in ("alignability-dual", "alignability-joint"):
        assert max_abs(a[k], g[k]) < 1e-4
        assert max_abs(b[k], g[k + "/interp_12_3"]) < 1e-4


def test_sine_position_table_vs_reference_fixture():
    """model/tfm_model.py:137-148 (host-side constant of the product)."""
    from temporalalignnet_b200.tfm_model import get_position_embedding_sine
    g = load_golden("g_blocks")
    assert np.array_equal(get_position_embedding_sine(8, 16).numpy(), g["sine_pos_16x8"])


def test_eager_port_matches_oracle_on_cpu():
    """oracle/eager_port.py (the torch-eager baseline bench.py times on the GPU: nn.MultiheadAttention, boolean-index
    loss) computes the same forward + loss as the oracle on the same weights (fp32, CPU)."""
    from oracle import eager_port as EP
    E, D, B, T, N = 2, 2, 3, 24, 5
    sd = synth.make_state_dict(E, D, seed=9)
    batch = synth.make_batch(B, T, N, seed=9, pad_video_every=2, force_full=True)
    m = EP.EagerTAN(E, D)
    missing, unexpected = m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    video, text = torch.from_numpy(batch["video"]), torch.from_numpy(batch["text"])
    vpm, tpm = torch.from_numpy(batch["video_padding_mask"]), torch.from_numpy(batch["text_padding_mask"])
    with torch.no_grad():
        out = m(video, text, vpm, tpm)
        loss = EP.eager_get_loss_init(out, batch["start"], batch["end"], tpm, T, N)
        ref = O.TanOracle(sd, E, D).forward(video, text, batch["video_padding_mask"], batch["text_padding_mask"])
        ref_loss = O.get_loss_init(ref["logits_dual"], ref["logits_joint"], batch["start"], batch["end"],
                                   batch["text_padding_mask"])
    assert max_abs(out["logits_dual"], ref["logits_dual"]) < FP32_TOL
    assert max_abs(out["logits_joint"], ref["logits_joint"]) < FP32_TOL
    for k in ("loss", "loss-dual", "loss-joint"):
        assert abs(float(loss[k]) - float(ref_loss[k])) < 1e-5 * abs(float(ref_loss[k])), k


def test_oracle_word2vec_vs_reference_fixture():
    """g_word2vec.npz: Word2VecModel.forward (model/word2vec_model.py:83-101) of the reference on synthetic weights,
    incl. an all-stop-word sentence (:93) and a repeated word; oracle forward + autograd against it."""
    from oracle.make_golden import make_word2vec_case
    g = load_golden("g_word2vec")
    sd, ids = make_word2vec_case()
    assert abs(sum(checksum(v) for v in sd.values()) + checksum(ids) - float(g["in_checksum"])) < 1e-6
    leaves = {k: torch.from_numpy(v).clone().requires_grad_(k != "word_embd.weight") for k, v in sd.items()}
    tok = torch.from_numpy(ids)
    pooled, last = O.word2vec_forward(leaves, tok, tok != 0)
    assert max_abs(pooled.detach(), g["pooler_output"]) < 1e-4
    assert max_abs(last.detach()[:, ::8], g["last_hidden_state"]) < 1e-4
    g_out = torch.from_numpy(synth._normal("w2v.gout", 888, tuple(pooled.shape), 1.0))
    (pooled * g_out).sum().backward()
    assert "grad_none/word_embd.weight" in g
    for k in ("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"):
        got = leaves[k].grad.double().reshape(-1)
        assert abs(float(got.norm()) - float(g["grad_norm/" + k])) < 1e-4 * float(g["grad_norm/" + k]), k
        assert max_abs(got[::97], g["grad_sub/" + k]) < 1e-4 * max(1.0, float(np.abs(g["grad_sub/" + k]).max())), k
