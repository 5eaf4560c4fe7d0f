"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: target/feature all-gather layout,
global column indexing (b_off), additivity of the fixed-shift column sums across row shards, and the
row (sum, count) reduction.  The per-rank exp-sums that the CUDA kernels would produce are computed
here with the torch oracle formulas; everything else is the product's own code
(loss.prepare_nce_inputs / gather_text_features / finish_loss)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import tan_oracle as O
from temporalalignnet_b200 import synth
from tests.helpers import cpu_pos_from_time, unpack_posbits


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cpu_reduce(row_sums, col_sums, out4, S=None, T=None, C=None, row_sel=None, col_sel=None):
    """torch restatement of tan_nce_reduce (checker for the CPU test)."""
    R = row_sums.shape[1]
    m = row_sums[1] > 0
    out4[0] += (row_sums[0][m].log() - row_sums[1][m].log()).double().sum()
    out4[1] += m.sum()
    c = col_sums.reshape(2, -1)
    m = c[1] > 0
    out4[2] += (c[0][m].log() - c[1][m].log()).double().sum()
    out4[3] += m.sum()


def _exp_sums(vn, tn, nce, B_loc, S, T, N):
    """What tan_sim_nce_fwd produces for local rows: vn [B_loc,S,T,d], tn [S,C,d] (global columns)."""
    cos = torch.einsum("bstd,scd->bstc", vn, tn)
    valid = nce.col_valid.bool()
    e = torch.exp((cos - 1.0) / 0.07) * valid.float()
    C = tn.shape[1]
    pos_own = unpack_posbits(nce.posbits, N).permute(0, 2, 1)                # [B_loc, T, N] targets of the local clips
    pos = torch.zeros(B_loc, T, C, dtype=torch.bool)
    for b in range(B_loc):
        pos[b, :, (nce.b_off + b) * N:(nce.b_off + b + 1) * N] = pos_own[b]
    pe = e * (pos & valid[None, None])[:, None].float()
    row = torch.stack((e.sum(-1).reshape(-1), pe.sum(-1).reshape(-1)))
    col = torch.stack((e.sum(dim=(0, 2)), pe.sum(dim=(0, 2))))
    return row.float(), col.float()


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from temporalalignnet_b200 import loss as L
    torch.set_num_threads(2)
    E = D = 2
    Bg, T, N = 4, 16, 4
    B_loc = Bg // world
    sd = synth.make_state_dict(E, D, seed=5)
    full = synth.make_batch(Bg, T, N, seed=5)
    sl = slice(rank * B_loc, (rank + 1) * B_loc)
    orc = O.TanOracle(sd, E, D)
    video = torch.from_numpy(full["video"][sl])
    text = torch.from_numpy(full["text"][sl])
    vpm = torch.from_numpy(full["video_padding_mask"][sl])
    tpm = torch.from_numpy(full["text_padding_mask"][sl])
    # local features (each clip's encoders are independent of the other clips)
    v = orc.get_visual_feature(video, vpm)
    vn = v / v.norm(dim=-1, keepdim=True)
    t = orc.get_textual_feature(text)
    tn = (t / t.norm(dim=-1, keepdim=True)).reshape(B_loc * N, -1)
    jv, jt = orc.get_joint_feature(video, vpm, t, tpm)
    jvn = jv / jv.norm(dim=-1, keepdim=True)
    jtn = (jt / jt.norm(dim=-1, keepdim=True)).permute(1, 0, 2, 3).reshape(D, B_loc * N, -1)
    nce = L.prepare_nce_inputs(full["start"][sl], full["end"][sl], tpm, T, N, torch.device("cpu"), shard=True,
                               pos_fn=cpu_pos_from_time)
    assert nce.b_off == rank * B_loc and nce.B_glob == Bg and nce.col_valid.numel() == Bg * N
    assert tuple(nce.posbits.shape) == (B_loc, T, 1)
    losses, sums = [], []
    for vfeat, tfeat, shared, S in ((vn, tn, True, E), (jvn, jtn, False, D)):
        tg = L.gather_text_features(tfeat.contiguous(), shared, dist)
        tg3 = tg[None].expand(S, -1, -1) if shared else tg
        row, col = _exp_sums(vfeat, tg3, nce, B_loc, S, T, N)
        col = col.contiguous()
        losses.append(L.finish_loss(row, col, dist, T, reduce_fn=_cpu_reduce))      # all-reduces `col` in place
        sums.append((row, col))
    loss = float((losses[0] + losses[1]) / 2)
    # ---- backward exchange (train.sim_coefficients + the all-reduce rule of train.step_backward) ----------------
    # The similarity-gradient kernels are replaced by their torch formula; the coefficient vectors, the stage-major
    # row layout, b_off and the sum of the text-feature gradients over ranks are the product's own host logic.
    from temporalalignnet_b200 import train as TR
    scale = torch.tensor(0.5)
    grad_err = 0.0
    for vfeat, tfeat, shared, S, (row, col) in ((vn, tn, True, E, sums[0]), (jvn, jtn, False, D, sums[1])):
        tg = L.gather_text_features(tfeat.contiguous(), shared, dist)
        tg3 = (tg[None].expand(S, -1, -1) if shared else tg).contiguous()
        C = tg3.shape[1]
        lg = type("LG", (), {})()
        lg.vfeat, lg.tfeat, lg.shared_text = vfeat, tfeat, shared
        ra, rap, cb, cbp = TR.sim_coefficients(TR.SimCtx(lg, row, col, nce), scale, dist)
        assert tuple(ra.shape) == (S, B_loc * T) and tuple(cb.shape) == (S, C)
        valid = nce.col_valid.bool()
        pos_own = unpack_posbits(nce.posbits, N).permute(0, 2, 1)
        pos = torch.zeros(B_loc * T, C)
        for b in range(B_loc):
            pos[b * T:(b + 1) * T, (nce.b_off + b) * N:(nce.b_off + b + 1) * N] = pos_own[b].float()
        d_v = torch.zeros(S, B_loc * T, vfeat.shape[-1])
        d_t = torch.zeros(S, C, vfeat.shape[-1])
        for s_ in range(S):
            a = vfeat[:, s_].reshape(B_loc * T, -1)
            e = torch.exp((a @ tg3[s_].t() - 1.0) / 0.07) * valid.float()[None]
            G = e * (ra[s_][:, None] + cb[s_][None] - pos * (rap[s_][:, None] + cbp[s_][None])) / 0.07
            d_v[s_] = G @ tg3[s_]
            d_t[s_] = G.t() @ a
        dist.all_reduce(d_t)                                    # text-feature gradients: sum over the ranks' rows
        if shared:
            d_t = d_t.sum(0, keepdim=True)
        # reference: autograd of the oracle's NCE on the GLOBAL batch
        vg = [torch.zeros_like(vfeat) for _ in range(world)]
        dist.all_gather(vg, vfeat.contiguous())
        vg = torch.cat(vg).clone().requires_grad_(True)
        tref = (tg.clone() if shared else tg3.clone()).requires_grad_(True)
        mask_g, _, _ = O.mask_from_time(full["start"], full["end"], T, N)
        tpm_g = torch.from_numpy(full["text_padding_mask"])
        tgt = torch.zeros(Bg, T, Bg, N, dtype=torch.bool)
        for b in range(Bg):
            tgt[b, :, b, :] = mask_g[b].t()
        cv = (~tpm_g).reshape(-1)
        tgt = tgt.reshape(Bg * T, Bg * N) & cv[None]
        logits = (torch.einsum("astc,kc->astk", vg, tref) if shared else
                  torch.einsum("astc,skc->astk", vg, tref)).reshape(Bg, S, T, Bg, N)
        (0.5 * O.nce_loss(logits, tgt, cv)).backward()
        ref_v = vg.grad[sl].permute(1, 0, 2, 3).reshape(S, B_loc * T, -1)
        ref_t = tref.grad[None] if shared else tref.grad
        grad_err = max(grad_err, float((d_v - ref_v).abs().max() / ref_v.abs().max()),
                       float((d_t - ref_t).abs().max() / ref_t.abs().max()))
    errs = [None] * world
    dist.all_gather_object(errs, grad_err)
    if rank == 0:
        ret["grad_err"] = max(errs)
    if rank == 0:
        ref_out = orc.forward(torch.from_numpy(full["video"]), torch.from_numpy(full["text"]),
                              full["video_padding_mask"], full["text_padding_mask"])
        ref = float(O.get_loss_init(ref_out["logits_dual"], ref_out["logits_joint"], full["start"], full["end"],
                                    full["text_padding_mask"])["loss"])
        ret["loss"], ret["ref"] = loss, ref
    gathered = [None] * world
    dist.all_gather_object(gathered, loss)
    if rank == 0:
        ret["all"] = gathered
    dist.destroy_process_group()


def test_two_rank_sharded_loss_equals_single_process_oracle():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert abs(ret["loss"] - ret["ref"]) < 1e-5 * abs(ret["ref"]), (ret["loss"], ret["ref"])
    assert all(abs(x - ret["loss"]) < 1e-7 for x in ret["all"])      # every rank returns the global loss
    # sharded feature gradients (coefficients + exchange rule of the backward pass) == autograd on the global batch
    assert ret["grad_err"] < 1e-4, ret["grad_err"]


def test_prepare_nce_inputs_single_process_layout():
    from temporalalignnet_b200 import loss as L
    b = synth.make_batch(3, 20, 5, seed=2)
    nce = L.prepare_nce_inputs(b["start"], b["end"], torch.from_numpy(b["text_padding_mask"]), 20, 5,
                               torch.device("cpu"), shard=False, pos_fn=cpu_pos_from_time)
    mask, start, end = O.mask_from_time(b["start"], b["end"], 20, 5)
    assert torch.equal(unpack_posbits(nce.posbits, 5), mask & ~torch.from_numpy(b["text_padding_mask"])[:, :, None])
    s2, e2 = L.padded_times(b["start"], b["end"], 20, 5, torch.device("cpu"))
    assert torch.equal(s2, start) and torch.equal(e2, end)
    assert torch.equal(nce.col_valid.view(3, 5).bool(), ~torch.from_numpy(b["text_padding_mask"]))
    m2, s2, e2 = L.get_mask_from_time(b["start"], b["end"], 20, 5, device="cpu")
    assert torch.equal(m2[:, :mask.shape[1]], mask[:, :m2.shape[1]])


def test_circulant_known_answer():
    from temporalalignnet_b200.loss import circulant
    assert circulant(torch.tensor([0, 1, 2]), dim=0).tolist() == [[0, 1, 2], [2, 0, 1], [1, 2, 0]]


def test_state_dict_keys_match_reference_names(monkeypatch):
    from temporalalignnet_b200 import TemporalAligner, TwinTemporalAligner, optim
    from tests import cpu_ops
    monkeypatch.setattr(optim, "ema_update", cpu_ops.ema_update)      # the product's EMA update is a CUDA kernel
    m = TemporalAligner(2, 3, random_pos_start=0, use_alignability_head=1)
    ref_keys = set(synth.make_state_dict(2, 3, use_alignability_head=True))
    assert set(m.state_dict().keys()) == ref_keys
    tw = TwinTemporalAligner(m=0.99, num_encoder_layers=1, num_decoder_layers=1)
    assert all(k.startswith(("online.", "target.")) for k in tw.state_dict())
    assert tw.target.random_pos_start == 0
    p0 = [p.clone() for p in tw.target.parameters()]
    for p in tw.online.parameters():
        p.data.add_(1.0)
    v0 = [p._version for p in tw.target.parameters()]
    tw._momentum_update()
    for a, b, o in zip(p0, tw.target.parameters(), tw.online.parameters()):
        assert torch.allclose(b, a * 0.99 + o * 0.01, atol=1e-6)
    # the bf16 weight shadows of the target watch the version counters (tfm_model._Bf16Cache): every update
    # must advance them, and so must _copy_param
    assert all(p._version > v for p, v in zip(tw.target.parameters(), v0))
    v1 = [p._version for p in tw.target.parameters()]
    tw._copy_param()
    assert all(p._version > v for p, v in zip(tw.target.parameters(), v1))
    assert all(torch.equal(a, b) for a, b in zip(tw.target.parameters(), tw.online.parameters()))


def _train_worker(rank, world, port, ret):
    """One rank of a sharded TRAINING step on CPU: the product's forward_train / get_loss / step_backward with the
    C-ABI wrappers replaced by torch stand-ins (tests/cpu_ops.py) and gloo collectives."""
    import types
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from tests import cpu_ops
    cpu_ops.install_plain()
    from temporalalignnet_b200 import TemporalAligner, get_loss
    E, D, Bg, T, N = 1, 2, 4, 16, 4
    B_loc = Bg // world
    sd = synth.make_state_dict(E, D, seed=5)
    full = synth.make_batch(Bg, T, N, seed=5, pad_video_every=3)
    args = types.SimpleNamespace(model="init", sim="cos", learn_agreement=0, temporal_agreement_type="keep",
                                 loss_threshold=0.0, use_alignability_head=0, optim_policy="default")
    m = TemporalAligner(E, D, random_pos_start=0, use_text_pos_enc=1)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    m.train()
    m.enable_autograd(True)

    def run(lo, hi, shard):
        for p in m.parameters():
            p.grad = None
        video, text = torch.from_numpy(full["video"][lo:hi]), torch.from_numpy(full["text"][lo:hi])
        vpm = torch.from_numpy(full["video_padding_mask"][lo:hi])
        tpm = torch.from_numpy(full["text_padding_mask"][lo:hi])
        out = m(video, text, video_padding_mask=vpm, lang_padding_mask=tpm)
        ld = get_loss({"start": full["start"][lo:hi], "end": full["end"][lo:hi], "text": full["text_str"][lo:hi]},
                      video, text, vpm.float(), tpm.float(), out, args, None, shard_batch=shard)
        ld["loss"].backward()
        return float(ld["loss"]), {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}

    loss_s, g_s = run(rank * B_loc, (rank + 1) * B_loc, True)
    loss_1, g_1 = run(0, Bg, False)
    worst = max(float((g_s[n] - g_1[n]).norm() / g_1[n].norm().clamp_min(1e-30)) for n in g_1)
    res = [None] * world
    dist.all_gather_object(res, (abs(loss_s - loss_1) / abs(loss_1), worst, set(g_s) == set(g_1)))
    if rank == 0:
        ret["train"] = res
    dist.destroy_process_group()


def test_two_rank_sharded_training_step_equals_single_process():
    """Sharded gradients (text features all-gathered, text-feature and weight gradients all-reduced inside
    loss.backward()) == the gradients of the same global batch in one process.  (CPU matmuls block differently for
    different batch sizes, so a few bf16 roundings of the stand-ins flip between the two runs: 2e-4 on the loss,
    where the GPU kernels -- bitwise batch-invariant per clip -- give 0 ulp, scripts/multigpu_train_check.py.)"""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_train_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for loss_err, grad_err, same_keys in ret["train"]:
        assert same_keys
        assert loss_err < 2e-4, loss_err
        assert grad_err < 2e-2, grad_err


def _shape_worker(rank, world, port, ret):
    """Ranks with different local N (each rank's own pad_sequence length, as with real data): get_loss must refuse
    with a clear error, pad_text_to_global must bring every rank to the longest N."""
    import types
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from temporalalignnet_b200 import TanError
    from temporalalignnet_b200 import loss as L
    B, T, N = 2, 16, 3 + rank                                  # rank 1 has one more sentence column
    assert L.shard_shapes(B, N, "cpu") == (2, 2, 3, 4)
    text = torch.randn(B, N, 8)
    tpm = torch.zeros(B, N, dtype=torch.bool)
    tpm[1, -1] = True
    emb2, mask2 = L.pad_text_to_global(text, tpm)
    ok = tuple(emb2.shape) == (B, 4, 8) and tuple(mask2.shape) == (B, 4)
    if rank == 0:
        ok = ok and bool(mask2[:, 3].all()) and torch.equal(emb2[:, 3], text[:, 2]) and torch.equal(emb2[:, :3], text)
    else:
        ok = ok and emb2 is text and mask2 is tpm
    # get_loss on the unpadded, disagreeing shapes refuses before any collective of mismatched size
    args = types.SimpleNamespace(model="init", sim="cos", learn_agreement=0, loss_threshold=0.0, use_alignability_head=0)
    lg = torch.zeros(B, 1, T, B, N)
    err = ""
    try:
        L.get_loss({"start": [[0.0]] * B, "end": [[1.0]] * B, "text": None}, torch.zeros(B, T, 4), text, None, tpm,
                   {"logits_dual": lg, "logits_joint": lg}, args, None)
    except TanError as e:
        err = str(e)
    # a loader that knows the global sentence counts passes them: no shape exchange (no host rendezvous) in the step.
    # (On CPU get_loss cannot get past its first kernel, tan_pos_from_time: reaching THAT error means the exchange
    # stage was passed without the collective.)
    hint_ok = True
    data_h = {"start": [[0.0]] * B, "end": [[1.0]] * B, "text": None}
    lg_h = torch.zeros(B, 1, T, world * B, 4)
    real_exchange = L.shard_exchange
    L.shard_exchange = lambda *a_, **k_: (_ for _ in ()).throw(AssertionError("exchange called despite the hint"))
    try:
        for hint, want in (([1] * (world * B), "pos_from_time"), ([1], "n_sentences_global")):
            try:
                L.get_loss(dict(data_h, n_sentences_global=hint), torch.zeros(B, T, 4), emb2, None, mask2,
                           {"logits_dual": lg_h, "logits_joint": lg_h}, args, None)
                hint_ok = False
            except TanError as e:
                hint_ok = hint_ok and want in str(e)
    finally:
        L.shard_exchange = real_exchange
    # different clip counts are an error of their own
    err_b = ""
    try:
        L.pad_text_to_global(torch.randn(B + rank, 4, 8), torch.zeros(B + rank, 4, dtype=torch.bool))
    except TanError as e:
        err_b = str(e)
    res = [None] * world
    dist.all_gather_object(res, (ok, "pad_text_to_global" in err, "same number of clips" in err_b, hint_ok))
    if rank == 0:
        ret["res"] = res
    dist.destroy_process_group()


def test_sharded_shape_disagreement_is_detected_and_padding_helper_fixes_it():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_shape_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(all(r) for r in ret["res"]), ret["res"]
