"""Sliding-window alignment (eval/eval_zeroshot_align.py 'overlap-seq'): all windows of a video batched into one
forward vs the oracle running every window on its own, unmasked, like the reference does."""
import numpy as np
import pytest
import torch

from oracle import tan_oracle as O
from temporalalignnet_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("head", [0, 1])
def test_sliding_window_alignment_vs_oracle(head):
    from temporalalignnet_b200 import TemporalAligner
    from temporalalignnet_b200.align import plan_windows, predicted_frames, sliding_window_alignment
    E, D, vlen, seq_len, n_text = 2, 3, 100, 32, 14
    sd = synth.make_state_dict(E, D, use_alignability_head=bool(head), seed=11)
    g = torch.Generator().manual_seed(3)
    video = torch.randn(vlen, 1024, generator=g)
    text = torch.randn(n_text, 512, generator=g)
    mid = np.linspace(2, 97, n_text)
    anchors = np.ones(n_text, bool)
    anchors[::3] = False
    windows = plan_windows(vlen, seq_len, mid, anchors)
    m = TemporalAligner(E, D, random_pos_start=0, use_alignability_head=head)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    m = m.to(DEV)
    res = sliding_window_alignment(m, video.to(DEV), text.to(DEV), windows)
    # the reference's procedure (oracle.overlap_seq_*: eval/eval_zeroshot_align.py:129-205): one batch-1 call per
    # window for the joint and for the dual model, accumulated sentence mask by sentence mask
    orc = O.TanOracle(sd, E, D, use_alignability_head=head)
    ref_windows = O.overlap_seq_windows(vlen, seq_len, mid, ~anchors)
    assert [(t0, t1) for t0, t1, _ in ref_windows] == [(t0, t1) for t0, t1, _, _ in windows]

    def sim_fn(t0, t1, mask):
        v, t = video[None, t0:t1], text[None, torch.from_numpy(mask)]
        o = {"sim": orc.get_text_visual_sim_joint(v, t).transpose(-1, -2) / 0.07,           # train/main.py:184-186
             "dual-sim": orc.get_text_visual_sim_dual(v, t).transpose(-1, -2) / 0.07}
        if head:
            o.update(orc.get_alignability(v, t))
        return o

    ref = O.overlap_seq_alignment(sim_fn, vlen, n_text, ref_windows, bool(head))
    assert torch.equal(res["overlap"].cpu(), ref["overlap"])
    for k in ("sim-joint", "sim-dual", "sim"):
        assert (res[k].cpu() - ref[k]).abs().max().item() < 0.08, k              # cosine error 4e-3 (bf16) / 0.07
    # per-sentence alignability: the head's logits (2e-2 like the forward parity tests, averaged over windows) or,
    # without a head, the window maxima of the similarities
    tol = 3e-2 if head else 0.08
    for k in ("alignability-dual", "alignability-joint"):
        assert res[k].shape == (n_text,)
        assert (res[k].cpu() - ref[k]).abs().max().item() < tol, k
    assert predicted_frames(res["sim"]).shape == (n_text,)


def test_align_stitch_kernel_equals_the_reference_loop_bitwise():
    """tan_align_stitch in two batches of windows == the reference's python loop of slice additions + division."""
    from temporalalignnet_b200 import ops
    g = torch.Generator().manual_seed(5)
    n_text, vlen, T, N = 23, 157, 32, 9
    wins = []
    for step in range(0, vlen - T // 2, T // 4):
        n0 = int(torch.randint(0, n_text - 2, (1,), generator=g))
        n1 = min(n_text, n0 + 1 + int(torch.randint(0, N, (1,), generator=g)))
        wins.append((step, min(vlen, step + T), n0, n1))
    W = len(wins)
    blk_j = torch.randn(W, T, N, generator=g).to(DEV)
    blk_d = torch.randn(W, T, N, generator=g).to(DEV)
    ref_j = torch.zeros(n_text, vlen, device=DEV)
    ref_d = torch.zeros_like(ref_j)
    cov = torch.zeros_like(ref_j)
    for i, (t0, t1, n0, n1) in enumerate(wins):
        ref_j[n0:n1, t0:t1] += (blk_j[i] / 0.07)[:t1 - t0, :n1 - n0].t()
        ref_d[n0:n1, t0:t1] += (blk_d[i] / 0.07)[:t1 - t0, :n1 - n0].t()
        cov[n0:n1, t0:t1] += 1
    ref_j, ref_d = ref_j / cov.clamp(min=1e-5), ref_d / cov.clamp(min=1e-5)
    sj = torch.full((n_text, vlen), float("nan"), device=DEV)
    sd, cv = sj.clone(), sj.clone()
    half = W // 2
    win_t = torch.tensor(wins, dtype=torch.int32, device=DEV)
    ops.align_stitch(blk_j[:half].contiguous(), blk_d[:half].contiguous(), win_t[:half].contiguous(), sj, sd, cv, False, False)
    ops.align_stitch(blk_j[half:].contiguous(), blk_d[half:].contiguous(), win_t[half:].contiguous(), sj, sd, cv, True, True)
    assert torch.equal(cv, cov)
    assert (sj - ref_j).abs().max().item() <= 1e-6 * ref_j.abs().max().item()
    assert (sd - ref_d).abs().max().item() <= 1e-6 * ref_d.abs().max().item()
    assert torch.equal(sj, ref_j) and torch.equal(sd, ref_d)         # same operations, same order, same roundings
    assert (cov == 0).any() and (cov > 1).any()                      # the case has uncovered and overlapped entries


def test_align_argmax_equals_softmax_argmax():
    from temporalalignnet_b200.align import predicted_frames
    g = torch.Generator().manual_seed(6)
    sim = (torch.randn(37, 301, generator=g) * 5).to(DEV)
    sim[:, 250:] = 0.0                                               # uncovered tail
    sim[3] = 0.0                                                     # a sentence no window covers -> frame 0
    sim[5, 17] = sim[5].max() + 1.0
    sim[5, 200] = sim[5, 17]                                         # exact tie -> first index
    ref = sim.masked_fill(sim == 0, -6e4).softmax(-1).argmax(-1)
    got = predicted_frames(sim)
    assert got.dtype == torch.int64 and torch.equal(got, ref)
    assert int(got[3]) == 0 and int(got[5]) == 17


def test_sliding_window_alignment_in_several_batches_equals_one_batch():
    from temporalalignnet_b200 import TemporalAligner
    from temporalalignnet_b200.align import plan_windows, sliding_window_alignment
    E, D, vlen, seq_len, n_text = 1, 3, 90, 32, 10                   # the head reads joint stage 2
    sd = synth.make_state_dict(E, D, use_alignability_head=True, seed=12)
    g = torch.Generator().manual_seed(4)
    video, text = torch.randn(vlen, 1024, generator=g).to(DEV), torch.randn(n_text, 512, generator=g).to(DEV)
    windows = plan_windows(vlen, seq_len, np.linspace(1, 88, n_text), np.ones(n_text, bool))
    m = TemporalAligner(E, D, random_pos_start=0, use_alignability_head=1)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    m = m.to(DEV)
    one = sliding_window_alignment(m, video, text, windows)
    many = sliding_window_alignment(m, video, text, windows, max_windows_per_batch=3)
    assert torch.equal(one["overlap"], many["overlap"])
    for k in ("sim-joint", "sim-dual", "alignability-dual", "alignability-joint"):
        assert (one[k] - many[k]).abs().max().item() < 2e-2 * max(one[k].abs().max().item(), 1.0), k


@pytest.mark.parametrize("head", [0, 1])
def test_global_alignment_and_meter_vs_oracle(head):
    """The 'global' method (one pass over the whole video, positional table interpolated from seq_len) against the
    oracle's restatement of eval/eval_zeroshot_align.py:207-215, and the Recall / AUC bookkeeping (:217-249) with
    the decision kernel on identical inputs."""
    from temporalalignnet_b200 import TemporalAligner
    from temporalalignnet_b200.align import AlignmentMeter, global_alignment
    E, D, seq_len = 2, 3, 32
    sd = synth.make_state_dict(E, D, use_alignability_head=bool(head), seed=13)
    m = TemporalAligner(E, D, random_pos_start=0, use_alignability_head=head)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    m = m.to(DEV)
    orc = O.TanOracle(sd, E, D, use_alignability_head=head)
    g = torch.Generator().manual_seed(14)
    rng = np.random.default_rng(14)
    meter, ref_videos = AlignmentMeter(head), []
    for vlen, n_text in [(50, 7), (32, 5), (200, 12)]:
        video, text = torch.randn(vlen, 1024, generator=g), torch.randn(n_text, 512, generator=g)

        def sim_fn():
            v, t = video[None], text[None]
            o = {"sim": orc.get_text_visual_sim_joint(v, t, seq_len).transpose(-1, -2) / 0.07,
                 "dual-sim": orc.get_text_visual_sim_dual(v, t, seq_len).transpose(-1, -2) / 0.07}
            if head:
                o.update(orc.get_alignability(v, t, seq_len))
            return o

        res = global_alignment(m, video.to(DEV), text.to(DEV), seq_len)
        ref = O.global_alignment(sim_fn, bool(head))
        assert res["sim"].shape == (n_text, vlen)
        for k in ("sim", "sim-dual"):
            assert (res[k].cpu() - ref[k]).abs().max().item() < 0.08, k          # cosine error 4e-3 (bf16) / 0.07
        for k in ("alignability-dual", "alignability-joint"):
            assert (res[k].cpu() - ref[k]).abs().max().item() < (3e-2 if head else 0.08), k
        start = np.sort(rng.uniform(0, vlen - 6, n_text))
        end = start + rng.uniform(1, 6, n_text)
        aligned = rng.random(n_text) < 0.6
        aligned[0], aligned[1] = True, False
        meter.update({k: v.to(DEV) for k, v in ref.items()}, aligned, start, end)
        ref_videos.append((ref, aligned, start, end))
    got, want = meter.compute(), O.htm_align_metrics(ref_videos, bool(head))
    assert got["Recall"] == want["Recall"] and abs(got["AUC"] - want["AUC"]) < 1e-6, (got, want)
