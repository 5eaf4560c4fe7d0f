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
    # the reference's procedure: one batch-1 call per window for the joint and for the dual model
    orc = O.TanOracle(sd, E, D, use_alignability_head=head)
    sim_j = torch.zeros(n_text, vlen)
    sim_d = torch.zeros(n_text, vlen)
    cover = torch.zeros(n_text, vlen)
    for t0, t1, n0, n1 in windows:
        v, t = video[None, t0:t1], text[None, n0:n1]
        sim_j[n0:n1, t0:t1] += orc.get_text_visual_sim_joint(v, t)[0, -1].t() / 0.07
        sim_d[n0:n1, t0:t1] += orc.get_text_visual_sim_dual(v, t)[0, -1].t() / 0.07
        cover[n0:n1, t0:t1] += 1
    ref_j, ref_d = sim_j / cover.clamp(min=1e-5), sim_d / cover.clamp(min=1e-5)
    assert torch.equal(res["overlap"].cpu(), cover)
    assert (res["sim-joint"].cpu() - ref_j).abs().max().item() < 0.08        # cosine error 4e-3 (bf16) / 0.07
    assert (res["sim-dual"].cpu() - ref_d).abs().max().item() < 0.08
    assert (res["sim"].cpu() - (ref_j + ref_d) / 2).abs().max().item() < 0.08
    assert predicted_frames(res["sim"]).shape == (n_text,)
    if head:
        assert res["alignability-joint"].shape == (n_text,) and torch.isfinite(res["alignability-joint"]).all()
