"""Host-side window planning of the sliding-window alignment (no GPU)."""
import numpy as np


def test_plan_windows_matches_reference_rules():
    from temporalalignnet_b200.align import plan_windows
    vlen, seq_len = 100, 32
    mid = np.linspace(2, 97, 14)
    anchors = np.ones(14, bool)
    anchors[::3] = False
    w = plan_windows(vlen, seq_len, mid, anchors)
    steps = [t0 for t0, _, _, _ in w]
    assert steps == sorted(steps) and all(s % 8 == 0 for s in steps)
    assert all(t1 - t0 <= seq_len and t1 <= vlen for t0, t1, _, _ in w)
    assert w[0][2] == 0 and w[-1][3] == 14                 # first / last windows reach the first / last sentence
    assert all(n1 > n0 for _, _, n0, n1 in w)


def _ranges_from_oracle(vlen, seq_len, mid, aligned):
    from oracle import tan_oracle as O
    out = []
    for t0, t1, mask in O.overlap_seq_windows(vlen, seq_len, mid, aligned):
        idx = np.flatnonzero(mask)
        assert np.array_equal(idx, np.arange(idx[0], idx[-1] + 1))       # the reference's masks are contiguous ranges
        out.append((t0, t1, int(idx[0]), int(idx[-1]) + 1))
    return out


def test_plan_windows_equals_the_reference_loop_on_random_videos():
    """plan_windows vs the oracle's restatement of eval/eval_zeroshot_align.py:129-177 (mask per step), over
    random lengths, sentence counts (also more sentences than frames), timestamps and alignable sets."""
    from temporalalignnet_b200.align import plan_windows
    rng = np.random.default_rng(7)
    cases = 0
    for _ in range(300):
        seq_len = int(rng.choice([4, 8, 32, 64]))
        vlen = int(rng.integers(1, 400))
        n_text = int(rng.integers(1, 60)) if rng.random() < 0.8 else vlen + int(rng.integers(2, 30))
        mid = np.sort(rng.uniform(-5, vlen + 5, n_text)) if rng.random() < 0.7 else rng.uniform(0, vlen, n_text)
        aligned = rng.random(n_text) < rng.choice([0.0, 0.3, 0.7, 1.0])
        want = _ranges_from_oracle(vlen, seq_len, mid, aligned)
        got = plan_windows(vlen, seq_len, mid, ~aligned)
        assert got == want, (vlen, seq_len, n_text)
        cases += len(want)
    assert cases > 1000


def test_plan_windows_edge_cases():
    from temporalalignnet_b200.align import plan_windows
    assert plan_windows(10, 32, [1.0, 5.0], [False, False]) == []          # no anchor sentence: no window
    assert plan_windows(10, 32, [], []) == []
    # a video shorter than half a window has no step at all (np.arange(0, vlen - seq_len // 2, .) is empty)
    assert plan_windows(16, 32, [3.0], [True]) == []
    w = plan_windows(17, 32, [3.0], [True])
    assert w == [(0, 17, 0, 1)]


import pytest  # noqa: E402
import torch  # noqa: E402


@pytest.mark.parametrize("head,per_batch", [(0, 256), (1, 256), (0, 3)])
def test_sliding_window_alignment_host_logic_vs_oracle(monkeypatch, head, per_batch):
    """The HOST logic of sliding_window_alignment (windows batched as clips, padded frames / sentences, the
    per-sentence accumulators with and without an alignability head, several batches of windows) with the C-ABI
    wrappers replaced by torch stand-ins (tests/cpu_ops.py), against the oracle's restatement of the reference loop
    (eval/eval_zeroshot_align.py:129-205) that runs every window on its own."""
    from oracle import tan_oracle as O
    from temporalalignnet_b200 import TemporalAligner, synth
    from temporalalignnet_b200.align import plan_windows, predicted_frames, sliding_window_alignment
    from tests import cpu_ops
    cpu_ops.install(monkeypatch)
    E, D, vlen, seq_len, n_text = 1, 3, 84, 32, 9                       # last window: 20 real frames of 32
    sd = synth.make_state_dict(E, D, use_alignability_head=bool(head), seed=5)
    g = torch.Generator().manual_seed(6)
    video, text = torch.randn(vlen, 1024, generator=g), torch.randn(n_text, 512, generator=g)
    mid = np.linspace(3, 80, n_text)
    anchors = np.ones(n_text, bool)
    anchors[1::4] = False
    windows = plan_windows(vlen, seq_len, mid, anchors)
    assert any(t1 - t0 < seq_len for t0, t1, _, _ in windows) and len({n1 - n0 for _, _, n0, n1 in windows}) > 1
    m = TemporalAligner(E, D, random_pos_start=0, use_alignability_head=head)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    # the inference forward drives two CUDA streams; on CPU the windows go through the training forward (one stream,
    # same kernels sequence per stack, same output dict), which the stand-ins can run
    from temporalalignnet_b200 import train
    monkeypatch.setattr(m, "_forward_impl", lambda v, t, vpm, tpm: train.forward_train(m, v, t, vpm, tpm, None))
    res = sliding_window_alignment(m, video, text, windows, max_windows_per_batch=per_batch)

    orc = O.TanOracle(sd, E, D, use_alignability_head=head)

    def sim_fn(t0, t1, mask):
        v, t = video[None, t0:t1], text[None, torch.from_numpy(mask)]
        o = {"sim": orc.get_text_visual_sim_joint(v, t).transpose(-1, -2) / 0.07,
             "dual-sim": orc.get_text_visual_sim_dual(v, t).transpose(-1, -2) / 0.07}
        if head:
            o.update(orc.get_alignability(v, t))
        return o

    ref = O.overlap_seq_alignment(sim_fn, vlen, n_text, O.overlap_seq_windows(vlen, seq_len, mid, ~anchors), bool(head))
    assert torch.equal(res["overlap"], ref["overlap"])
    for k in ("sim-joint", "sim-dual", "sim"):
        assert (res[k] - ref[k]).abs().max().item() < 0.08, k           # bf16 features: cosine error 4e-3 / 0.07
    for k in ("alignability-dual", "alignability-joint"):
        assert (res[k] - ref[k]).abs().max().item() < (3e-2 if head else 0.08), k
    want = torch.where(ref["sim"] != 0, ref["sim"], torch.full_like(ref["sim"], -6e4)).argmax(-1)
    got = predicted_frames(res["sim"])
    top = ref["sim"].gather(1, got[:, None])[:, 0]                     # near-ties may flip under bf16: compare values
    assert ((ref["sim"].gather(1, want[:, None])[:, 0] - top).abs() < 0.16).all()


def test_roc_auc_equals_sklearn_with_and_without_ties():
    metrics = pytest.importorskip("sklearn.metrics")
    from temporalalignnet_b200.align import roc_auc
    rng = np.random.default_rng(3)
    for n in (2, 5, 64, 1000):
        for ties in (False, True):
            y = rng.random(n) < 0.4
            y[0], y[1] = True, False
            s = rng.integers(0, 6, n).astype(np.float64) if ties else rng.normal(size=n)
            assert abs(roc_auc(y, s) - metrics.roc_auc_score(y, s)) < 1e-12
    with pytest.raises(ValueError):
        roc_auc([1, 1, 1], [0.1, 0.2, 0.3])


@pytest.mark.parametrize("head", [0, 1])
def test_global_alignment_and_meter_host_logic_vs_oracle(monkeypatch, head):
    """global_alignment (one pass, interpolated positional table) and AlignmentMeter (Recall / AUC bookkeeping) with
    the stand-in kernels, against the oracle's restatement of eval/eval_zeroshot_align.py:207-249, over three videos
    (two 'global', one 'overlap-seq' result)."""
    from oracle import tan_oracle as O
    from temporalalignnet_b200 import TemporalAligner, synth, train
    from temporalalignnet_b200.align import (AlignmentMeter, global_alignment, plan_windows,
                                             sliding_window_alignment)
    from tests import cpu_ops
    cpu_ops.install(monkeypatch)
    E, D, seq_len = 1, 3, 32
    sd = synth.make_state_dict(E, D, use_alignability_head=bool(head), seed=8)
    m = TemporalAligner(E, D, random_pos_start=0, use_alignability_head=head)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    monkeypatch.setattr(m, "_forward_impl", lambda v, t, vpm, tpm: train.forward_train(m, v, t, vpm, tpm, None))
    orc = O.TanOracle(sd, E, D, use_alignability_head=head)
    g = torch.Generator().manual_seed(9)
    rng = np.random.default_rng(9)
    meter, ref_videos = AlignmentMeter(head), []
    for vi, (vlen, n_text) in enumerate([(50, 7), (32, 5), (70, 8)]):
        video, text = torch.randn(vlen, 1024, generator=g), torch.randn(n_text, 512, generator=g)
        start = np.sort(rng.uniform(0, vlen - 6, n_text))
        end = start + rng.uniform(1, 6, n_text)
        aligned = rng.random(n_text) < 0.6
        aligned[0], aligned[1] = True, False

        def sim_fn(t0=0, t1=vlen, mask=np.ones(n_text, bool), interp=seq_len):
            v, t = video[None, t0:t1], text[None, torch.from_numpy(mask)]
            o = {"sim": orc.get_text_visual_sim_joint(v, t, interp).transpose(-1, -2) / 0.07,
                 "dual-sim": orc.get_text_visual_sim_dual(v, t, interp).transpose(-1, -2) / 0.07}
            if head:
                o.update(orc.get_alignability(v, t, interp))
            return o

        if vi < 2:
            res = global_alignment(m, video, text, seq_len)
            ref = O.global_alignment(sim_fn, bool(head))
            assert res["sim"] is res["sim-joint"]
        else:
            mid = (start + end) / 2
            res = sliding_window_alignment(m, video, text, plan_windows(vlen, seq_len, mid, ~aligned))
            ref = O.overlap_seq_alignment(lambda t0, t1, mask: sim_fn(t0, t1, mask, None), vlen, n_text,
                                          O.overlap_seq_windows(vlen, seq_len, mid, aligned), bool(head))
        for k in ("sim", "sim-dual"):
            assert (res[k] - ref[k]).abs().max().item() < 0.08, (vi, k)
        for k in ("alignability-dual", "alignability-joint"):
            assert (res[k] - ref[k]).abs().max().item() < (3e-2 if head else 0.08), (vi, k)
        # the bookkeeping itself is compared on IDENTICAL inputs (the oracle's fp32 result), so that a near-tie
        # between two frames under bf16 cannot flip a hit
        meter.update(ref, aligned, start, end)
        ref_videos.append((ref, aligned, start, end))
    got, want = meter.compute(), O.htm_align_metrics(ref_videos, bool(head))
    assert got["Recall"] == want["Recall"] and abs(got["AUC"] - want["AUC"]) < 1e-12, (got, want)


@pytest.mark.parametrize("method", ["overlap-seq", "global"])
def test_evaluate_alignment_loop_vs_oracle(monkeypatch, method):
    """evaluate_alignment == the oracle's composition of the reference's loop (eval_zeroshot_align.py:97-252) on a
    small synthetic test set; the sentences of a video are separated well enough for bf16 not to flip a decision."""
    from oracle import tan_oracle as O
    from temporalalignnet_b200 import TemporalAligner, synth, train
    from temporalalignnet_b200.align import evaluate_alignment
    from tests import cpu_ops
    cpu_ops.install(monkeypatch)
    E, D, seq_len = 1, 1, 32
    sd = synth.make_state_dict(E, D, seed=21)
    m = TemporalAligner(E, D, random_pos_start=0)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    monkeypatch.setattr(m, "_forward_impl", lambda v, t, vpm, tpm: train.forward_train(m, v, t, vpm, tpm, None))
    orc = O.TanOracle(sd, E, D)
    g = torch.Generator().manual_seed(22)
    rng = np.random.default_rng(22)
    samples, ref_videos = [], []
    for vlen, n_text in [(48, 6), (64, 9), (40, 4)]:
        video, text = torch.randn(vlen, 1024, generator=g), torch.randn(n_text, 512, generator=g)
        start = np.sort(rng.uniform(0, vlen - 6, n_text))
        end = start + rng.uniform(1, 6, n_text)
        aligned = rng.random(n_text) < 0.6
        aligned[0], aligned[1] = True, False
        samples.append({"video": video[None], "str": [f"s{i}" for i in range(n_text)], "text_embed_": text,
                        "start": torch.tensor(start), "end": torch.tensor(end), "aligned": torch.tensor(aligned)})

        def sim_fn(t0=0, t1=vlen, mask=np.ones(n_text, bool), interp=seq_len, video=video, text=text):
            v, t = video[None, t0:t1], text[None, torch.from_numpy(mask)]
            return {"sim": orc.get_text_visual_sim_joint(v, t, interp).transpose(-1, -2) / 0.07,
                    "dual-sim": orc.get_text_visual_sim_dual(v, t, interp).transpose(-1, -2) / 0.07}

        if method == "global":
            ref = O.global_alignment(sim_fn, False)
        else:
            ref = O.overlap_seq_alignment(lambda t0, t1, mask: sim_fn(t0, t1, mask, None), vlen, n_text,
                                          O.overlap_seq_windows(vlen, seq_len, (start + end) / 2, aligned), False)
        ref_videos.append((ref, aligned, start, end))
    by_str = {tuple(s["str"]): s["text_embed_"] for s in samples}
    got = evaluate_alignment(m, samples, seq_len, method, embed_text=lambda strs: by_str[tuple(strs)])
    want = O.htm_align_metrics(ref_videos, False)
    # decisions: allow one sentence of the 11 alignable ones to flip under bf16 features; scores: AUC over 19 points
    n_al = sum(int(a.sum()) for _, a, _, _ in ref_videos)
    assert abs(got["Recall"] - want["Recall"]) <= 1.0 / n_al + 1e-9, (got, want)
    assert abs(got["AUC"] - want["AUC"]) < 0.05, (got, want)
    with pytest.raises(Exception):
        evaluate_alignment(m, samples, seq_len, "nearest")
