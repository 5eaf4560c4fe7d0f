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
