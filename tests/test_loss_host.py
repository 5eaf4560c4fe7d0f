"""Host-side helpers of get_loss (no GPU): the sync-free quantile / standardisation / bit packing glue of the
threshold and alignability branches against torch's own functions and the reference's formulas."""
import pytest
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from temporalalignnet_b200 import loss as L
from tests.helpers import pack_posbits


@settings(max_examples=60, deadline=None)
@given(st.integers(2, 40), st.floats(0.0, 1.0), st.integers(0, 2 ** 31 - 1))
def test_masked_quantile_equals_torch_quantile(n, q, seed):
    """torch.quantile(x[valid], q) (train/loss.py:286,:318-319 use it on boolean-indexed tensors = a host sync)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, generator=g)
    valid = torch.rand(n, generator=g) > 0.3
    if valid.sum() == 0:
        valid[0] = True
    got = L.masked_quantile(x, valid, q)
    ref = torch.quantile(x[valid], q)
    assert abs(float(got) - float(ref)) <= 1e-6 * max(1.0, abs(float(ref)))


def test_masked_standardise_equals_reference_formula():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(50, generator=g)
    valid = torch.rand(50, generator=g) > 0.4
    got = L._masked_standardise(x, valid)[valid]
    sel = x[valid]
    ref = (sel - sel.mean()) / sel.std()                     # train/loss.py:281,:283
    assert torch.allclose(got, ref, atol=1e-5)


@pytest.mark.parametrize("B,N", [(3, 5), (2, 32), (4, 40), (1, 64)])
def test_pack_bits_layout_matches_posbits_words(B, N):
    """_pack_bits([B, N]) uses the word layout of the packed targets (bit n % 32 of word n / 32)."""
    g = torch.Generator().manual_seed(B * 100 + N)
    flags = torch.rand(B, N, generator=g) > 0.5
    words = L._pack_bits(flags)                              # [B, W] int32
    ref = pack_posbits(flags[:, :, None])[:, 0, :]           # [B, N, T=1] -> [B, 1, W] -> [B, W]
    assert torch.equal(words, ref)


def test_nce_stats_to_loss_and_circulant():
    out4 = torch.tensor([6.0, 3.0, 8.0, 2.0], dtype=torch.float64)
    assert float(L.nce_stats_to_loss(out4)) == pytest.approx((6 / 3 + 8 / 2) / 2)
    assert L.circulant(torch.tensor([0, 1, 2]), dim=0).tolist() == [[0, 1, 2], [2, 0, 1], [1, 2, 0]]


def test_compact_column_layout_and_poison_flag():
    """Ragged columns (NceInputs.col_off / col_src): clip b owns columns [off[b], off[b+1]) in sentence order; a padding
    mask that marks a sentence beyond the list length as real poisons the loss instead of silently dropping it."""
    import torch
    from temporalalignnet_b200 import loss as L, synth
    from tests.helpers import cpu_pos_from_time
    b = synth.make_batch(4, 16, 6, seed=21)
    tpm = torch.from_numpy(b["text_padding_mask"])
    nce = L.prepare_nce_inputs(b["start"], b["end"], tpm, 16, 6, torch.device("cpu"), shard=False,
                               pos_fn=cpu_pos_from_time, compact=True)
    n = [len(s) for s in b["start"]]
    assert nce.compact and nce.C == sum(n) and nce.C_pad == 24
    assert nce.col_off.tolist() == [0] + list(torch.tensor(n).cumsum(0).tolist())
    assert nce.col_src.tolist() == [i * 6 + j for i, k in enumerate(n) for j in range(k)]
    assert bool(nce.col_valid.all()) and not bool(nce.poison)
    x = torch.arange(2 * nce.C * 3, dtype=torch.float32).view(2, nce.C, 3)
    full = nce.scatter_columns(x)
    assert tuple(full.shape) == (2, 24, 3) and torch.equal(full[:, nce.col_src], x)
    assert float(full.sum()) == float(x.sum())
    assert float(L.NceInputs.guard(nce, torch.tensor(1.5))) == 1.5
    tpm2 = tpm.clone()
    short = min(range(4), key=lambda i: n[i])
    if n[short] < 6:
        tpm2[short, n[short]] = False                      # a "real" sentence the lists know nothing about
        nce2 = L.prepare_nce_inputs(b["start"], b["end"], tpm2, 16, 6, torch.device("cpu"), shard=False,
                                    pos_fn=cpu_pos_from_time, compact=True)
        assert bool(nce2.poison) and torch.isnan(nce2.guard(torch.tensor(1.5)))
