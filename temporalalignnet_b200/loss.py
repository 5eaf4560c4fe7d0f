"""`get_loss` and helpers with the reference's signatures (train/loss.py), computed by the fused
similarity + MIL-NCE kernels of libtan_b200.so.

Supported: `args.model in ('init', 'cotrain')`, `sim='cos'`, and every loss flag of the reference:
`learn_agreement` with `temporal_agreement_type in ('i', 'u', 'keep', 'keep-joint')` (train/loss.py:88-229),
`loss_threshold` (:277-304) and `use_alignability_head` (:306-357; BASELINE config 5).

Closed form (verified bit-exact against the reference on CPU by the oracle tests):
with z = logits / 0.07, valid columns = real sentences, positives = same clip and start <= t < end,
    v[s, r] = LSE_{c valid} z - LSE_{c positive} z      for rows with a positive
    t[s, c] = LSE_r z       - LSE_{r positive} z        for valid columns with a positive
    loss_x  = (mean v + mean t) / 2 ;  loss = (loss_dual + loss_joint) / 2.
"""
from __future__ import annotations

import os
import warnings
from typing import Optional

import torch
from torch.nn.utils.rnn import pad_sequence

from . import ops
from ._lib import TanError
from .tan_model import LazyLogits


# sharded mode: verify per call that all ranks pass the same (B_loc, N) (one 32-byte all-reduce on a side stream, no
# main-stream synchronisation); TAN_SHARD_CHECK=0 skips it for loops whose shapes are fixed by construction
SHARD_CHECK = os.environ.get("TAN_SHARD_CHECK", "1") != "0"


def circulant(tensor, dim):
    """train/loss.py:16-23: circulant(tensor([0,1,2]), 0) -> [[0,1,2],[2,0,1],[1,2,0]]."""
    S = tensor.shape[dim]
    flipped = tensor.flip((dim,))
    tmp = torch.cat([flipped, torch.narrow(flipped, dim=dim, start=0, length=S - 1)], dim=dim)
    return tmp.unfold(dim, S, 1).flip((-1,))


def _pad_times(start_list, end_list, T, device):
    start = pad_sequence([torch.as_tensor(i, dtype=torch.float32) for i in start_list], batch_first=True,
                         padding_value=T + 1e2)
    end = pad_sequence([torch.as_tensor(i, dtype=torch.float32) for i in end_list], batch_first=True,
                       padding_value=-1e2)
    return start.to(device, non_blocking=True), end.to(device, non_blocking=True)


def get_mask_from_time(start_list, end_list, num_timestamp, num_text, device='cuda'):
    """train/loss.py:26-41 -> (mask [B,N,T] bool, start [B,N], end [B,N]).  Host-side input
    preparation (python lists -> small tensors); not on the measured path."""
    start, end = _pad_times(start_list, end_list, num_timestamp, device)
    steps = torch.arange(num_timestamp, device=device)[None, None, :]
    mask = (start[:, :, None] <= steps) & (steps < end[:, :, None])
    if mask.shape[1] < num_text:      # fewer sentences than num_text in the whole batch
        mask = torch.nn.functional.pad(mask, (0, 0, 0, num_text - mask.shape[1]))
    return mask, start, end


def get_text_pos(start_list, end_list, device='cuda'):
    """train/loss.py:44-52."""
    start = pad_sequence([torch.as_tensor(i, dtype=torch.float32) for i in start_list], batch_first=True,
                         padding_value=0).to(device, non_blocking=True)
    end = pad_sequence([torch.as_tensor(i, dtype=torch.float32) for i in end_list], batch_first=True,
                       padding_value=0).to(device, non_blocking=True)
    return torch.stack((start, end), dim=-1)


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


_shape_streams = {}


def shard_shapes(B: int, N: int, device):
    """(B_min, B_max, N_min, N_max) over the ranks of the default process group, exchanged on a SIDE stream: the host only
    waits for a 32-byte all-reduce, not for the forward kernels already queued on the main stream."""
    dist = _dist()
    if dist is None:
        return B, B, N, N
    device = torch.device(device)
    st = _shape_streams.get(str(device))
    if st is None:
        st = _shape_streams[str(device)] = torch.cuda.Stream(device=device) if device.type == "cuda" else False
    v = torch.tensor([B, -B, N, -N], dtype=torch.int64)
    if st:
        with torch.cuda.stream(st):
            t = v.to(device, non_blocking=True)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out = t.cpu()                      # synchronises the side stream only
    else:                                      # gloo on CPU tensors (host tests)
        t = v.clone()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out = t
    return int(-out[1]), int(out[0]), int(-out[3]), int(out[2])


def pad_text_to_global(lang_embed: torch.Tensor, lang_padding_mask: torch.Tensor):
    """Sharded training with real data: every rank pads its batch to ITS OWN longest clip (`pad_sequence` in
    data/loader_htm.py:111-129), so N differs between ranks, while the all-gathers of the sharded loss need one
    N.  Call this on the text inputs BEFORE `model(...)` / `get_loss(...)`: pads to the longest N of any rank with
    masked sentences (mask = 1, embedding = the clip's last row, as pad_sequence_by_last, data/loader_htm.py:13-23).
    Returns (lang_embed, lang_padding_mask) unchanged when nothing needs padding."""
    B, N = lang_embed.shape[0], lang_embed.shape[1]
    b_min, b_max, _, n_max = shard_shapes(B, N, lang_embed.device)
    if b_min != b_max:
        raise TanError(f"sharded loss: every rank needs the same number of clips (this rank {B}, ranks have "
                       f"{b_min}..{b_max}); use drop_last=True / a distributed sampler with equal shares")
    if n_max == N:
        return lang_embed, lang_padding_mask
    pad = n_max - N
    emb = torch.cat((lang_embed, lang_embed[:, -1:, :].expand(B, pad, lang_embed.shape[2])), dim=1).contiguous()
    mask = torch.cat((lang_padding_mask, torch.ones(B, pad, dtype=lang_padding_mask.dtype,
                                                    device=lang_padding_mask.device)), dim=1).contiguous()
    return emb, mask


class NceInputs:
    """Device-side description of the targets: packed positive bits of the LOCAL clips [B_loc, T, W] and the
    column-valid mask over the GLOBAL columns; optional row_kill / row_sel / col_sel (see include/tan_b200.h).

    Ragged ("compact") columns: the reference drops the padded sentences before the loss (train/loss.py:235); with
    `col_off` set, the similarity kernels never compute them.  `col_src` [C] maps every compact column to its column
    in the padded [B_glob, N] layout (the layout of the text features), `col_valid` is already compact, `C_pad` =
    B_glob * N.  The lengths come from the HOST lists (`input_data['start']`), like the reference's own padding mask
    (train/main.py:52,:62-65); `poison` is a device flag set when the padding mask marks a sentence BEYOND a clip's
    list length as real -- the compact layout would silently drop it, so the loss is turned into NaN instead."""

    def __init__(self, posbits, col_valid, N, T, b_off, B_glob, row_kill=None, row_sel=None, col_sel=None,
                 col_off=None, col_src=None, poison=None):
        self.posbits, self.col_valid, self.N, self.T, self.b_off, self.B_glob = posbits, col_valid, N, T, b_off, B_glob
        self.row_kill, self.row_sel, self.col_sel = row_kill, row_sel, col_sel
        self.col_off, self.col_src, self.poison = col_off, col_src, poison
        self.C_pad = B_glob * N
        self._stage_ids = {}

    @property
    def compact(self) -> bool:
        return self.col_off is not None

    @property
    def C(self) -> int:
        return int(self.col_src.numel()) if self.compact else self.C_pad

    def geom(self, B_loc, S, T, d):
        return ops.sim_geom(B_loc, S, T, self.C, self.N, d, self.b_off, col_off=self.col_off)

    def compact_features(self, tfeat: torch.Tensor) -> torch.Tensor:
        """[C_pad, d] -> [C, d] or [S, C_pad, d] -> [S, C, d] (bf16 rows gathered by tan_embed_gather_bf16)."""
        if not self.compact:
            return tfeat
        d = tfeat.shape[-1]
        S = 1 if tfeat.dim() == 2 else tfeat.shape[0]
        ids = self._stage_ids.get(S)
        if ids is None:
            ids = (torch.arange(S, device=self.col_src.device, dtype=torch.int64)[:, None] * self.C_pad +
                   self.col_src[None]).reshape(-1).contiguous()
            self._stage_ids[S] = ids
        out = ops.embed_gather(ids, tfeat.reshape(S * self.C_pad, d).contiguous())
        return out.view(self.C, d) if tfeat.dim() == 2 else out.view(S, self.C, d)

    def scatter_columns(self, x: torch.Tensor) -> torch.Tensor:
        """[S, >= C, d] over compact columns -> [S, C_pad, d] over padded columns (zeros at padded sentences)."""
        if not self.compact:
            return x
        out = torch.zeros(x.shape[0], self.C_pad, x.shape[2], dtype=x.dtype, device=x.device)
        out.index_copy_(1, self.col_src, x[:, :self.C])
        return out

    def guard(self, loss: torch.Tensor) -> torch.Tensor:
        return loss if self.poison is None else torch.where(self.poison, torch.full_like(loss, float("nan")), loss)


def padded_times(start_list, end_list, T: int, N: int, device):
    """Python lists -> padded [B, N] start/end (train/loss.py:32-39: missing sentences get start = T+100,
    end = -100, i.e. never positive), via ONE pinned host buffer and one H2D copy."""
    B = len(start_list)
    host = torch.empty(2, B, N, dtype=torch.float32, pin_memory=torch.cuda.is_available())
    host[0].fill_(float(T) + 1e2)
    host[1].fill_(-1e2)
    for b in range(B):
        nb = len(start_list[b])
        if nb > N:
            raise TanError(f"clip {b} has {nb} sentences but text_embed has N={N}")
        if nb:
            host[0, b, :nb] = torch.as_tensor(start_list[b], dtype=torch.float32)
            host[1, b, :nb] = torch.as_tensor(end_list[b], dtype=torch.float32)
    dev = host.to(device, non_blocking=True)
    return dev[0], dev[1]


_ROW_KILL_WARNED = False
COMPACT_COLUMNS = os.environ.get("TAN_COMPACT_COLUMNS", "1") != "0"


EXCH_CAP = 1022      # clips per rank the one-collective shape exchange can describe (2 + EXCH_CAP int32 = 4 KB)
_exch_cache = {}


def shard_exchange(B: int, N: int, n_loc, device):
    """ONE fixed-size collective on the side stream per call: every rank contributes [B, N, n_1 .. n_B] (sentence
    counts of its clips, zero-padded to EXCH_CAP), so that the shape agreement check and the global per-clip
    sentence counts of the ragged-column layout cost a single 4 KB all-gather and a single host wait.
    Returns (b_min, b_max, n_min, n_max, flat list of the sentence counts of all ranks' clips or None)."""
    dist = _dist()
    if dist is None:
        return B, B, N, N, list(n_loc) if n_loc is not None else None
    if B > EXCH_CAP:
        b = shard_shapes(B, N, device)
        return (*b, _host_all_gather_int(n_loc, device) if n_loc is not None else None)
    device = torch.device(device)
    W = dist.get_world_size()
    cuda = device.type == "cuda"
    key = (str(device), W)
    ent = _exch_cache.get(key)
    if ent is None:
        host = torch.zeros(2 + EXCH_CAP, dtype=torch.int32)
        recv = torch.zeros(W * (2 + EXCH_CAP), dtype=torch.int32)
        if cuda:
            host, recv = host.pin_memory(), recv.pin_memory()
        # torch.empty on purpose: a zero-fill would be a kernel on the MAIN stream, which may run (after whatever is
        # queued there) in the middle of the side stream's copy -> gather -> copy sequence below
        ent = _exch_cache[key] = (host, recv, torch.empty(2 + EXCH_CAP, dtype=torch.int32, device=device),
                                  torch.empty(W * (2 + EXCH_CAP), dtype=torch.int32, device=device))
    host, recv, d_send, d_recv = ent
    host.zero_()
    host[0], host[1] = B, N
    if n_loc is not None:
        host[2:2 + B] = torch.as_tensor(n_loc, dtype=torch.int32)
    st = _shape_streams.get(str(device))
    if st is None:
        st = _shape_streams[str(device)] = torch.cuda.Stream(device=device) if cuda else False
    if st:
        with torch.cuda.stream(st):
            d_send.copy_(host, non_blocking=True)
            dist.all_gather_into_tensor(d_recv, d_send)
            recv.copy_(d_recv, non_blocking=True)
        st.synchronize()                       # the side stream only: the main stream's queued kernels keep running
        got = recv.view(W, 2 + EXCH_CAP)
    else:                                      # gloo on CPU tensors (host tests)
        outs = [torch.empty_like(host) for _ in range(W)]
        dist.all_gather(outs, host.clone())
        got = torch.stack(outs)
    bs, ns = got[:, 0].tolist(), got[:, 1].tolist()
    n_glob = None
    if n_loc is not None and min(bs) == max(bs):
        n_glob = got[:, 2:2 + B].reshape(-1).tolist()
    return min(bs), max(bs), min(ns), max(ns), n_glob


def _host_all_gather_int(vec, device):
    """All-gather a short list of ints over the ranks through a SIDE stream (see shard_shapes) -> flat python list."""
    dist = _dist()
    device = torch.device(device)
    t = torch.tensor(vec, dtype=torch.int64)
    W = dist.get_world_size()
    st = _shape_streams.get(str(device))
    if st is None:
        st = _shape_streams[str(device)] = torch.cuda.Stream(device=device) if device.type == "cuda" else False
    if st:
        with torch.cuda.stream(st):
            td = t.to(device, non_blocking=True)
            out = torch.empty(W * td.numel(), dtype=torch.int64, device=device)
            dist.all_gather_into_tensor(out, td)
            return out.cpu().tolist()
    outs = [torch.empty_like(t) for _ in range(W)]
    dist.all_gather(outs, t)
    return torch.cat(outs).tolist()


_compact_cache = {}


def prepare_nce_inputs(start_list, end_list, text_padding_mask, T: int, N: int, device, shard: bool,
                       pos_fn=None, padded=None, compact: bool = False, n_glob=None) -> NceInputs:
    """Targets of the `--model init` recipe: bit (b, t, n) = real sentence and start <= t < end
    (train/loss.py:26-41,:80-85), built on the device by tan_pos_from_time; the column-valid mask
    (~text_padding_mask, :235) is all-gathered over ranks with `shard` so that columns are global.
    compact: ragged columns from the host-side sentence counts (see NceInputs)."""
    B = len(start_list)
    if padded is not None:                      # data.collate_fn already padded the times (pinned -> one async copy)
        start, end = (t.to(device, non_blocking=True).float() for t in padded)
        if start.shape[1] < N:                  # sharded runs pad the text to the longest N of any rank
            start = torch.nn.functional.pad(start, (0, N - start.shape[1]), value=float(T) + 1e2)
            end = torch.nn.functional.pad(end, (0, N - end.shape[1]), value=-1e2)
    else:
        start, end = padded_times(start_list, end_list, T, N, device)
    valid = (~text_padding_mask.to(device).bool()).to(torch.uint8).contiguous()
    pos_fn = ops.pos_from_time if pos_fn is None else pos_fn      # (tests inject a torch checker on CPU)
    posbits = pos_fn(start.contiguous(), end.contiguous(), valid, B, T, N)
    dist = _dist() if shard else None
    if dist is None:
        valid_g, b_off, B_glob = valid.view(-1), 0, B
    else:
        W, rank = dist.get_world_size(), dist.get_rank()
        valid_g = torch.empty(W * B * N, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(valid_g, valid.view(-1))
        b_off, B_glob = rank * B, W * B
    if not (compact and COMPACT_COLUMNS):
        return NceInputs(posbits, valid_g, N, T, b_off, B_glob)
    import numpy as np
    if n_glob is None:
        n_loc = [min(len(s_), N) for s_ in start_list]
        n_glob = n_loc if dist is None else _host_all_gather_int(n_loc, device)
    n_glob = [min(int(n_), N) for n_ in n_glob]
    # the layout tensors only depend on (sentence counts, N, device): batches that repeat them (fixed-length
    # loaders, benchmarks) reuse the device copies
    ckey = (tuple(n_glob), N, str(device))
    ent = _compact_cache.get(ckey)
    if ent is None:
        off = np.zeros(B_glob + 1, np.int32)
        np.cumsum(np.asarray(n_glob, np.int64), out=off[1:])
        src = np.concatenate([b * N + np.arange(n, dtype=np.int64) for b, n in enumerate(n_glob)]) if off[-1] else \
            np.zeros(0, np.int64)
        pin = torch.cuda.is_available() and torch.device(device).type == "cuda"
        col_off = torch.from_numpy(off)
        col_src = torch.from_numpy(src)
        if pin:
            col_off, col_src = col_off.pin_memory(), col_src.pin_memory()
        col_off, col_src = col_off.to(device, non_blocking=True), col_src.to(device, non_blocking=True)
        covered = torch.zeros(B_glob * N, dtype=torch.bool, device=device)
        covered[col_src] = True
        if len(_compact_cache) > 8:
            _compact_cache.clear()
        ent = _compact_cache[ckey] = (col_off, col_src, covered)
    col_off, col_src, covered = ent
    poison = (valid_g.bool() & ~covered).any()
    return NceInputs(posbits, valid_g.index_select(0, col_src).contiguous(), N, T, b_off, B_glob, col_off=col_off,
                     col_src=col_src, poison=poison)


def nce_stats_to_loss(out4: torch.Tensor) -> torch.Tensor:
    """(sum_v, n_v, sum_t, n_t) fp64 -> loss_x = (mean v + mean t) / 2 (train/loss.py:256) as fp32."""
    return ((out4[0] / out4[1] + out4[2] / out4[3]) * 0.5).float()


def gather_text_features(tfeat: torch.Tensor, shared_text: bool, dist) -> torch.Tensor:
    """The one exchange step of the path: every rank needs every clip's L2-normalised
    text features.  [B_loc*N, d] -> [B*N, d] (dual) or [S, B_loc*N, d] -> [S, B*N, d] (joint), rank-major
    so that global column c = (rank*B_loc + b)*N + n."""
    W = dist.get_world_size()
    d = tfeat.shape[-1]
    if shared_text:
        full = torch.empty(W * tfeat.shape[0], d, dtype=tfeat.dtype, device=tfeat.device)
        dist.all_gather_into_tensor(full, tfeat.contiguous())
        return full
    S = tfeat.shape[0]
    g = torch.empty(W * S, tfeat.shape[1], d, dtype=tfeat.dtype, device=tfeat.device)   # concat along dim 0
    dist.all_gather_into_tensor(g, tfeat.contiguous())
    return g.view(W, S, tfeat.shape[1], d).permute(1, 0, 2, 3).reshape(S, W * tfeat.shape[1], d).contiguous()


def finish_loss(row_sums: torch.Tensor, col_sums: torch.Tensor, dist, T: int, reduce_fn=None, row_sel=None,
                col_sel=None, reduce_cols: bool = True) -> torch.Tensor:
    """Exp-sums -> loss_x.  row_sums [2, R_local], col_sums [2, S, C] (partial over the local rows unless
    reduce_cols is False).  Distributed: column sums add across ranks (fixed-shift exp sums); row terms are
    reduced locally and their (sum, count) all-reduced, so every rank returns the global-batch loss."""
    reduce_fn = ops.nce_reduce if reduce_fn is None else reduce_fn
    out4 = torch.zeros(4, dtype=torch.float64, device=row_sums.device)
    S, C = col_sums.shape[1], col_sums.shape[2]
    if dist is None:
        reduce_fn(row_sums, col_sums, out4, S, T, C, row_sel, col_sel)
    else:
        if reduce_cols:
            dist.all_reduce(col_sums)
        reduce_fn(row_sums, col_sums, out4, S, T, C, row_sel, col_sel)   # cols now global, identical on every rank
        rows = out4[:2].clone()
        dist.all_reduce(rows)
        out4 = torch.cat((rows, out4[2:]))
    return nce_stats_to_loss(out4)


def nce_sums_one_model(logits, nce: NceInputs, shard: bool):
    """(row_sums [2, B_loc*S*T], col_sums [2, S, C] summed over ALL ranks' rows, T) for one model's logits
    (LazyLogits -> fused kernel; tensor -> streaming kernel)."""
    dist = _dist() if shard else None
    if isinstance(logits, LazyLogits):
        vfeat, tfeat = logits.vfeat, logits.tfeat
        B, S, T, d = vfeat.shape
        dev = vfeat.device
        if dist is not None:
            tfeat = gather_text_features(tfeat, logits.shared_text, dist)
        tfeat = nce.compact_features(tfeat)
        C = tfeat.shape[-2]
        g = nce.geom(B, S, T, d)
        row_sums = torch.empty(2, B * S * T, dtype=torch.float32, device=dev)
        col_sums = torch.empty(2, S, C, dtype=torch.float32, device=dev)
        ws = torch.empty(ops.sim_workspace_bytes(g), dtype=torch.uint8, device=dev)
        ops.sim_nce_fwd(vfeat, tfeat, 0 if logits.shared_text else C * d, g, nce.posbits, nce.col_valid, None,
                        row_sums, col_sums, ws, row_kill=nce.row_kill)
    else:
        if logits.dim() != 5:
            raise TanError(f"logits must be [B,S,T,B,N], got {tuple(logits.shape)}")
        if nce.compact:
            raise TanError("materialised logits keep the padded [B,S,T,B,N] layout: prepare the targets with compact=False")
        B, S, T, B2, N = logits.shape
        dev = logits.device
        if not logits.is_cuda:
            raise TanError("get_loss runs on a CUDA (sm_100a) device only; there is no CPU path")
        x = logits.detach()
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        x = x.contiguous()
        C = B2 * N
        g = ops.sim_geom(B, S, T, C, N, 1, nce.b_off)
        row_sums = torch.empty(2, B * S * T, dtype=torch.float32, device=dev)
        col_sums = torch.empty(2, S, C, dtype=torch.float32, device=dev)
        ws = torch.empty(ops.sim_workspace_bytes(g), dtype=torch.uint8, device=dev)
        ops.nce_from_logits(x, g, nce.posbits, nce.col_valid, row_sums, col_sums, ws, row_kill=nce.row_kill)
    if dist is not None:
        dist.all_reduce(col_sums)
    return row_sums, col_sums, T


def nce_loss_one_model(logits, nce: NceInputs, shard: bool) -> torch.Tensor:
    """loss_x for one model's logits."""
    dist = _dist() if shard else None
    row_sums, col_sums, T = nce_sums_one_model(logits, nce, shard)
    return nce.guard(finish_loss(row_sums, col_sums, dist, T, row_sel=nce.row_sel, col_sel=nce.col_sel, reduce_cols=False))


def pack_text_features(td: torch.Tensor, tj: torch.Tensor) -> torch.Tensor:
    """[1 + Sj, B_loc*N, d]: the dual model's text features stacked on the joint model's stages.  The forward
    allocates them as adjacent slices of one buffer (TemporalAligner._forward_impl), so this is normally a view."""
    Sj, BN, d = tj.shape
    if (td.is_contiguous() and tj.is_contiguous() and td.untyped_storage().data_ptr() == tj.untyped_storage().data_ptr()
            and tj.storage_offset() == td.storage_offset() + BN * d):
        return torch.as_strided(td, (1 + Sj, BN, d), (BN * d, d, 1))
    return torch.cat((td[None], tj), dim=0).contiguous()


# measured on 2 x B200 (profiles/r02f): the exchange alone takes 0.18 ms as one all-gather + permuting copy and 0.62 ms
# as a coalesced group of per-stage all-gathers that write the stage-major layout directly -- so the group is opt-in
_COALESCE = os.environ.get("TAN_GATHER_COALESCED", "0") == "1"


def exchange_text_features(packed: torch.Tensor, dist) -> torch.Tensor:
    """The one exchange step of the path: [1 + Sj, B_loc*N, d] per rank -> stage-major
    [1 + Sj, W*B_loc*N, d] (global column c = rank * B_loc*N + local column), the layout the similarity kernel's 2-D
    TMA map reads: one all-gather into a rank-major buffer + a permuting copy (or, TAN_GATHER_COALESCED=1, a
    coalesced group of per-stage all-gathers that writes the layout directly; measured slower)."""
    W = dist.get_world_size()
    S1, BN, d = packed.shape
    full = torch.empty(S1, W * BN, d, dtype=packed.dtype, device=packed.device)
    if _COALESCE and packed.is_cuda and hasattr(dist, "_coalescing_manager"):
        try:
            with dist._coalescing_manager(device=packed.device, async_ops=False):
                for s in range(S1):
                    dist.all_gather_into_tensor(full[s], packed[s])
            return full
        except Exception:                     # backend without coalescing support: fall through
            pass
    gathered = torch.empty(W * S1, BN, d, dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(gathered, packed)
    return gathered.view(W, S1, BN, d).permute(1, 0, 2, 3).reshape(S1, W * BN, d).contiguous()


def sim_pair_sums(vd, vj, full, nce: NceInputs):
    """Fused similarity + exp-sum kernels of both models on the local rows x the columns of `full` [1 + Sj, C, d].
    Returns (rs_d, cs_d, rs_j, cs_j, cols): `cols` is the flat buffer holding both column-sum tensors (one
    all-reduce)."""
    B, Sd, T, d = vd.shape
    Sj = vj.shape[1]
    dev = vd.device
    full = nce.compact_features(full)
    C = full.shape[1]
    cols = torch.empty(2 * (Sd + Sj) * C, dtype=torch.float32, device=dev)
    cs_d, cs_j = cols[:2 * Sd * C].view(2, Sd, C), cols[2 * Sd * C:].view(2, Sj, C)
    rs_d = torch.empty(2, B * Sd * T, dtype=torch.float32, device=dev)
    rs_j = torch.empty(2, B * Sj * T, dtype=torch.float32, device=dev)
    g_d = nce.geom(B, Sd, T, d)
    g_j = nce.geom(B, Sj, T, d)
    ws = torch.empty(max(ops.sim_workspace_bytes(g_d), ops.sim_workspace_bytes(g_j)), dtype=torch.uint8, device=dev)
    ops.sim_nce_fwd(vd, full[0], 0, g_d, nce.posbits, nce.col_valid, None, rs_d, cs_d, ws, row_kill=nce.row_kill)
    ops.sim_nce_fwd(vj, full[1:], C * d, g_j, nce.posbits, nce.col_valid, None, rs_j, cs_j, ws, row_kill=nce.row_kill)
    return rs_d, cs_d, rs_j, cs_j, cols


def nce_losses_pair(logits_dual, logits_joint, nce: NceInputs, shard: bool):
    """(loss_dual, loss_joint).  Sharded + fused path: the exchange of BOTH models is batched into one gather of the
    text features (dual model's stacked on the joint model's stages), one all-reduce (column sums) and one tiny
    all-reduce (row sums / counts) -- the collectives are latency-bound at these sizes (1-6 MB per rank), so their
    number, not their volume, is what a step pays for."""
    dist = _dist() if shard else None
    if dist is None or not (isinstance(logits_dual, LazyLogits) and isinstance(logits_joint, LazyLogits)):
        return nce_loss_one_model(logits_dual, nce, shard), nce_loss_one_model(logits_joint, nce, shard)
    vd, vj = logits_dual.vfeat, logits_joint.vfeat
    Sd, T, Sj = vd.shape[1], vd.shape[2], vj.shape[1]
    full = exchange_text_features(pack_text_features(logits_dual.tfeat, logits_joint.tfeat), dist)
    rs_d, cs_d, rs_j, cs_j, cols = sim_pair_sums(vd, vj, full, nce)
    C = cs_d.shape[2]
    dist.all_reduce(cols)
    out8 = torch.zeros(8, dtype=torch.float64, device=vd.device)
    ops.nce_reduce(rs_d, cs_d, out8[0:4], Sd, T, C, nce.row_sel, nce.col_sel)
    ops.nce_reduce(rs_j, cs_j, out8[4:8], Sj, T, C, nce.row_sel, nce.col_sel)
    rows = torch.cat((out8[0:2], out8[4:6]))
    dist.all_reduce(rows)
    loss_d = nce.guard(nce_stats_to_loss(torch.cat((rows[0:2], out8[2:4]))))
    loss_j = nce.guard(nce_stats_to_loss(torch.cat((rows[2:4], out8[6:8]))))
    return loss_d, loss_j


# ------------------------------------------------------------------------------------------------
# helpers of the agreement / threshold / alignability branches: tiny [B*N]-sized device vectors, no host sync
# ------------------------------------------------------------------------------------------------
def _gather_flat(x: torch.Tensor, dist) -> torch.Tensor:
    """[B_loc*N] per rank -> [B*N] in global column order (rank-major)."""
    x = x.contiguous().view(-1)
    if dist is None:
        return x
    out = torch.empty(dist.get_world_size() * x.numel(), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x)
    return out


def masked_quantile(x: torch.Tensor, valid: torch.Tensor, q: float) -> torch.Tensor:
    """torch.quantile(x[valid], q) (linear interpolation) without the boolean-index host sync."""
    xs, _ = torch.sort(torch.where(valid, x.float(), torch.full_like(x, float("inf"), dtype=torch.float32)))
    m1 = (valid.sum() - 1).clamp(min=0)
    pos = q * m1.to(torch.float32)
    lo = pos.floor()
    w = pos - lo
    lo_i = lo.long()
    hi_i = torch.minimum(lo_i + 1, m1)
    a, b = xs[lo_i], xs[hi_i]
    return torch.where(w > 0, a + (b - a) * w, a)


def _masked_standardise(x: torch.Tensor, valid: torch.Tensor) -> torch.Tensor:
    """(x - mean) / std (unbiased) over the valid entries (train/loss.py:281,:283)."""
    v = valid.float()
    m = v.sum()
    mean = (x * v).sum() / m
    var = (((x - mean) ** 2) * v).sum() / (m - 1)
    return (x - mean) / var.sqrt()


def _own_last_stage(lg, B: int, T: int, N: int, b_off: int) -> torch.Tensor:
    """Own-clip cosines of the LAST stage [B_loc, T, N] fp32 (torch.diagonal(...)[:, -1], train/loss.py:92-105)."""
    if isinstance(lg, LazyLogits):
        Bv, S, Tv, d = lg.vfeat.shape
        return ops.own_clip_sim(lg.vfeat, lg.tfeat, lg.shared_text, Bv, S, Tv, N, d, s_first=S - 1, s_count=1)[:, 0]
    idx = torch.arange(B, device=lg.device)
    return lg.detach()[idx, -1, :, b_off + idx, :].float().contiguous()


def _pack_bits(flags_bn: torch.Tensor) -> torch.Tensor:
    """[B, N] bool -> [B, W] int32 words (bit n % 32 of word n / 32)."""
    B, N = flags_bn.shape
    W = (N + 31) // 32
    f = torch.zeros(B, W * 32, dtype=torch.int64, device=flags_bn.device)
    f[:, :N] = flags_bn.to(torch.int64)
    words = (f.view(B, W, 32) << torch.arange(32, device=f.device, dtype=torch.int64)).sum(-1)
    return torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)


def get_loss(input_data, video_seq, text_embed, video_padding_mask, text_padding_mask, logits, args,
             abs_text_pos=None, shard_batch: Optional[bool] = None):
    """train/loss.py:55-422.  Same arguments and returned keys ('loss', 'loss-dual', 'loss-joint' and, with the
    respective flags, 'confidence-ratio', 'iou-threshold', 'loss-dual-all', 'loss-joint-all', 'loss-total',
    'loss-joint-bce', 'alignability_top1'); every value supports `.item()` (train/main.py:127).

    The [B,S,T,B,N] passes run in the fused similarity+NCE kernel (or the streaming kernel for plain tensors);
    agreement self-labelling (--learn_agreement) runs on the own-clip blocks (tan_own_clip_sim, tan_agree_scan,
    tan_agree_targets); what remains in torch are elementwise / sort operations on [B*N]-sized vectors
    (quantiles, standardisation, the BCE of the alignability head), without host synchronisation.

    shard_batch (extension): when torch.distributed is initialised with world_size > 1, each rank passes its
    LOCAL clips and the contrastive matrix, the quantiles and the BCE span the global batch.  Default: on iff a
    process group exists."""
    model = getattr(args, "model", "init")
    if model not in ("init", "cotrain"):
        raise TanError(f"unknown args.model {model!r}")
    if getattr(args, "sim", "cos") != "cos":
        raise NotImplementedError("only sim='cos' (the reference default) is implemented")
    learn = bool(getattr(args, "learn_agreement", 0))
    thr = float(getattr(args, "loss_threshold", 0.0) or 0.0)
    head = bool(getattr(args, "use_alignability_head", 0))
    kind = getattr(args, "temporal_agreement_type", "keep")
    if learn and kind not in ops.AGREE_KINDS:
        raise TanError(f"unknown temporal_agreement_type {kind!r}")
    logits_dual, logits_joint = logits['logits_dual'], logits['logits_joint']
    B, T, _ = video_seq.shape
    N = text_embed.shape[1]
    device = logits_dual.device
    shard = (_dist() is not None) if shard_batch is None else bool(shard_batch)
    dist = _dist() if shard else None
    n_glob = None
    will_compact = (COMPACT_COLUMNS and not (learn or thr > 0 or head) and N <= 64 and isinstance(logits_dual, LazyLogits)
                    and isinstance(logits_joint, LazyLogits))
    hint = input_data.get('n_sentences_global') if isinstance(input_data, dict) else None
    if dist is not None and hint is not None:
        # a loader that slices ONE global batch over the ranks knows every clip's sentence count: no exchange, no
        # host wait in the step (the per-step 4 KB gather below is a host-side rendezvous of all ranks)
        if len(hint) != dist.get_world_size() * B:
            raise TanError(f"input_data['n_sentences_global'] has {len(hint)} entries, expected world_size * B = "
                           f"{dist.get_world_size() * B}")
        n_glob = [min(int(n_), N) for n_ in hint] if will_compact else None
    elif dist is not None and (SHARD_CHECK or will_compact):
        # the all-gathers below assume one (B_loc, N) on every rank; with real data N is each rank's own
        # pad_sequence length -- fail loudly instead of hanging in NCCL or mis-indexing columns.  The same 4 KB
        # exchange carries every clip's sentence count for the ragged-column layout.
        n_loc = [min(len(s_), N) for s_ in input_data['start']] if will_compact else None
        b_min, b_max, n_min, n_max, n_glob = shard_exchange(B, N, n_loc, device)
        if b_min != b_max or n_min != n_max:          # the same verdict on every rank: nobody enters a collective alone
            raise TanError(f"sharded get_loss: ranks disagree on the batch shape (this rank B={B}, N={N}; ranks have "
                           f"B {b_min}..{b_max}, N {n_min}..{n_max}).  Pad the text inputs with "
                           "temporalalignnet_b200.loss.pad_text_to_global(lang_embed, lang_padding_mask) before the "
                           "forward, and give every rank the same number of clips.")
    padded = (input_data['start_pad'], input_data['end_pad']) if 'start_pad' in input_data else None
    # ragged columns (padded sentences never computed): the plain recipe on fused logits; the flag branches index
    # [B*N]-shaped vectors by padded column and the N > 64 gradient path has no ragged variant
    nce = prepare_nce_inputs(input_data['start'], input_data['end'], text_padding_mask, T, N, device, shard, padded=padded,
                             compact=will_compact, n_glob=n_glob)
    loss_dict = {}
    # training step: the forward ran with a tape (model.enable_autograd) -> the returned loss carries ONE autograd
    # node whose backward is the hand-written backward pass (train.py)
    tape = getattr(logits_dual, "tape", None) if isinstance(logits_dual, LazyLogits) else None
    want_grad = tape is not None and torch.is_grad_enabled()
    if want_grad and not (learn or thr > 0 or head):
        from . import train
        rs_d, cs_d, _ = nce_sums_one_model(logits_dual, nce, shard)
        rs_j, cs_j, _ = nce_sums_one_model(logits_joint, nce, shard)
        loss_dual = nce.guard(finish_loss(rs_d, cs_d, dist, T, reduce_cols=False))
        loss_joint = nce.guard(finish_loss(rs_j, cs_j, dist, T, reduce_cols=False))
        loss_dict['loss-dual'], loss_dict['loss-joint'] = loss_dual.detach(), loss_joint.detach()
        loss_dict['loss'] = train.attach_autograd((loss_dual + loss_joint) / 2, tape,
                                                  train.SimCtx(logits_dual, rs_d, cs_d, nce),
                                                  train.SimCtx(logits_joint, rs_j, cs_j, nce), 1.0, dist)
        return loss_dict
    if not (learn or thr > 0 or head):
        loss_dual, loss_joint = nce_losses_pair(logits_dual, logits_joint, nce, shard)
        loss_dict['loss-dual'], loss_dict['loss-joint'] = loss_dual.detach(), loss_joint.detach()
        loss_dict['loss'] = (loss_dual + loss_joint) / 2
        return loss_dict

    tpm_b = text_padding_mask.to(device).bool().view(B, N)
    tpm_u8 = tpm_b.to(torch.uint8).contiguous()
    vpm_u8 = None
    if video_padding_mask is not None:
        vpm_u8 = video_padding_mask.to(device).bool().to(torch.uint8).contiguous()
    valid_g = nce.col_valid.bool()                                           # [C] global
    md = mj = None
    if learn:
        # ---- agreement self-labelling (train/loss.py:88-229) -------------------------------------
        if model == "cotrain":
            src_d, src_j = logits['ema-logits_dual'], logits['ema-logits_joint']
        else:
            src_d, src_j = logits_dual, logits_joint
        fill = model != "cotrain"          # the reference fills the loss logits in place only in `init` mode
        own_j = _own_last_stage(src_j, B, T, N, nce.b_off)
        own_d = _own_last_stage(src_d, B, T, N, nce.b_off)
        win_j, mean_j, max_j = ops.agree_scan(own_j, nce.posbits, vpm_u8, tpm_u8, B, T, N, fill_max=fill)
        win_d, mean_d, max_d = ops.agree_scan(own_d, nce.posbits, vpm_u8, tpm_u8, B, T, N, fill_max=fill)
        inter = (torch.minimum(win_j[..., 1], win_d[..., 1]) - torch.maximum(win_j[..., 0], win_d[..., 0])).clamp(min=0)
        union = (win_j[..., 1] - win_j[..., 0]) + (win_d[..., 1] - win_d[..., 0]) - inter
        iou = inter.float() / union.float().clamp(min=1e-5)
        mean_d_g, mean_j_g = _gather_flat(mean_d, dist), _gather_flat(mean_j, dist)
        q_d = masked_quantile(mean_d_g, valid_g, 0.3)
        q_j = masked_quantile(mean_j_g, valid_g, 0.3)
        conf_iou = iou >= 0.5
        conf = (mean_d >= q_d) & (mean_j >= q_j) & conf_iou
        replace = conf if kind in ("i", "u") else conf_iou
        nce.posbits = ops.agree_targets(nce.posbits, win_j, win_d, replace.to(torch.uint8).contiguous(), B, T, N, kind)
        if fill and vpm_u8 is not None:
            nce.row_kill = vpm_u8
            global _ROW_KILL_WARNED
            if not _ROW_KILL_WARNED:
                _ROW_KILL_WARNED = True
                warnings.warn("get_loss(model='init', learn_agreement=1) with a video padding mask: padded frames lose "
                              "their own-clip entries and drop out of the row mean here; the reference keeps a padded "
                              "frame that carries a positive with a ~6e4 loss term (a view side effect of "
                              "train/loss.py:96-100, DESIGN.md section 2).  Loss values differ from reference runs in "
                              "that configuration only.", stacklevel=2)
        conf_g = _gather_flat(conf, dist)
        loss_dict['confidence-ratio'] = (conf_g & valid_g).float().sum() / valid_g.float().sum()
        loss_dict['iou-threshold'] = torch.tensor(0.5, device=device)
        if fill:
            md, mj = max_d, max_j
    rs_d, cs_d, _ = nce_sums_one_model(logits_dual, nce, shard)
    rs_j, cs_j, _ = nce_sums_one_model(logits_joint, nce, shard)
    loss_dual = finish_loss(rs_d, cs_d, dist, T, reduce_cols=False)
    loss_joint = finish_loss(rs_j, cs_j, dist, T, reduce_cols=False)
    loss_dict['loss-dual'], loss_dict['loss-joint'] = loss_dual.detach(), loss_joint.detach()
    loss_th = None
    loss_bce = None
    row_sel = col_sel = None
    bce_dx = None
    if thr > 0 or head:
        # ---- keep the most alignable sentences (train/loss.py:277-304) -----------------------------
        if md is None:
            own_d = _own_last_stage(logits_dual, B, T, N, nce.b_off)
            own_j = _own_last_stage(logits_joint, B, T, N, nce.b_off)
            _, _, md = ops.agree_scan(own_d, nce.posbits, vpm_u8, tpm_u8, B, T, N, fill_max=False)
            _, _, mj = ops.agree_scan(own_j, nce.posbits, vpm_u8, tpm_u8, B, T, N, fill_max=False)
        md_g, mj_g = _gather_flat(md, dist), _gather_flat(mj, dist)
        metric = -(_masked_standardise(md_g, valid_g) + _masked_standardise(mj_g, valid_g))
        keep_g = (metric <= masked_quantile(metric, valid_g, thr)) & valid_g
        if thr > 0:
            loss_dict['loss-dual-all'], loss_dict['loss-joint-all'] = loss_dual.detach(), loss_joint.detach()
            keep_loc = keep_g.view(-1, N)[nce.b_off:nce.b_off + B]
            row_sel = ((nce.posbits & _pack_bits(keep_loc)[:, None, :]) != 0).any(-1).to(torch.uint8).contiguous()
            col_sel = keep_g.to(torch.uint8).contiguous()
            loss_dual_th = finish_loss(rs_d, cs_d, dist, T, row_sel=row_sel, col_sel=col_sel, reduce_cols=False)
            loss_joint_th = finish_loss(rs_j, cs_j, dist, T, row_sel=row_sel, col_sel=col_sel, reduce_cols=False)
            loss_dict['loss-dual'], loss_dict['loss-joint'] = loss_dual_th.detach(), loss_joint_th.detach()
            loss_th = (loss_dual_th + loss_joint_th) / 2
        if head:
            # ---- alignability BCE (train/loss.py:306-357) ------------------------------------------
            label = torch.full_like(md_g, 2.0)
            qd5, qj5 = masked_quantile(md_g, valid_g, 0.5), masked_quantile(mj_g, valid_g, 0.5)
            label = torch.where((md_g > qd5) & (mj_g > qj5), torch.ones_like(label), label)
            label = torch.where((md_g < qd5) & (mj_g < qj5), torch.zeros_like(label), label)
            if abs_text_pos is not None:
                centre = _gather_flat(abs_text_pos.to(device).float().mean(-1), dist)
                label = torch.where((centre < 0.2) | (centre > 0.8), torch.zeros_like(label), label)
            has_pos = cs_d[1].amax(0) > 0                                   # real sentences with a positive (global)
            xj = _gather_flat(logits['joint_logits_alignability'][:, 2, :, 0].float(), dist)
            sel = ((label != 2.0) & valid_g & has_pos).float()
            n_sel = sel.sum()
            y = torch.where(label == 1.0, torch.ones_like(label), torch.zeros_like(label))
            pw = 1.0 / ((y * sel).sum() / n_sel) - 1.0
            sp = torch.nn.functional.softplus
            loss_bce = ((pw * y * sp(-xj) + (1 - y) * sp(xj)) * sel).sum() / n_sel
            if want_grad:     # d loss_bce / d x of the local sentences (pos_weight and labels carry no gradient)
                dxj = sel * ((1 - y) * torch.sigmoid(xj) - pw * y * torch.sigmoid(-xj)) / n_sel
                bce_dx = dxj.view(-1, N)[nce.b_off:nce.b_off + B].contiguous()
            loss_dict['loss-joint-bce'] = loss_bce.detach()
            loss_dict['alignability_top1'] = ((((xj > 0).float() == y).float()) * sel).sum() / n_sel
    nce_weight = 0.0 if getattr(args, "optim_policy", "default") == "bce" else 1.0
    if thr > 0:
        loss_dict['loss-total'] = ((loss_dual + loss_joint) / 2).detach()
        loss = loss_th
    else:
        loss = (loss_dual + loss_joint) / 2
    if head:
        loss = loss * nce_weight + loss_bce
    if want_grad:
        from . import train
        loss = train.attach_autograd(loss, tape, train.SimCtx(logits_dual, rs_d, cs_d, nce, row_sel, col_sel),
                                     train.SimCtx(logits_joint, rs_j, cs_j, nce, row_sel, col_sel), nce_weight, dist,
                                     bce_dx)
    loss_dict['loss'] = loss
    return loss_dict
