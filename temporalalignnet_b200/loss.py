"""`get_loss` and helpers with the reference's signatures (train/loss.py), computed by the fused
similarity + MIL-NCE kernels of libtan_b200.so.

Supported configuration (this round): `args.model in ('init', 'cotrain')` with `sim='cos'`,
`learn_agreement=0`, `loss_threshold=0`, `use_alignability_head=0` -- the `--model init` recipe
(train/readme.md:10).  The self-labelling / threshold / BCE branches (train/loss.py:88-229,
:277-357; BASELINE config 5) raise NotImplementedError instead of silently computing something else.

Closed form (verified bit-exact against the reference on CPU by the oracle tests):
with z = logits / 0.07, valid columns = real sentences, positives = same clip and start <= t < end,
    v[s, r] = LSE_{c valid} z - LSE_{c positive} z      for rows with a positive
    t[s, c] = LSE_r z       - LSE_{r positive} z        for valid columns with a positive
    loss_x  = (mean v + mean t) / 2 ;  loss = (loss_dual + loss_joint) / 2.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch.nn.utils.rnn import pad_sequence

from . import ops
from ._lib import TanError
from .tan_model import LazyLogits


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


class NceInputs:
    """Device-side description of the targets: packed positive bits of the LOCAL clips [B_loc, T, W] and the
    column-valid mask over the GLOBAL columns; optional row_kill / row_sel / col_sel (see include/tan_b200.h)."""

    def __init__(self, posbits, col_valid, N, T, b_off, B_glob, row_kill=None, row_sel=None, col_sel=None):
        self.posbits, self.col_valid, self.N, self.T, self.b_off, self.B_glob = posbits, col_valid, N, T, b_off, B_glob
        self.row_kill, self.row_sel, self.col_sel = row_kill, row_sel, col_sel


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


def prepare_nce_inputs(start_list, end_list, text_padding_mask, T: int, N: int, device, shard: bool,
                       pos_fn=None) -> NceInputs:
    """Targets of the `--model init` recipe: bit (b, t, n) = real sentence and start <= t < end
    (train/loss.py:26-41,:80-85), built on the device by tan_pos_from_time; the column-valid mask
    (~text_padding_mask, :235) is all-gathered over ranks with `shard` so that columns are global."""
    B = len(start_list)
    start, end = padded_times(start_list, end_list, T, N, device)
    valid = (~text_padding_mask.to(device).bool()).to(torch.uint8).contiguous()
    pos_fn = ops.pos_from_time if pos_fn is None else pos_fn      # (tests inject a torch checker on CPU)
    posbits = pos_fn(start.contiguous(), end.contiguous(), valid, B, T, N)
    dist = _dist() if shard else None
    if dist is None:
        return NceInputs(posbits, valid.view(-1), N, T, 0, B)
    W, rank = dist.get_world_size(), dist.get_rank()
    gathered = torch.empty(W * B * N, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(gathered, valid.view(-1))
    return NceInputs(posbits, gathered, N, T, rank * B, W * B)


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
                col_sel=None) -> torch.Tensor:
    """Exp-sums -> loss_x.  row_sums [2, R_local], col_sums [2, S, C] (partial over the local rows).
    Distributed: column sums add across ranks (fixed-shift exp sums); row terms are reduced locally and
    their (sum, count) all-reduced, so every rank returns the global-batch loss."""
    reduce_fn = ops.nce_reduce if reduce_fn is None else reduce_fn
    out4 = torch.zeros(4, dtype=torch.float64, device=row_sums.device)
    S, C = col_sums.shape[1], col_sums.shape[2]
    if dist is None:
        reduce_fn(row_sums, col_sums, out4, S, T, C, row_sel, col_sel)
    else:
        dist.all_reduce(col_sums)
        reduce_fn(row_sums, col_sums, out4, S, T, C, row_sel, col_sel)   # cols now global, identical on every rank
        rows = out4[:2].clone()
        dist.all_reduce(rows)
        out4 = torch.cat((rows, out4[2:]))
    return nce_stats_to_loss(out4)


def nce_loss_one_model(logits, nce: NceInputs, shard: bool) -> torch.Tensor:
    """loss_x for one model's logits (LazyLogits -> fused kernel; tensor -> streaming kernel)."""
    dist = _dist() if shard else None
    if isinstance(logits, LazyLogits):
        vfeat, tfeat = logits.vfeat, logits.tfeat
        B, S, T, d = vfeat.shape
        dev = vfeat.device
        if dist is not None:
            tfeat = gather_text_features(tfeat, logits.shared_text, dist)
        C = tfeat.shape[-2]
        g = ops.sim_geom(B, S, T, C, nce.N, d, nce.b_off)
        row_sums = torch.empty(2, B * S * T, dtype=torch.float32, device=dev)
        col_sums = torch.empty(2, S, C, dtype=torch.float32, device=dev)
        ws = torch.empty(ops.sim_workspace_bytes(g), dtype=torch.uint8, device=dev)
        ops.sim_nce_fwd(vfeat, tfeat, 0 if logits.shared_text else C * d, g, nce.posbits, nce.col_valid, None,
                        row_sums, col_sums, ws, row_kill=nce.row_kill)
    else:
        if logits.dim() != 5:
            raise TanError(f"logits must be [B,S,T,B,N], got {tuple(logits.shape)}")
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
    return finish_loss(row_sums, col_sums, dist, T, row_sel=nce.row_sel, col_sel=nce.col_sel)


def get_loss(input_data, video_seq, text_embed, video_padding_mask, text_padding_mask, logits, args,
             abs_text_pos=None, shard_batch: Optional[bool] = None):
    """train/loss.py:55-422 for the `--model init` recipe.  Same arguments and returned keys
    ('loss', 'loss-dual', 'loss-joint'); every value supports `.item()` (train/main.py:127).

    shard_batch (extension): when torch.distributed is initialised with world_size > 1, each rank
    passes its LOCAL clips and the contrastive matrix spans the global batch (text features and
    targets are all-gathered, column sums all-reduced).  Default: on iff a process group exists."""
    model = getattr(args, "model", "init")
    if model not in ("init", "cotrain"):
        raise TanError(f"unknown args.model {model!r}")
    if getattr(args, "sim", "cos") != "cos":
        raise NotImplementedError("only sim='cos' (the reference default) is implemented")
    if getattr(args, "learn_agreement", 0) or getattr(args, "loss_threshold", 0) > 0 or \
            getattr(args, "use_alignability_head", 0):
        raise NotImplementedError("learn_agreement / loss_threshold / alignability-head losses "
                                  "(train/loss.py:88-229,:277-357) are not implemented in this round")
    logits_dual, logits_joint = logits['logits_dual'], logits['logits_joint']
    B, T, _ = video_seq.shape
    N = text_embed.shape[1]
    device = logits_dual.device
    shard = (_dist() is not None) if shard_batch is None else bool(shard_batch)
    nce = prepare_nce_inputs(input_data['start'], input_data['end'], text_padding_mask, T, N, device, shard)
    loss_dual = nce_loss_one_model(logits_dual, nce, shard)
    loss_joint = nce_loss_one_model(logits_joint, nce, shard)
    loss_dict = {'loss-dual': loss_dual.detach(), 'loss-joint': loss_joint.detach()}
    loss_dict['loss'] = (loss_dual + loss_joint) / 2
    return loss_dict
