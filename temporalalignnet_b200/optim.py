"""Optimizer step of the (co-)training loop: per-parameter gradient clipping
(utils/train_utils.py:3-13), AdamW (train/main.py:397) and the EMA update of the target network
(model/tan_model.py:340-344) as multi-tensor kernels of libtan_b200.so -- two launches for ALL parameters, no host
synchronisation (the reference does one `.item()` per parameter for the clipping and ~5 small kernels per parameter
for AdamW and the EMA).

    opt = FusedAdamW(optim_policy(model, args), lr=args.lr, weight_decay=args.wd, clip_grad=args.clip_grad,
                     ema=(model.target.parameters(), model.online.parameters(), model.m))     # ema: cotrain only
    ...
    loss.backward()
    opt.step()            # train/main.py:113-122 in one call (clip + AdamW + _momentum_update)
    opt.zero_grad()

`FusedAdamW` is a `torch.optim.Optimizer` (param groups, state_dict, lr schedulers and GradScaler.step work as
usual); an unclipped step is bit-identical to `torch.optim.AdamW(foreach=True)` on fp32 parameters.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional

import torch

from . import _lib
from ._lib import OptimTensor, TanError, check, lib

CHUNK = 16384          # elements per CTA


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class _Plan:
    """Chunk tables of a fixed list of tensors (device int arrays) + a pinned host staging table that is refilled
    and copied to the device every step (gradient pointers and step sizes change)."""

    def __init__(self, numels: List[int], device):
        chunk_tensor, chunk_start, first = [], [], [0]
        for t, n in enumerate(numels):
            for s in range(0, max(n, 1), CHUNK):
                chunk_tensor.append(t)
                chunk_start.append(s)
            first.append(len(chunk_tensor))
        self.n_tensors, self.n_chunks = len(numels), len(chunk_tensor)
        self.chunk_tensor = torch.tensor(chunk_tensor, dtype=torch.int32, device=device)
        self.chunk_start = torch.tensor(chunk_start, dtype=torch.int64, device=device)
        self.first = torch.tensor(first, dtype=torch.int32, device=device)
        self.partial = torch.empty(self.n_chunks, dtype=torch.float32, device=device)
        self.norms = torch.zeros(self.n_tensors, dtype=torch.float32, device=device)
        nbytes = C.sizeof(OptimTensor) * self.n_tensors
        # the host table is refilled every step while earlier asynchronous copies may still be queued: a ring of
        # pinned staging buffers, each guarded by the event of its last copy (waiting on a 3-steps-old event never
        # blocks in practice; no device synchronisation)
        self.hosts = [torch.empty(nbytes, dtype=torch.uint8, pin_memory=True) for _ in range(3)]
        self.events = [None, None, None]
        self.slot = 0
        self.dev = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.rows = None
        self.next_rows()

    def next_rows(self):
        """Switch to the next staging buffer (after its previous copy has left the host) and return its rows."""
        self.slot = (self.slot + 1) % len(self.hosts)
        ev = self.events[self.slot]
        if ev is not None:
            ev.synchronize()
        self.rows = (OptimTensor * self.n_tensors).from_address(self.hosts[self.slot].data_ptr())
        return self.rows

    def upload(self) -> None:
        self.dev.copy_(self.hosts[self.slot], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events[self.slot] = ev


def _check(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
        raise TanError(f"{what}: the fused optimizer step needs contiguous fp32 CUDA tensors "
                       f"(got {t.dtype}, {t.device}, contiguous={t.is_contiguous()})")


_ema_plans = {}


@torch.no_grad()
def ema_update(target_params: Iterable[torch.Tensor], online_params: Iterable[torch.Tensor], m: float) -> None:
    """p_t = m p_t + (1 - m) p_o for every parameter pair (TwinTemporalAligner._momentum_update,
    model/tan_model.py:340-344): one launch of tan_ema_update, same two-products-one-sum arithmetic as the reference.
    Bumps the targets' version counters so that the bf16 weight shadows of the target model are re-cast."""
    tgt, src = list(target_params), list(online_params)
    if len(tgt) != len(src) or not tgt:
        raise TanError("ema_update: parameter lists differ in length or are empty")
    for a, b in zip(tgt, src):
        _check(a, "ema target")
        _check(b, "ema source")
        if a.shape != b.shape:
            raise TanError("ema_update: shape mismatch")
    key = tuple(p.data_ptr() for p in tgt) + tuple(p.data_ptr() for p in src)
    plan = _ema_plans.get(key)
    if plan is None:
        _ema_plans.clear()
        plan = _ema_plans[key] = _Plan([p.numel() for p in tgt], tgt[0].device)
        for r, a, b in zip(plan.rows, tgt, src):
            r.param, r.grad, r.exp_avg, r.exp_avg_sq, r.ema, r.numel = b.data_ptr(), None, None, None, a.data_ptr(), a.numel()
            r.decay, r.neg_step_size = 1.0, 0.0
        plan.upload()
    check(lib().tan_ema_update(plan.dev.data_ptr(), plan.chunk_tensor.data_ptr(), plan.chunk_start.data_ptr(),
                               plan.n_chunks, CHUNK, float(m), _stream()), "tan_ema_update")
    _bump_versions(tgt)                # written by a kernel outside torch's view: advance the version counters


def _bump_versions(tensors: List[torch.Tensor]) -> None:
    """Tensors updated by our kernels behind torch's back: bump `_version` (what _Bf16Cache watches) with one
    multi-tensor no-op write."""
    torch._foreach_add_(tensors, 0.0)


class FusedAdamW(torch.optim.Optimizer):
    """AdamW with the reference's per-parameter gradient clipping and (optionally) the EMA target update folded into
    the same two kernel launches.  `ema=(target_params, online_params, m)` pairs every online parameter with its
    target copy by position (TwinTemporalAligner: `model.target.parameters()`, `model.online.parameters()`)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, clip_grad: float = 0.0, ema=None):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        b = {tuple(g["betas"]) for g in self.param_groups} | {g["eps"] for g in self.param_groups}
        if len(b) != 2:
            raise TanError("FusedAdamW: betas / eps must be the same in every parameter group")
        self.clip_grad = float(clip_grad)
        self._ema_of = {}
        self._ema_m = 0.0
        if ema is not None:
            tgt, src, m = ema
            self._ema_of = {id(s): t for t, s in zip(list(tgt), list(src))}
            self._ema_m = float(m)
        self._plan: Optional[_Plan] = None
        self._flat: List[torch.Tensor] = []
        self._steps = 0
        self.last_norms: Optional[torch.Tensor] = None      # device vector [n_params], in param-group order

    def _build(self):
        self._flat = [p for g in self.param_groups for p in g["params"]]
        for p in self._flat:
            _check(p, "parameter")
            st = self.state[p]
            if "exp_avg" not in st:
                st["step"] = torch.tensor(0.0)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        self._plan = _Plan([p.numel() for p in self._flat], self._flat[0].device)

    @torch.no_grad()
    def step(self, closure=None, inv_scale: Optional[torch.Tensor] = None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._plan is None or len(self._flat) != sum(len(g["params"]) for g in self.param_groups):
            self._build()
        plan = self._plan
        plan.next_rows()
        self._steps += 1
        step = self._steps
        i = 0
        beta1, beta2 = self.param_groups[0]["betas"]
        eps = self.param_groups[0]["eps"]
        touched = []
        for g in self.param_groups:
            lr, wd = g["lr"], g["weight_decay"]
            decay = 1.0 - lr * wd
            neg_step = (lr / (1.0 - beta1 ** step)) * -1.0
            for p in g["params"]:
                r = plan.rows[i]
                i += 1
                st = self.state[p]
                st["step"] += 1
                grad = p.grad
                if grad is not None:
                    if grad.dtype != torch.float32 or not grad.is_contiguous():
                        grad = p.grad = grad.float().contiguous()
                    touched.append(p)
                ema = self._ema_of.get(id(p))
                r.param, r.grad = p.data_ptr(), (grad.data_ptr() if grad is not None else None)
                r.exp_avg, r.exp_avg_sq = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
                r.ema = ema.data_ptr() if (ema is not None and grad is not None) else None
                r.numel, r.decay, r.neg_step_size = p.numel(), decay, neg_step
                if ema is not None and grad is not None:
                    touched.append(ema)
        plan.upload()
        check(lib().tan_optim_adamw_step(plan.dev.data_ptr(), plan.n_tensors, plan.chunk_tensor.data_ptr(),
                                         plan.chunk_start.data_ptr(), plan.first.data_ptr(), plan.n_chunks, CHUNK,
                                         self.clip_grad, float(beta1), float(beta2), float(eps), step, self._ema_m,
                                         None if inv_scale is None else inv_scale.data_ptr(), plan.partial.data_ptr(),
                                         plan.norms.data_ptr(), _stream()), "tan_optim_adamw_step")
        self.last_norms = plan.norms
        if touched:
            _bump_versions(touched)
        return loss
