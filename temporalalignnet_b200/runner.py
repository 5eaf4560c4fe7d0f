"""Step runner for the TAN hot path: one step = TemporalAligner.forward + get_loss over one batch of clips.
Used by bench.py and the multi-GPU check scripts.

  * `run_api_steps(n)`       the public API exactly as train/main.py:81-105 calls it, starting from pinned HOST
                             tensors / python lists: H2D copies, forward, get_loss, `.item()`.
  * `step_resident()`        same kernels on inputs already resident in HBM, no host sync; with `use_graph=True` the
                             whole step (about 100 kernel launches) is replayed from one CUDA graph.
  * `step_train()`           forward with tape + get_loss + loss.backward() (no optimizer).

Multi-GPU (one process per GPU): ONE global batch is generated from the seed and every rank takes its slice
[rank * B_loc, (rank + 1) * B_loc), so the loss of the global batch is the same number at every world size (a free
cross-N parity check); the contrastive matrix spans the global batch (loss.py all-gathers text features / targets
and all-reduces column sums).
"""
from __future__ import annotations

import os
import types
from typing import Optional

import torch

from . import loss as loss_mod
from . import ops, synth
from .tan_model import TemporalAligner


def default_loss_args(**kw):
    d = dict(model="init", sim="cos", learn_agreement=0, temporal_agreement_type="keep", loss_threshold=0.0,
             use_alignability_head=0, optim_policy="default")
    d.update(kw)
    return types.SimpleNamespace(**d)


def slice_batch(batch: dict, lo: int, hi: int) -> dict:
    """Clips [lo, hi) of a synth.make_batch dict."""
    out = {}
    for k, v in batch.items():
        out[k] = v[lo:hi]
    return out


class TanStepRunner:
    def __init__(self, num_encoder_layers=6, num_decoder_layers=6, B_loc=32, T=256, N=None, width=512,
                 video_dim=1024, seed=888, device="cuda", rank=0, world_size=1, use_graph=True, loss_flags=None):
        self.E, self.D, self.B, self.T = num_encoder_layers, num_decoder_layers, B_loc, T
        self.N = N if N is not None else max(T // 8, 1)
        self.width, self.video_dim = width, video_dim
        self.device = torch.device(device)
        self.rank, self.world = rank, world_size
        self.shard = world_size > 1
        self.flags = dict(loss_flags or {})
        self.args = default_loss_args(**self.flags)
        # loss recipes with flags (config 5) run through get_loss itself: python glue on [B*N] vectors, pinned
        # staging buffers -> eager launches.  The plain recipe is graph-captured: one GPU: the whole step is one CUDA
        # graph; several GPUs: the forward (no collectives) replays from the model's own graph cache and the loss
        # (NCCL all-gather / all-reduce + a dozen launches) is captured together with its collectives in ONE step graph
        # (measured on 2 and 8 GPUs: 2.57 -> 2.46 ms per step at N = 8, profiles/r02y); TAN_GRAPH_NCCL=0 enqueues the
        # loss eagerly behind the forward graph instead, and a failed capture falls back to that by itself
        self.graph_nccl = os.environ.get("TAN_GRAPH_NCCL", "1") == "1"
        self.use_graph = use_graph and not self.flags and (world_size == 1 or self.graph_nccl)
        self.use_model_graph = use_graph
        head = int(self.flags.get("use_alignability_head", 0))
        sd = synth.make_state_dict(self.E, self.D, width=width, d_in=video_dim, seed=seed, perturb=False,
                                   use_alignability_head=bool(head))
        self.model = TemporalAligner(self.E, self.D, random_pos_start=0, width=width, video_dim=video_dim,
                                     use_alignability_head=head)
        self.model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        self.model = self.model.to(self.device)
        if os.environ.get("TAN_TWO_STREAMS") is not None:            # A/B aid
            self.model.two_streams = os.environ["TAN_TWO_STREAMS"] != "0"
        if self.use_model_graph:
            self.model.enable_cuda_graphs(True)
        # ONE global batch; this rank's slice
        B_glob = B_loc * world_size
        full = synth.make_batch(B_glob, T, self.N, d_in=video_dim, seed=seed, tag="global")
        self.batch = slice_batch(full, rank * B_loc, (rank + 1) * B_loc)
        self.global_n_b = [len(s) for s in full["start"]]
        del full
        # pinned host copies (the e2e path starts here) and device-resident copies
        self.h_video = torch.from_numpy(self.batch["video"]).pin_memory()
        self.h_text = torch.from_numpy(self.batch["text"]).pin_memory()
        self.h_vpm = torch.from_numpy(self.batch["video_padding_mask"]).pin_memory()
        self.h_tpm = torch.from_numpy(self.batch["text_padding_mask"]).pin_memory()
        self.d_video = self.h_video.to(self.device)
        self.d_text = self.h_text.to(self.device)
        self.d_vpm = self.h_vpm.to(self.device)
        self.d_tpm = self.h_tpm.to(self.device)
        self.input_data = {"start": self.batch["start"], "end": self.batch["end"], "text": self.batch["text_str"]}
        if self.shard and os.environ.get("TAN_NO_COUNT_HINT") is None:
            self.input_data["n_sentences_global"] = self.global_n_b    # the runner sliced ONE global batch: no exchange
        self.nce = loss_mod.prepare_nce_inputs(self.batch["start"], self.batch["end"], self.d_tpm, T, self.N,
                                               self.device, self.shard, compact=not self.flags and self.N <= 64)
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in (self.h_video, self.h_text, self.h_vpm, self.h_tpm))
        self.d2h_bytes = 4
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._graph_loss = None
        self.launches_per_step = None

    # -- the public-API step, from host memory ------------------------------------------------------
    def _api_step(self, video, text, vpm, tpm):
        out = self.model(video, text, video_padding_mask=vpm, lang_padding_mask=tpm, text_timestamp=None,
                         abs_text_pos=None)
        return loss_mod.get_loss(self.input_data, video, text, vpm, tpm, out, self.args, None, shard_batch=self.shard)

    def run_api_steps(self, n: int) -> float:
        """n public-API steps with the NEXT step's host->device copies issued on a copy stream while the current
        step computes (two device staging sets, events for the hand-over) -- the device-side counterpart of the
        reference's background batch prefetcher (utils/data_utils.py:9-47, which overlaps the host side only).
        Every step still performs its own H2D copy from pinned memory and its own `.item()` read."""
        dev = self.device
        copy_stream = getattr(self, "_copy_stream", None)
        if copy_stream is None:
            copy_stream = self._copy_stream = torch.cuda.Stream(device=dev)
            self._stage = [tuple(torch.empty_like(t, device=dev) for t in (self.h_video, self.h_text, self.h_vpm, self.h_tpm))
                           for _ in range(2)]
        main = torch.cuda.current_stream(dev)
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]

        def issue(i):
            slot = i & 1
            with torch.cuda.stream(copy_stream):
                if i >= 2:
                    copy_stream.wait_event(consumed[slot])            # step i-2 has finished reading this staging set
                for dst, src in zip(self._stage[slot], (self.h_video, self.h_text, self.h_vpm, self.h_tpm)):
                    dst.copy_(src, non_blocking=True)
                ready[slot].record(copy_stream)

        loss = float("nan")
        issue(0)
        for i in range(n):
            if i + 1 < n:
                issue(i + 1)
            slot = i & 1
            main.wait_event(ready[slot])
            video, text, vpm, tpm = self._stage[slot]
            ld = self._api_step(video, text, vpm, tpm)
            consumed[slot].record(main)
            loss = ld["loss"].item()                                  # D2H read of the step's result
        return loss

    # -- device-resident step -----------------------------------------------------------------------
    def _step_kernels(self) -> torch.Tensor:
        if self.flags:
            return self._api_step(self.d_video, self.d_text, self.d_vpm, self.d_tpm)["loss"]
        out = self.model(self.d_video, self.d_text, video_padding_mask=self.d_vpm, lang_padding_mask=self.d_tpm)
        l_dual, l_joint = loss_mod.nce_losses_pair(out["logits_dual"], out["logits_joint"], self.nce, self.shard)
        return (l_dual + l_joint) / 2

    def warmup(self, n=3):
        # kernels per step, counted on one EAGER step (launches replayed from a CUDA graph do not pass through
        # the ops wrappers; capture passes do, but are not steps)
        graphs_on, self.model._graphs_on = self.model._graphs_on, False
        self._step_kernels()                         # first call also casts the weights to their bf16 shadows
        n0 = ops.launches()
        loss = self._step_kernels()
        self.launches_per_step = ops.launches() - n0
        self.model._graphs_on = graphs_on
        for _ in range(max(n - 2, 1)):
            loss = self._step_kernels()
        torch.cuda.synchronize()
        if self.use_graph and self._graph is None:
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._graph_loss = self._step_kernels()
                self._graph = g
                g.replay()
                torch.cuda.synchronize()
            except Exception as e:                       # noqa: BLE001 -- any capture failure: eager loss instead
                if self.world == 1:
                    raise
                import sys
                print(f"[runner] step-graph capture with NCCL failed on rank {self.rank} ({type(e).__name__}: {e}); "
                      "falling back to the eagerly enqueued loss", file=sys.stderr, flush=True)
                self._graph, self._graph_loss, self.use_graph = None, None, False
                torch.cuda.synchronize()
        return float(loss)

    def step_resident(self) -> torch.Tensor:
        if self._graph is not None:
            self._graph.replay()
            return self._graph_loss
        return self._step_kernels()

    def close(self) -> None:
        """Drop captured graphs (before the process group is destroyed: a graph holding NCCL kernels must not
        outlive its communicator)."""
        self._graph = None
        self._graph_loss = None
        self.model.enable_cuda_graphs(False)
        torch.cuda.synchronize()

    # -- training step: forward with tape + get_loss + backward (no optimizer) -------------------------
    def step_train(self) -> torch.Tensor:
        """fwd + loss + bwd on the resident inputs through the public API (model(...), get_loss, loss.backward()):
        SURVEY.md 8(d) "full train step (fwd+loss+bwd, no optimizer)".  Gradients land in `.grad`."""
        m = self.model
        m.train()
        m.enable_autograd(True)
        for p in m.parameters():
            p.grad = None
        out = m(self.d_video, self.d_text, video_padding_mask=self.d_vpm, lang_padding_mask=self.d_tpm)
        ld = loss_mod.get_loss(self.input_data, self.d_video, self.d_text, self.d_vpm, self.d_tpm, out, self.args,
                               None, shard_batch=self.shard)
        ld["loss"].backward()
        m.enable_autograd(False)                     # the forward-only steps of this runner stay on the inference path
        return ld["loss"].detach()

    # -- work accounting -----------------------------------------------------------
    def flops_per_clip(self, B_glob: Optional[int] = None) -> dict:
        d, T, N, E, D = self.width, self.T, self.N, self.E, self.D
        Bg = B_glob if B_glob is not None else self.B * self.world
        f_layer = lambda L: 24 * L * d * d + 4 * L * L * d
        pre = 2 * T * self.video_dim * d + 2 * N * 512 * d
        enc = E * f_layer(T)
        joint = D * f_layer(T + N)
        sim = (E + D) * 2 * T * (Bg * N) * d
        return {"pre": pre, "enc": enc, "joint": joint, "sim": sim, "total": pre + enc + joint + sim,
                "gemm": pre + (E * T + D * (T + N)) * 24 * d * d}
