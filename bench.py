#!/usr/bin/env python
"""bench.py -- clips/sec of the TAN hot path (forward + MIL-NCE / alignment loss) on B200.

    python bench.py --gpus 1 --steps 20 --warmup 5 [--config {2,3,4,5}] [--scaling {strong,weak}]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...       # the reference's CPU path (oracle port) on the host cores

Default workload = BASELINE.json configs[2] ("config 3" in SURVEY.md 8(d)), the configuration the metric is quoted
on: E6D6, T=256 frames, d=512, GLOBAL batch 256 clips (N=32 sentences per clip), contrastive negatives spanning the
global batch.  It fits one B200 (the fused similarity+NCE kernel never materialises the 12.9 GB logits), so N=1 runs
all 256 clips on one GPU and N GPUs shard the same batch, 256/N clips each (strong scaling: total work is fixed; at
N=8 this is exactly configs[2], 32 clips per GPU).  `--scaling weak` keeps 32 clips per GPU instead (global batch
32 N).  `--config 2 / 4 / 5` select the other BASELINE shapes (SURVEY numbering; one JSON per config is committed
under profiles/).  Synthetic features (seed 888; ONE global batch sliced by rank, so `loss` is the same number at
every N), reference-init weights.  One "step" = one `TemporalAligner.forward` + `get_loss` over the batch.

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (CUDA events, max over ranks);
`e2e` = the same through the public API from pinned host memory (H2D + D2H inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# SURVEY.md 8(d) numbering (config k = BASELINE.json configs[k-1]); B = global batch
CONFIGS = {
    2: dict(E=6, D=6, T=64, width=512, B=64, N=8, video_dim=1024, flags={}, gpus=1,
            name="BASELINE configs[1]: E6D6 T=64 d=512 batch 64 (paper HTM-370K default shape)"),
    3: dict(E=6, D=6, T=256, width=512, B=256, N=32, video_dim=1024, flags={}, gpus=8,
            name="BASELINE configs[2]: E6D6 T=256 d=512 global batch 256, global NCE"),
    4: dict(E=6, D=6, T=1024, width=768, B=32, N=128, video_dim=768, flags={}, gpus=1,
            name="BASELINE configs[3]: E6D6 T=1024 d=768 (12 heads) batch 32 (long-video attention stress)"),
    5: dict(E=12, D=12, T=512, width=512, B=128, N=64, video_dim=1024, gpus=4,
            flags=dict(learn_agreement=1, use_alignability_head=1, loss_threshold=0.5, temporal_agreement_type="keep"),
            name="BASELINE configs[4]: E12D12 T=512 d=512 batch 128, agreement self-labelling + threshold + "
                 "alignability-head BCE (train/loss.py:88-357)"),
}
CPU_SAMPLE_CLIPS = 16
# module-level aliases of the default configuration (tests/test_bench_contract.py shrinks them)
E_LAYERS, D_LAYERS, T_FRAMES, WIDTH, B_GLOBAL, N_TEXT, VIDEO_DIM = 6, 6, 256, 512, 256, 32, 1024


def get_config(k: int) -> dict:
    c = dict(CONFIGS[k])
    if k == 3:           # the default follows the module-level aliases
        c.update(E=E_LAYERS, D=D_LAYERS, T=T_FRAMES, width=WIDTH, B=B_GLOBAL, N=N_TEXT, video_dim=VIDEO_DIM)
    return c


def metric_name(c: dict) -> str:
    loss = "MIL-NCE loss" if not c["flags"] else "alignment loss (self-labelling + threshold + alignability BCE)"
    return (f"clips/sec (forward + {loss}; E{c['E']}D{c['D']}, T={c['T']}, d={c['width']}, global batch {c['B']}, "
            f"global negatives)")


def workload_config(c: dict, n_gpus: int, scaling: str, b_per_gpu: int) -> dict:
    return {"workload": f"{c['name']} (N={c['N']} sentences/clip, D_in={c['video_dim']}), {b_per_gpu} clips per GPU, "
                        f"global batch {b_per_gpu * n_gpus}",
            "step": "TemporalAligner.forward + get_loss" + ("(model=init), fused similarity+NCE (no logits in HBM)"
                                                             if not c["flags"] else f"({c['flags']})"),
            "global_batch": b_per_gpu * n_gpus, "seq_len": c["T"], "scaling_mode": scaling,
            "parallelism": f"dp{n_gpus} (video batch sharded; text features all-gathered)" if n_gpus > 1 else "single GPU",
            "l2": "flushed between timed steps (256 MiB write), each step timed by its own CUDA event pair"}


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's path on the host cores
# --------------------------------------------------------------------------------------------------
def cpu_reference_clips_per_sec(c: dict, steps, warmup, clips=CPU_SAMPLE_CLIPS, train=True):
    """Times oracle/tan_oracle.py (fp32 torch-CPU restatement of model/tan_model.py forward +
    train/loss.py get_loss; the reference itself is Python-on-torch and /root/reference does not
    exist on the GPU box) on a bounded sample of the workload: `clips` clips of the same shape."""
    import torch

    from oracle import tan_oracle as O
    from temporalalignnet_b200 import synth
    from temporalalignnet_b200.runner import default_loss_args
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    E, D, T, N = c["E"], c["D"], c["T"], c["N"]
    head = int(c["flags"].get("use_alignability_head", 0))
    if T * N * clips * clips * (E + D) > 3e9:           # keep the sample's fp32 logits + loss temporaries in RAM
        clips = max(2, clips // 4)
    sd = synth.make_state_dict(E, D, width=c["width"], d_in=c["video_dim"], perturb=False, use_alignability_head=bool(head))
    batch = synth.make_batch(clips, T, N, d_in=c["video_dim"], force_full=bool(c["flags"]))
    orc = O.TanOracle(sd, E, D, use_alignability_head=head)
    video, text = torch.from_numpy(batch["video"]), torch.from_numpy(batch["text"])
    args = default_loss_args(**c["flags"])

    def loss_of(out):
        if c["flags"]:
            return O.get_loss_full(out, batch["start"], batch["end"], torch.from_numpy(batch["video_padding_mask"]),
                                   torch.from_numpy(batch["text_padding_mask"]), args)["loss"]
        return O.get_loss_init(out["logits_dual"], out["logits_joint"], batch["start"], batch["end"],
                               batch["text_padding_mask"])["loss"]

    def step():
        with torch.no_grad():
            return float(loss_of(orc.forward(video, text, batch["video_padding_mask"], batch["text_padding_mask"])))
    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)

    # the same sample as a full training step (fwd + loss + bwd through torch autograd, no optimizer): the CPU
    # counterpart of the GPU arm's `train_step` leg (SURVEY.md 8(d) asks for both); 1 warm-up + best of 2
    train_value = None
    if train:
        def train_step_cpu():
            leaves = {k: torch.from_numpy(v).clone().requires_grad_(True) for k, v in sd.items()}
            orc.sd = leaves
            loss_of(orc.forward(video, text, batch["video_padding_mask"], batch["text_padding_mask"])).backward()
        train_times = []
        for i in range(3):
            t0 = time.perf_counter()
            train_step_cpu()
            if i > 0:
                train_times.append(time.perf_counter() - t0)
        train_value = clips / min(train_times)
    mean = sum(times) / len(times)
    # per-clip cost of the reference on the FULL batch: encoders scale with clips, the similarity + loss with
    # clips x global columns -- extrapolated from the measured sample by the flop model (stated, not measured)
    f_layer = lambda L: 24 * L * c["width"] ** 2 + 4 * L * L * c["width"]
    enc = E * f_layer(T) + D * f_layer(T + N) + 2 * T * c["video_dim"] * c["width"]
    sim = lambda Bg: (E + D) * 2 * T * (Bg * N) * c["width"]
    scale = (enc + sim(c["B"])) / (enc + sim(clips))
    return {"value": clips / mean, "best": clips / min(times), "cores": cores, "train_value": train_value,
            "full_batch_extrapolated": clips / mean / scale,
            "sample": f"{clips} clips (E{E}D{D}, T={T}, N={N}; negatives span only the {clips}-clip sample, "
                      f"i.e. 1/{max(c['B'] // clips, 1)} of the workload's similarity work per clip -- the reference's fp32 "
                      f"logits of the full batch do not fit the time budget; flop-model extrapolation to the full "
                      f"batch in full_batch_extrapolated), fp32 torch-CPU oracle port, {warmup} warm-up + mean of {steps} steps",
            "ms_per_step": 1e3 * mean}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = get_config(getattr(args, "config", 3))
    r = cpu_reference_clips_per_sec(c, args.steps, max(args.warmup, 1))
    b_per_gpu = c["B"] // args.gpus if getattr(args, "scaling", "strong") == "strong" else c["B"] // c["gpus"]
    line = {"impl": "reference", "metric": metric_name(c), "value": round(r["value"], 3), "unit": "clips/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(r["ms_per_step"], 2),
            "higher_is_better": True, "scaling": getattr(args, "scaling", "strong"), "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(c, args.gpus, getattr(args, "scaling", "strong"), b_per_gpu),
            "cpu_baseline": {"value": round(r["value"], 3), "unit": "clips/s", "cores": r["cores"], "kind": "port",
                             "sample": r["sample"], "train_step_value": round(r["train_value"], 3),
                             "full_batch_extrapolated": round(r["full_batch_extrapolated"], 3)},
            "e2e": {"value": round(r["value"], 3), "unit": "clips/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock + throttle reasons DURING the timed region with NVML from a background thread
    (every ~2 ms: the timed region of a default run is only tens of milliseconds, too short for an
    `nvidia-smi -lms` loop)."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4,
               "hw_power_brake_slowdown": 0x80}

    def __init__(self, gpu_index=0):
        import threading
        self.samples, self.reasons, self.max_mhz, self.power = [], set(), None, []
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for nm, bit in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(nm)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        if self._t is None:
            return out
        self._stop.set()
        self._t.join(timeout=2)
        if self.samples:
            out.update(sm_mhz=statistics.median(self.samples), reasons=sorted(self.reasons), samples=len(self.samples),
                       power_w_max=round(max(self.power), 1) if self.power else None)
        return out


# --------------------------------------------------------------------------------------------------
# measurement instrumentation (lives HERE, not in the product): per-kernel-class CUDA graphs
# --------------------------------------------------------------------------------------------------
class only_class:
    """Inside this context every kernel launch of temporalalignnet_b200.ops except class `name` is dropped, so a CUDA
    graph captured around one step holds exactly that class's launches (in step order, on the step's real buffers).
    Kernels do not branch on data values, so skipping the producers changes no launch.  Implemented by wrapping the
    product's single launch choke point (`ops._launch`) from outside; the product itself has no such switch."""

    def __init__(self, name: str):
        self.name = name
        self.acc = [0.0, 0]

    def __enter__(self):
        from temporalalignnet_b200 import ops
        self._ops, self._orig = ops, ops._launch
        name, acc, orig = self.name, self.acc, ops._launch

        def filtered(cls, work, n, call):
            if cls == name:
                acc[0] += work
                acc[1] += n
                orig(cls, work, n, call)
        ops._launch = filtered
        return self.acc

    def __exit__(self, *a):
        self._ops._launch = self._orig


def class_graph(runner, name: str):
    """(graph, work, launches) holding ONLY the launches of kernel class `name` of one resident step.  Replaying it
    gives that class's device time with CUDA events alone -- per-launch event pairs on eagerly launched kernels also
    time the host's launch preparation whenever the GPU runs dry.  With several GPUs the text features are gathered
    once OUTSIDE the graph, so that the similarity class runs at its real geometry (local rows x GLOBAL columns)."""
    import torch

    from temporalalignnet_b200 import loss as loss_mod
    graphs_on, runner.model._graphs_on = runner.model._graphs_on, False
    full = None
    if runner.shard:
        out = runner.model(runner.d_video, runner.d_text, video_padding_mask=runner.d_vpm, lang_padding_mask=runner.d_tpm)
        full = loss_mod.exchange_text_features(loss_mod.pack_text_features(out["logits_dual"].tfeat,
                                                                           out["logits_joint"].tfeat), loss_mod._dist())
        torch.cuda.synchronize()

    def step():
        if runner.flags:
            return runner._step_kernels()
        out = runner.model(runner.d_video, runner.d_text, video_padding_mask=runner.d_vpm, lang_padding_mask=runner.d_tpm)
        if full is None:
            return loss_mod.nce_losses_pair(out["logits_dual"], out["logits_joint"], runner.nce, False)
        return loss_mod.sim_pair_sums(out["logits_dual"].vfeat, out["logits_joint"].vfeat, full, runner.nce)

    try:
        with only_class(name) as acc:
            step()                                   # eager dry run: allocations
            torch.cuda.synchronize()
            acc[0], acc[1] = 0.0, 0
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                step()
    finally:
        runner.model._graphs_on = graphs_on
    g.replay()
    torch.cuda.synchronize()
    return g, acc[0], acc[1]


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one process per GPU)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from temporalalignnet_b200 import ops
    from temporalalignnet_b200.runner import TanStepRunner

    c = get_config(args.config)
    if args.scaling == "weak":
        B_PER_GPU = c["B"] // c["gpus"]
    else:
        if c["B"] % world != 0:
            raise SystemExit(f"--gpus must divide the global batch {c['B']}")
        B_PER_GPU = c["B"] // world
    steps, warmup = args.steps, max(args.warmup, 3)
    dev = f"cuda:{local}"
    runner = TanStepRunner(c["E"], c["D"], B_PER_GPU, c["T"], c["N"], c["width"], c["video_dim"], device=dev,
                           rank=rank, world_size=world, use_graph=not args.no_graph, loss_flags=c["flags"])
    loss0 = runner.warmup(warmup)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- timed region 1: device-resident steps ------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i in range(steps):
        flush.zero_()
        ev[i][0].record()
        loss_t = runner.step_resident()
        ev[i][1].record()
    barrier()
    total_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in ev))
    clocks = sampler.stop() if sampler else None
    ms_per_step = total_ms / steps
    value = B_PER_GPU * world / (ms_per_step * 1e-3)
    launches_per_step = runner.launches_per_step
    loss_val = float(loss_t)

    # ---- timed region 2: end to end through the public API, from pinned host memory ---------------
    runner.run_api_steps(2)
    barrier()
    t0 = time.perf_counter()
    loss_api = runner.run_api_steps(steps)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = B_PER_GPU * world * steps / e2e_s

    # ---- roofline pass: one CUDA graph per kernel class holding exactly that class's launches of a step
    # (class_graph), replayed `steps` times with the L2 flushed in between; duration = CUDA events around each
    # replay, so achieved = algorithmic flops of the class / its measured device time.
    agg = {}
    if runner.flags:
        # get_loss with flags stages its inputs through pinned host buffers (not capturable): per-class times from
        # CUDA event pairs around every launch of eager steps instead (small classes include launch gaps)
        agg = eager_class_times(runner, steps, flush)
    for name in ("linear", "attention", "layernorm", "sim_nce_fwd", "glue", "cast", "nce_reduce"):
        if runner.flags:
            break
        g, work, n_launch = class_graph(runner, name)
        if n_launch == 0:
            continue
        tot = 0.0
        for _ in range(steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            g.replay()
            b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        agg[name] = [work * steps, tot, n_launch * steps]
        del g
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # dram__bytes_read + dram__bytes_write per launch from the committed `ncu --set full` captures of the same
    # shapes (profiles/ncu_traffic.json, written by scripts/ncu_traffic.py from the .ncu-rep files)
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        pass
    tf_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else \
        "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    extra = {}
    lin = agg["linear"]
    lin_tf = lin[0] / (lin[1] * 1e-3) / 1e12
    kernel_ms = {k: round(v[1] / steps, 4) for k, v in agg.items()}
    default_shape = args.config == 3 and world == 1
    roofline = {"kernel": "umma_gemm2_kernel<LinearEpi2> + gemm_res_ln_kernel (tan_linear_bf16 / tan_linear_res_ln_bf16: "
                          "pre / QKV / MLP projections, out-projection fused with residual + ln_2; the kernel class "
                          "with the largest share of the step)",
                "bound": "tensor", "achieved": round(lin_tf, 1), "peak": tf_peak, "unit": "TFLOP/s",
                "frac": round(lin_tf / tf_peak, 4),
                "traffic": (traffic.get("linear") or {}).get("bytes_per_launch") if default_shape else None,
                "traffic_source": (traffic.get("linear") or {}).get("source") if default_shape else None,
                "peak_source": peak_src,
                "launches_per_step": lin[2] // steps, "avg_launch_us": round(1e3 * lin[1] / max(lin[2], 1), 2),
                "share_of_step": round(lin[1] / steps / ms_per_step, 3),
                "how": "CUDA graph of the step's tan_linear_bf16 launches alone, CUDA events per replay, "
                       "L2 flushed between replays"}
    s_ = agg["sim_nce_fwd"]
    fl = runner.flops_per_clip()
    # useful flops of the similarity: only real sentences are columns of the reference's matrix (it drops the padded
    # ones before the loss, train/loss.py:235).  With ragged columns (default) the kernel computes exactly those and
    # `achieved` counts only them; without, it also computes the padded columns
    nce_ = runner.nce
    col_frac = 1.0 if nce_.compact else float(nce_.col_valid.float().mean().item())
    extra["roofline_sim"] = {"kernel": "sim_fused_kernel + sim_reduce_partials_kernel (tan_sim_nce_fwd, fused mode: "
                                       "the logits never reach HBM)", "bound": "tensor",
                             "achieved": round(s_[0] / (s_[1] * 1e-3) / 1e12, 1), "peak": tf_peak,
                             "unit": "TFLOP/s", "frac": round(s_[0] / (s_[1] * 1e-3) / 1e12 / tf_peak, 4),
                             "ragged_columns": bool(nce_.compact), "columns_computed": int(nce_.C),
                             "columns_padded_layout": int(nce_.C_pad), "valid_column_fraction": round(col_frac, 4),
                             "frac_useful": round(col_frac * s_[0] / (s_[1] * 1e-3) / 1e12 / tf_peak, 4),
                             "traffic": (traffic.get("sim_nce_fwd") or {}).get("bytes_per_launch") if default_shape else None,
                             "columns": "local rows x GLOBAL columns (text features pre-gathered outside the class graph)"}
    a_ = agg["attention"]
    extra["roofline_attention"] = {"kernel": "attention_kernel (tan_attention_bf16, tcgen05)", "bound": "tensor",
                                   "achieved": round(a_[0] / (a_[1] * 1e-3) / 1e12, 1), "peak": tf_peak,
                                   "unit": "TFLOP/s", "frac": round(a_[0] / (a_[1] * 1e-3) / 1e12 / tf_peak, 4),
                                   "avg_launch_us": round(1e3 * a_[1] / max(a_[2], 1), 2)}
    extra["kernel_ms_per_step"] = kernel_ms
    extra["kernel_ms_sum"] = round(sum(kernel_ms.values()), 4)
    # north_star's "attention / encoder path": both transformer stacks (projections, attention core, LayerNorms)
    enc_ms = (agg["linear"][1] + agg["attention"][1] + agg["layernorm"][1]) / steps
    enc_fl = fl["pre"] + fl["enc"] + fl["joint"]
    extra["encoder_path"] = {"flops_per_clip": enc_fl, "ms_per_step": round(enc_ms, 4),
                             "achieved": round(enc_fl * B_PER_GPU / (enc_ms * 1e-3) / 1e12, 1),
                             "unit": "TFLOP/s", "peak": tf_peak,
                             "frac": round(enc_fl * B_PER_GPU / (enc_ms * 1e-3) / 1e12 / tf_peak, 4)}
    extra["whole_step_tensor_frac"] = round(fl["total"] * B_PER_GPU / (ms_per_step * 1e-3) / 1e12 / tf_peak, 4)

    # ---- multi-GPU: the step's collectives timed alone (same message sizes, nothing to overlap with) ----------
    if world > 1 and not c["flags"]:
        extra["comm_ms_per_step"] = comm_only_ms(runner, steps, max_over_ranks)

    # ---- HBM-bound contrastive pass on materialised logits (API-preserving mode), rank 0 only --------
    if rank == 0 and not args.skip_hbm and not c["flags"]:
        extra["roofline_nce_hbm"] = hbm_nce_roofline(runner, peaks, flush)

    # ---- parity carried by the bench line itself: this model's loss on a fixed 8-clip sub-batch, GPU path vs the
    # CPU oracle (fp32) -- rank 0, local computation
    if rank == 0 and not args.skip_cpu:
        extra["loss_parity"] = loss_parity_leg(runner, c)

    # ---- training step (forward with tape + get_loss + backward, no optimizer) ----------
    # reported beside the headline metric, never as it: eager launches, 3 warm-up + `steps` timed steps
    if not args.skip_train and (world == 1 or args.train_multi):
        extra["train_step"] = train_step_leg(runner, min(steps, 5), flush, fl, tf_peak, world, max_over_ranks)

    # ---- the existing GPU implementation: torch eager running the reference's path on the same B200 ----------
    if rank == 0 and world == 1 and not args.skip_eager and not c["flags"]:
        runner.close()
        torch.cuda.empty_cache()
        extra["eager_gpu_baseline"] = eager_gpu_baseline(c)

    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        r = cpu_reference_clips_per_sec(c, 3, 1, train=not c["flags"])
        cpu = {"value": round(r["best"], 3), "unit": "clips/s", "cores": r["cores"], "kind": "port",
               "sample": r["sample"].replace("mean of 3", "best of 3"),
               "full_batch_extrapolated": round(r["full_batch_extrapolated"], 3),
               "train_step_value": round(r["train_value"], 3) if r["train_value"] else None}

    if world > 1:
        dist.barrier()
    if rank == 0:
        line = {"metric": metric_name(c), "value": round(value, 1), "unit": "clips/s", "n_gpus": world, "steps": steps,
                "warmup": warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": args.scaling,
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": workload_config(c, world, args.scaling, B_PER_GPU),
                "loss": round(loss_val, 6), "loss_api": round(loss_api, 6), "cuda_graph": runner._graph is not None or
                (world == 1 and not args.no_graph and not c["flags"]),
                "e2e": {"value": round(e2e_value, 1), "unit": "clips/s", "h2d_bytes_per_step": runner.h2d_bytes,
                        "d2h_bytes_per_step": runner.d2h_bytes},
                "gpu_launches": launches_per_step * steps, "gpu_launches_per_step": launches_per_step,
                "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
        line.update(extra)
        print(json.dumps(line), flush=True)
    runner.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def eager_class_times(runner, steps, flush):
    """{class: [work, ms, launches]} over `steps` eager steps (ops.profile: CUDA event pairs around every launch)."""
    import torch

    from temporalalignnet_b200 import ops
    agg = {}
    graphs_on, runner.model._graphs_on = runner.model._graphs_on, False
    try:
        for _ in range(steps):
            flush.zero_()
            with ops.profile() as prof:
                runner._step_kernels()
                torch.cuda.synchronize()
                for cls, work, a, b in prof:
                    e = agg.setdefault(cls, [0.0, 0.0, 0])
                    e[0] += work
                    e[1] += a.elapsed_time(b)
                    e[2] += 1
    finally:
        runner.model._graphs_on = graphs_on
    return agg


def comm_only_ms(runner, steps, max_over_ranks):
    """The step's exchange (text-feature gather, column-sum all-reduce, row-scalar all-reduce) on buffers of the
    step's sizes, timed alone with CUDA events: an upper bound of what the step pays for communication."""
    import torch

    from temporalalignnet_b200 import loss as loss_mod
    dist = loss_mod._dist()
    W = dist.get_world_size()
    BN, d = runner.B * runner.N, runner.width
    packed = torch.zeros(1 + runner.D, BN, d, dtype=torch.bfloat16, device=runner.device)
    cols = torch.zeros(2 * (runner.E + runner.D) * W * BN, dtype=torch.float32, device=runner.device)
    rows = torch.zeros(4, dtype=torch.float64, device=runner.device)

    def once():
        loss_mod.exchange_text_features(packed, dist)
        dist.all_reduce(cols)
        dist.all_reduce(rows)
    for _ in range(3):
        once()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        once()
    b.record()
    torch.cuda.synchronize()
    return round(max_over_ranks(a.elapsed_time(b) / steps), 4)


def loss_parity_leg(runner, c, clips=8):
    """Forward + get_loss of the bench's own model on its first `clips` clips: GPU path vs the fp32 CPU oracle."""
    import torch

    from oracle import tan_oracle as O
    from temporalalignnet_b200 import loss as loss_mod
    from temporalalignnet_b200 import synth
    try:
        clips = min(clips, runner.B)
        sl = slice(0, clips)
        m = runner.model
        graphs_on, m._graphs_on = m._graphs_on, False
        idata = {k: v[sl] for k, v in runner.input_data.items()}
        out = m(runner.d_video[sl], runner.d_text[sl], video_padding_mask=runner.d_vpm[sl], lang_padding_mask=runner.d_tpm[sl])
        got = float(loss_mod.get_loss(idata, runner.d_video[sl], runner.d_text[sl], runner.d_vpm[sl], runner.d_tpm[sl], out,
                                      runner.args, None, shard_batch=False)["loss"].item())
        m._graphs_on = graphs_on
        head = int(c["flags"].get("use_alignability_head", 0))
        sd = synth.make_state_dict(c["E"], c["D"], width=c["width"], d_in=c["video_dim"], perturb=False,
                                   use_alignability_head=bool(head))
        b = runner.batch
        with torch.no_grad():
            ref_out = O.TanOracle(sd, c["E"], c["D"], use_alignability_head=head).forward(
                torch.from_numpy(b["video"][sl]), torch.from_numpy(b["text"][sl]), b["video_padding_mask"][sl],
                b["text_padding_mask"][sl])
            if c["flags"]:
                ref = float(O.get_loss_full(ref_out, b["start"][sl], b["end"][sl], torch.from_numpy(b["video_padding_mask"][sl]),
                                            torch.from_numpy(b["text_padding_mask"][sl]), runner.args)["loss"])
            else:
                ref = float(O.get_loss_init(ref_out["logits_dual"], ref_out["logits_joint"], b["start"][sl], b["end"][sl],
                                            b["text_padding_mask"][sl])["loss"])
        return {"clips": clips, "loss_gpu": round(got, 6), "loss_ref": round(ref, 6), "rel_err": round(abs(got - ref) / abs(ref), 7),
                "ref": "oracle/tan_oracle.py (fp32 CPU restatement, pinned to the reference by tests/golden)"}
    except Exception as e:
        return {"error": f"{type(e).__name__}: {e}"[:300]}


def train_step_leg(runner, steps, flush, fl, tf_peak, world=1, max_over_ranks=None):
    """fwd + loss + bwd through the public API (`model(...)`, `get_loss`, `loss.backward()`), CUDA events per
    step, L2 flushed between steps.  Algorithmic flops = 3 x the forward's (backward = 2 x forward)."""
    import torch
    try:
        for _ in range(3):
            runner.step_train()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            loss = runner.step_train()
            b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        ms = tot / steps
        if max_over_ranks is not None:
            ms = max_over_ranks(ms)
        gn = sum(float(p.grad.float().norm() ** 2) for p in runner.model.parameters() if p.grad is not None) ** 0.5
        # + the optimizer step (fused clip + AdamW, two launches) as a separate number
        from temporalalignnet_b200.optim import FusedAdamW
        opt = FusedAdamW([p for p in runner.model.parameters() if p.requires_grad], lr=1e-5, weight_decay=1e-5, clip_grad=3.0)
        sd0 = {k: v.detach().clone() for k, v in runner.model.state_dict().items()}
        opt.step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            opt.step()
        b.record()
        torch.cuda.synchronize()
        opt_ms = a.elapsed_time(b) / 5
        runner.model.load_state_dict(sd0)                      # the benchmark's weights stay the seeded ones
        for p in runner.model.parameters():
            p.grad = None
        tf = 3.0 * fl["total"] * runner.B / (ms * 1e-3) / 1e12
        return {"value": round(runner.B * world / (ms * 1e-3), 1), "unit": "clips/s", "ms_per_step": round(ms, 3),
                "steps": steps, "what": "forward (activations kept) + get_loss + loss.backward(), no optimizer; eager "
                "launches (DESIGN.md section 7)", "loss": round(float(loss), 6),
                "grad_norm": round(gn, 6), "achieved": round(tf, 1), "peak": tf_peak, "unit_flops": "TFLOP/s",
                "frac": round(tf / tf_peak, 4), "max_memory_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1),
                "optimizer_step_ms": round(opt_ms, 4),
                "optimizer": "FusedAdamW: per-parameter clip + AdamW in two launches (tan_optim_adamw_step), no host sync"}
    except Exception as e:                                   # never lose the headline line to the extra leg
        return {"error": f"{type(e).__name__}: {e}"[:300]}


def eager_gpu_baseline(c, steps=3):
    """The existing GPU implementation: the reference's path as torch eager runs it on the SAME
    B200 -- oracle/eager_port.py restates the reference's torch calls module for module (nn.MultiheadAttention,
    nn.LayerNorm, einsum, boolean-index loss; /root/reference itself does not exist on the GPU box) -- ENTIRELY on
    the GPU including the loss, under fp16 autocast as train/main.py:81 and under bf16 autocast, forward + loss and
    the full training step (+ backward), at the largest batch that fits (its logits are materialised)."""
    import torch

    from oracle import eager_port as EP
    from temporalalignnet_b200 import synth
    E, D, T, N = c["E"], c["D"], c["T"], c["N"]
    sd = synth.make_state_dict(E, D, width=c["width"], d_in=c["video_dim"], perturb=False)
    model = EP.EagerTAN(E, D, width=c["width"], video_dim=c["video_dim"]).cuda()
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    out = {"unit": "clips/s", "what": "oracle/eager_port.py: torch-eager port of the reference (same torch modules / "
           "calls), forward + get_loss and forward + get_loss + backward, all on the GPU, logits materialised"}

    def run(B, dtype, train):
        batch = synth.make_batch(B, T, N, d_in=c["video_dim"], tag="global")
        video, text = torch.from_numpy(batch["video"]).cuda(), torch.from_numpy(batch["text"]).cuda()
        vpm = torch.from_numpy(batch["video_padding_mask"]).cuda()
        tpm = torch.from_numpy(batch["text_padding_mask"]).cuda()

        def step():
            with torch.autocast("cuda", dtype=dtype):
                o = model(video, text, vpm, tpm)
                loss = EP.eager_get_loss_init(o, batch["start"], batch["end"], tpm, T, N)["loss"]
            if train:
                for p in model.parameters():
                    p.grad = None
                loss.backward()
            return loss
        ctx = torch.enable_grad() if train else torch.no_grad()
        with ctx:
            loss = step()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(steps):
                loss = step()
            b.record()
            torch.cuda.synchronize()
        return B * steps / (a.elapsed_time(b) * 1e-3), float(loss)

    for dname, dtype in (("fp16_autocast", torch.float16), ("bf16_autocast", torch.bfloat16)):
        for train in (False, True):
            key = dname + ("_train_step" if train else "")
            B = c["B"]
            while B >= 8:
                try:
                    v, loss = run(B, dtype, train)
                    out[key] = {"value": round(v, 1), "clips": B, "loss": round(loss, 5)}
                    break
                except torch.cuda.OutOfMemoryError:
                    torch.cuda.empty_cache()
                    B //= 2
                except Exception as e:
                    out[key] = f"{type(e).__name__}: {e}"[:200]
                    break
            torch.cuda.empty_cache()
    return out


def hbm_nce_roofline(runner, peaks, flush):
    """tan_nce_from_logits on the materialised bf16 logits of this workload: algorithmic bytes =
    the logits read once; measured with CUDA events, L2 flushed.  With several GPUs: this rank's
    rows x the GLOBAL columns (text features gathered first), the shape north_star's HBM target is quoted on."""
    import torch

    from temporalalignnet_b200 import loss as loss_mod
    from temporalalignnet_b200 import ops
    from temporalalignnet_b200.tan_model import LazyLogits
    try:
        graphs_on, runner.model._graphs_on = runner.model._graphs_on, False
        out = runner.model(runner.d_video, runner.d_text, video_padding_mask=runner.d_vpm, lang_padding_mask=runner.d_tpm)
        runner.model._graphs_on = graphs_on
        lg = out["logits_joint"]
        if runner.shard:
            return None          # rank-0-only leg: a collective here would deadlock; see comm_ms_per_step / roofline_sim
        dense = lg.materialize()
        B, S, T, B2, N = dense.shape
        g = ops.sim_geom(B, S, T, B2 * N, N, 1, 0)
        rs = torch.empty(2, B * S * T, dtype=torch.float32, device=dense.device)
        cs = torch.empty(2, S, B2 * N, dtype=torch.float32, device=dense.device)
        ws = torch.empty(ops.sim_workspace_bytes(g), dtype=torch.uint8, device=dense.device)
        # materialised logits keep the padded [B, S, T, B, N] layout: targets without the ragged-column compaction
        nce = loss_mod.prepare_nce_inputs(runner.batch["start"], runner.batch["end"], runner.d_tpm, runner.T, runner.N,
                                          runner.device, False, compact=False)
        ts = []
        for i in range(8):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ops.nce_from_logits(dense, g, nce.posbits, nce.col_valid, rs, cs, ws)
            b.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(a.elapsed_time(b))
        ms = sum(ts) / len(ts)
        nbytes = dense.numel() * 2
        hbm = float(peaks.get("hbm_gbs", 6650.0)) if peaks else 6650.0
        gbs = nbytes / (ms * 1e-3) / 1e9
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        except Exception:
            pass
        return {"kernel": "nce_from_logits_kernel<bf16> + partial reduce (tan_nce_from_logits)", "bound": "hbm",
                "achieved": round(gbs, 1), "peak": hbm, "unit": "GB/s", "frac": round(gbs / hbm, 4),
                "traffic": (traffic.get("nce_from_logits") or {}).get("bytes_per_launch"),
                "bytes": nbytes, "ms": round(ms, 4)}
    except Exception as e:
        return {"error": f"{type(e).__name__}: {e}"[:300]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS),
                    help="SURVEY.md 8(d) configuration number (k = BASELINE.json configs[k-1]); default 3 = the headline")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: the configuration's global batch is split over the GPUs; weak: fixed clips per GPU")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline and loss_parity legs")
    ap.add_argument("--skip-hbm", action="store_true", help="skip the materialised-logits HBM roofline leg")
    ap.add_argument("--skip-train", action="store_true", help="skip the training-step (fwd+loss+bwd) leg")
    ap.add_argument("--skip-eager", action="store_true", help="skip the torch-eager-on-GPU baseline leg")
    ap.add_argument("--train-multi", action="store_true", help="run the training-step leg with several GPUs as well")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
