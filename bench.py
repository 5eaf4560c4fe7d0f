#!/usr/bin/env python
"""bench.py -- clips/sec of the TAN hot path (forward + MIL-NCE loss) on B200.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...       # the reference's CPU path (oracle port) on the host cores

Workload = BASELINE.json configs[2], the configuration the metric is quoted on: E6D6, T=256 frames,
d=512, GLOBAL batch 256 clips (N=32 sentences per clip), contrastive negatives spanning the global
batch.  It fits one B200 (the fused similarity+NCE kernel never materialises the 12.9 GB logits), so
N=1 runs all 256 clips on one GPU and N GPUs shard the same batch, 256/N clips each (strong scaling:
total work is fixed; at N=8 this is exactly configs[2], 32 clips per GPU).  Synthetic features (seed
888), reference-init weights.  One "step" = one
`TemporalAligner.forward` + `get_loss` (`--model init`) over the batch; the reference's backward /
optimizer are not part of this metric (forward+loss is what §8 row (a) covers this round).

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

E_LAYERS, D_LAYERS, T_FRAMES, WIDTH, B_GLOBAL, N_TEXT, VIDEO_DIM = 6, 6, 256, 512, 256, 32, 1024
CPU_SAMPLE_CLIPS = 16
METRIC = "clips/sec (forward + MIL-NCE loss; E6D6, T=256, d=512, global batch 256, global negatives)"


def workload_config(n_gpus):
    return {"workload": f"BASELINE configs[2]: E6D6 T={T_FRAMES} d={WIDTH} global batch {B_GLOBAL} "
                        f"(N={N_TEXT} sentences/clip, D_in={VIDEO_DIM}), {B_GLOBAL // n_gpus} clips per GPU",
            "step": "TemporalAligner.forward + get_loss(model=init), fused similarity+NCE (no logits in HBM)",
            "global_batch": B_GLOBAL, "seq_len": T_FRAMES,
            "parallelism": f"dp{n_gpus} (video batch sharded; text features all-gathered)" if n_gpus > 1 else "single GPU",
            "l2": "flushed between timed steps (256 MiB write), each step timed by its own CUDA event pair"}


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's path on the host cores
# --------------------------------------------------------------------------------------------------
def cpu_reference_clips_per_sec(steps, warmup, clips=CPU_SAMPLE_CLIPS):
    """Times oracle/tan_oracle.py (fp32 torch-CPU restatement of model/tan_model.py forward +
    train/loss.py get_loss; the reference itself is Python-on-torch and /root/reference does not
    exist on the GPU box) on a bounded sample of the workload: `clips` clips of the same shape."""
    import torch

    from oracle import tan_oracle as O
    from temporalalignnet_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synth.make_state_dict(E_LAYERS, D_LAYERS, perturb=False)
    batch = synth.make_batch(clips, T_FRAMES, N_TEXT)
    orc = O.TanOracle(sd, E_LAYERS, D_LAYERS)
    video, text = torch.from_numpy(batch["video"]), torch.from_numpy(batch["text"])

    def step():
        with torch.no_grad():
            out = orc.forward(video, text, batch["video_padding_mask"], batch["text_padding_mask"])
            return float(O.get_loss_init(out["logits_dual"], out["logits_joint"], batch["start"], batch["end"],
                                         batch["text_padding_mask"])["loss"])
    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)

    # the same sample as a full training step (fwd + loss + bwd through torch autograd, no optimizer): the CPU
    # counterpart of the GPU arm's `train_step` leg (SURVEY.md 8(d) asks for both); 1 warm-up + best of 2
    def train_step_cpu():
        leaves = {k: torch.from_numpy(v).clone().requires_grad_(True) for k, v in sd.items()}
        orc.sd = leaves
        out = orc.forward(video, text, batch["video_padding_mask"], batch["text_padding_mask"])
        O.get_loss_init(out["logits_dual"], out["logits_joint"], batch["start"], batch["end"],
                        batch["text_padding_mask"])["loss"].backward()
    train_times = []
    for i in range(3):
        t0 = time.perf_counter()
        train_step_cpu()
        if i > 0:
            train_times.append(time.perf_counter() - t0)
    return {"value": clips / (sum(times) / len(times)), "best": clips / min(times), "cores": cores,
            "train_value": clips / min(train_times),
            "sample": f"{clips} clips (E6D6, T={T_FRAMES}, N={N_TEXT}; negatives span only the {clips}-clip sample, "
                      f"i.e. 1/{B_GLOBAL // clips} of the workload's similarity work per clip -- the reference's fp32 "
                      f"logits of the full batch, 2 x 12.9 GB, do not fit the time budget), "
                      f"fp32 torch-CPU oracle port, {warmup} warm-up + mean of {steps} steps",
            "ms_per_step": 1e3 * sum(times) / len(times)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_clips_per_sec(args.steps, max(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": round(r["value"], 3), "unit": "clips/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(r["ms_per_step"], 2),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": round(r["value"], 3), "unit": "clips/s", "cores": r["cores"], "kind": "port",
                             "sample": r["sample"], "train_step_value": round(r["train_value"], 3)},
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

    if B_GLOBAL % world != 0:
        raise SystemExit(f"--gpus must divide the global batch {B_GLOBAL}")
    B_PER_GPU = B_GLOBAL // world
    steps, warmup = args.steps, max(args.warmup, 3)
    runner = TanStepRunner(E_LAYERS, D_LAYERS, B_PER_GPU, T_FRAMES, N_TEXT, WIDTH, VIDEO_DIM, device=f"cuda:{local}",
                           rank=rank, world_size=world, use_graph=not args.no_graph)
    loss0 = runner.warmup(warmup)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region 1: device-resident steps ------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    n0 = ops.launches()
    for i in range(steps):
        flush.zero_()
        ev[i][0].record()
        loss_t = runner.step_resident()
        ev[i][1].record()
    barrier()
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
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
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = B_PER_GPU * world * steps / e2e_s

    # ---- roofline pass: one CUDA graph per kernel class holding exactly that class's launches of a step
    # (runner.class_graph), replayed `steps` times with the L2 flushed in between; duration = CUDA events
    # around each replay, so achieved = algorithmic flops of the class / its measured device time.
    agg = {}
    if True:
        for name in ("linear", "attention", "layernorm", "sim_nce_fwd", "glue"):
            g, work, n_launch = runner.class_graph(name)
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
    roofline = None
    if agg:
        lin = agg["linear"]
        lin_tf = lin[0] / (lin[1] * 1e-3) / 1e12
        kernel_ms = {k: round(v[1] / steps, 4) for k, v in agg.items()}
        roofline = {"kernel": "umma_gemm2_kernel<LinearEpi2> + gemm_res_ln_kernel (tan_linear_bf16 / tan_linear_res_ln_bf16: "
                              "pre / QKV / MLP projections, out-projection fused with residual + ln_2; the kernel class "
                              "with the largest share of the step)",
                    "bound": "tensor", "achieved": round(lin_tf, 1), "peak": tf_peak, "unit": "TFLOP/s",
                    "frac": round(lin_tf / tf_peak, 4),
                    "traffic": (traffic.get("linear") or {}).get("bytes_per_launch") if world == 1 else None,
                    "traffic_source": (traffic.get("linear") or {}).get("source"), "peak_source": peak_src,
                    "launches_per_step": lin[2] // steps, "avg_launch_us": round(1e3 * lin[1] / max(lin[2], 1), 2),
                    "share_of_step": round(lin[1] / steps / ms_per_step, 3),
                    "how": "CUDA graph of the step's tan_linear_bf16 launches alone, CUDA events per replay, "
                           "L2 flushed between replays"}
        s_ = agg["sim_nce_fwd"]
        extra["roofline_sim"] = {"kernel": "sim_fused_kernel + sim_reduce_partials_kernel (tan_sim_nce_fwd, fused mode: "
                                           "the logits never reach HBM)", "bound": "tensor",
                                 "achieved": round(s_[0] / (s_[1] * 1e-3) / 1e12, 1), "peak": tf_peak,
                                 "unit": "TFLOP/s", "frac": round(s_[0] / (s_[1] * 1e-3) / 1e12 / tf_peak, 4),
                                 "traffic": (traffic.get("sim_nce_fwd") or {}).get("bytes_per_launch") if world == 1 else None,
                                 "columns": "local columns only (collectives are excluded from the per-class graphs)"
                                 if world > 1 else "global"}
        a_ = agg["attention"]
        extra["roofline_attention"] = {"kernel": "attention_kernel (tan_attention_bf16, tcgen05)", "bound": "tensor",
                                       "achieved": round(a_[0] / (a_[1] * 1e-3) / 1e12, 1), "peak": tf_peak,
                                       "unit": "TFLOP/s", "frac": round(a_[0] / (a_[1] * 1e-3) / 1e12 / tf_peak, 4),
                                       "avg_launch_us": round(1e3 * a_[1] / max(a_[2], 1), 2)}
        extra["kernel_ms_per_step"] = kernel_ms
        # north_star's "attention / encoder path": both transformer stacks (projections, attention core, LayerNorms)
        fl_ = runner.flops_per_clip()
        enc_ms = (agg["linear"][1] + agg["attention"][1] + agg["layernorm"][1]) / steps
        extra["encoder_path"] = {"flops_per_clip": fl_["pre"] + fl_["enc"] + fl_["joint"], "ms_per_step": round(enc_ms, 4),
                                 "achieved": round((fl_["pre"] + fl_["enc"] + fl_["joint"]) * B_PER_GPU / (enc_ms * 1e-3) / 1e12, 1),
                                 "unit": "TFLOP/s", "peak": tf_peak,
                                 "frac": round((fl_["pre"] + fl_["enc"] + fl_["joint"]) * B_PER_GPU / (enc_ms * 1e-3) / 1e12 / tf_peak, 4)}
    fl = runner.flops_per_clip()
    extra["whole_step_tensor_frac"] = round(fl["total"] * B_PER_GPU / (ms_per_step * 1e-3) / 1e12 / tf_peak, 4)

    # ---- HBM-bound contrastive pass on materialised logits (API-preserving mode), rank 0 only --------
    if rank == 0 and not args.skip_hbm:
        extra["roofline_nce_hbm"] = hbm_nce_roofline(runner, peaks, flush)

    # ---- training step (forward with tape + get_loss + backward, no optimizer) ----------
    # reported beside the headline metric, never as it: eager launches, 3 warm-up + `steps` timed steps
    if world == 1 and not args.skip_train:
        extra["train_step"] = train_step_leg(runner, min(steps, 5), flush, fl, tf_peak)

    if rank == 0 and world == 1 and args.eager_gpu:
        extra["eager_gpu_baseline"] = eager_gpu_baseline()

    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        r = cpu_reference_clips_per_sec(3, 1)
        cpu = {"value": round(r["best"], 3), "unit": "clips/s", "cores": r["cores"], "kind": "port",
               "sample": r["sample"].replace("mean of 3", "best of 3"),
               "train_step_value": round(r["train_value"], 3)}      # fwd + loss + bwd (torch autograd) on the same sample

    if world > 1:
        dist.barrier()
    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 1), "unit": "clips/s", "n_gpus": world, "steps": steps,
                "warmup": warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(world),
                "loss": round(loss_val, 6), "loss_api": round(loss_api, 6), "cuda_graph": runner._graph is not None,
                "e2e": {"value": round(e2e_value, 1), "unit": "clips/s", "h2d_bytes_per_step": runner.h2d_bytes,
                        "d2h_bytes_per_step": runner.d2h_bytes},
                "gpu_launches": launches_per_step * steps, "gpu_launches_per_step": launches_per_step,
                "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def train_step_leg(runner, steps, flush, fl, tf_peak):
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
        gn = sum(float(p.grad.float().norm() ** 2) for p in runner.model.parameters() if p.grad is not None) ** 0.5
        for p in runner.model.parameters():
            p.grad = None
        tf = 3.0 * fl["total"] * runner.B / (ms * 1e-3) / 1e12
        return {"value": round(runner.B / (ms * 1e-3), 1), "unit": "clips/s", "ms_per_step": round(ms, 3),
                "steps": steps, "what": "forward (activations kept) + get_loss + loss.backward(), no optimizer; eager "
                "launches; first correct backward path (DESIGN.md section 7)", "loss": round(float(loss), 6),
                "grad_norm": round(gn, 6), "achieved": round(tf, 1), "peak": tf_peak, "unit_flops": "TFLOP/s",
                "frac": round(tf / tf_peak, 4), "max_memory_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}
    except Exception as e:                                   # never lose the headline line to the extra leg
        return {"error": f"{type(e).__name__}: {e}"[:300]}


def eager_gpu_baseline(clips=32, steps=5):
    """OPTIONAL leg (--eager-gpu): the torch port of the reference's path (oracle/tan_oracle.py: torch matmul /
    softmax / layer_norm / logsumexp, i.e. cuBLAS + torch eager kernels) on the SAME B200, fp32 and under bf16
    autocast, on a `clips`-clip sample of the workload (its logits are materialised, so the full batch does not
    fit): SURVEY.md 8(d)'s "existing GPU kernel" bar.  A baseline measurement like cpu_baseline, never the product."""
    import torch

    from oracle import tan_oracle as O
    from temporalalignnet_b200 import synth
    sd = {k: torch.from_numpy(v).cuda() for k, v in synth.make_state_dict(E_LAYERS, D_LAYERS, perturb=False).items()}
    batch = synth.make_batch(clips, T_FRAMES, N_TEXT)
    orc = O.TanOracle(sd, E_LAYERS, D_LAYERS)
    orc.sd = sd
    video, text = torch.from_numpy(batch["video"]).cuda(), torch.from_numpy(batch["text"]).cuda()
    vpm = torch.from_numpy(batch["video_padding_mask"]).cuda()
    tpm = torch.from_numpy(batch["text_padding_mask"]).cuda()
    out = {}
    for name, ctx in (("fp32", torch.autocast("cuda", enabled=False)), ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
        try:
            def step():
                with torch.no_grad(), ctx:
                    o = orc.forward(video, text, vpm, tpm)
                    return O.get_loss_init(o["logits_dual"].float().cpu(), o["logits_joint"].float().cpu(), batch["start"],
                                           batch["end"], batch["text_padding_mask"])["loss"]
            step()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(steps):
                step()
            torch.cuda.synchronize()
            out[name] = round(clips * steps / (time.perf_counter() - t0), 1)
        except Exception as e:
            out[name] = f"{type(e).__name__}: {e}"[:200]
    out["unit"] = "clips/s"
    out["sample"] = f"{clips} clips; encoders on the GPU (torch eager), the loss of the port on the host (its mask logic is CPU code)"
    return out


def hbm_nce_roofline(runner, peaks, flush):
    """tan_nce_from_logits on the materialised bf16 logits of this workload: algorithmic bytes =
    the logits read once; measured with CUDA events, L2 flushed."""
    import torch

    from temporalalignnet_b200 import loss as loss_mod
    from temporalalignnet_b200 import ops
    out = runner.model(runner.d_video, runner.d_text, video_padding_mask=runner.d_vpm, lang_padding_mask=runner.d_tpm)
    lg = out["logits_joint"]
    if runner.shard:
        return None
    dense = lg.materialize()
    B, S, T, B2, N = dense.shape
    g = ops.sim_geom(B, S, T, B2 * N, N, 1, 0)
    rs = torch.empty(2, B * S * T, dtype=torch.float32, device=dense.device)
    cs = torch.empty(2, S, B2 * N, dtype=torch.float32, device=dense.device)
    ws = torch.empty(ops.sim_workspace_bytes(g), dtype=torch.uint8, device=dense.device)
    nce = runner.nce
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--skip-hbm", action="store_true", help="skip the materialised-logits HBM roofline leg")
    ap.add_argument("--skip-train", action="store_true", help="skip the training-step (fwd+loss+bwd) leg")
    ap.add_argument("--eager-gpu", action="store_true", help="also time the torch-eager port of the reference on the GPU (sample)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
