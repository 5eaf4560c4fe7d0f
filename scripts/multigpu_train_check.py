"""Multi-GPU check of the training step (run under torchrun, one process per GPU): the parameter gradients of a
GLOBAL batch sharded over W ranks (text features all-gathered, column sums all-reduced, text-feature gradients and
weight gradients all-reduced over NCCL inside loss.backward()) must equal the gradients of the same global batch
computed on ONE GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29512 scripts/multigpu_train_check.py
"""
import os
import sys
import types

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from temporalalignnet_b200 import TemporalAligner, get_loss, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = f"cuda:{local}"
    E, D, T, N, B_loc = 2, 3, 64, 8, 6
    Bg = B_loc * world
    sd = synth.make_state_dict(E, D)
    batch = synth.make_batch(Bg, T, N, pad_video_every=3)
    m = TemporalAligner(E, D, random_pos_start=0, use_text_pos_enc=1)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    m = m.to(dev)
    m.train()
    m.enable_autograd(True)
    args = types.SimpleNamespace(model="init", sim="cos", learn_agreement=0, loss_threshold=0.0, use_alignability_head=0,
                                 optim_policy="default", temporal_agreement_type="keep")

    def run(lo, hi, shard):
        for p in m.parameters():
            p.grad = None
        video = torch.from_numpy(batch["video"][lo:hi]).to(dev)
        text = torch.from_numpy(batch["text"][lo:hi]).to(dev)
        vpm = torch.from_numpy(batch["video_padding_mask"][lo:hi]).to(dev)
        tpm = torch.from_numpy(batch["text_padding_mask"][lo:hi]).to(dev)
        out = m(video, text, video_padding_mask=vpm, lang_padding_mask=tpm)
        ld = get_loss({"start": batch["start"][lo:hi], "end": batch["end"][lo:hi], "text": batch["text_str"][lo:hi]},
                      video, text, vpm.float(), tpm.float(), out, args, None, shard_batch=shard)
        ld["loss"].backward()
        torch.cuda.synchronize()
        return float(ld["loss"]), {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}

    loss_s, g_s = run(rank * B_loc, (rank + 1) * B_loc, True)
    loss_1, g_1 = run(0, Bg, False)                       # every rank recomputes the global batch alone
    assert set(g_s) == set(g_1), (set(g_s) ^ set(g_1))
    worst, worst_name = 0.0, ""
    for n in g_1:
        rel = float((g_s[n].double() - g_1[n].double()).norm() / g_1[n].double().norm().clamp_min(1e-30))
        if rel > worst:
            worst, worst_name = rel, n
    t = torch.tensor([worst, abs(loss_s - loss_1) / abs(loss_1)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"world={world} train-step: loss sharded {loss_s:.6f} single {loss_1:.6f} (rel {float(t[1]):.1e}); "
              f"{len(g_1)} parameter gradients, worst rel-Frobenius difference {float(t[0]):.2e} ({worst_name})")
    assert float(t[1]) < 1e-5 and float(t[0]) < 2e-2, "sharded training step differs from the single-GPU one"
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
