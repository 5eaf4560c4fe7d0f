"""Multi-GPU parity check (run under torchrun, one process per GPU):
the loss of a GLOBAL batch sharded over W ranks (text features / targets all-gathered, column sums
all-reduced over NCCL) must equal the loss of the same global batch computed on ONE GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 scripts/multigpu_check.py
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
    sd = synth.make_state_dict(E, D, use_alignability_head=True)
    m = TemporalAligner(E, D, random_pos_start=0, use_alignability_head=1)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    m = m.to(dev)
    arg_sets = [
        dict(model="init", sim="cos", learn_agreement=0, loss_threshold=0.0, use_alignability_head=0,
             optim_policy="default", temporal_agreement_type="keep"),
        # every loss branch: the quantiles / standardisation / BCE span the global batch
        dict(model="init", sim="cos", learn_agreement=1, loss_threshold=0.5, use_alignability_head=1,
             optim_policy="default", temporal_agreement_type="keep"),
    ]

    def run(lo, hi, shard, args):
        video = torch.from_numpy(batch["video"][lo:hi]).to(dev)
        text = torch.from_numpy(batch["text"][lo:hi]).to(dev)
        vpm = torch.from_numpy(batch["video_padding_mask"][lo:hi]).to(dev)
        tpm = torch.from_numpy(batch["text_padding_mask"][lo:hi]).to(dev)
        out = m(video, text, video_padding_mask=vpm, lang_padding_mask=tpm)
        ld = get_loss({"start": batch["start"][lo:hi], "end": batch["end"][lo:hi], "text": batch["text_str"][lo:hi]},
                      video, text, vpm.float(), tpm.float(), out, args, None, shard_batch=shard)
        return {k: float(v) for k, v in ld.items()}

    for kw in arg_sets:
        args = types.SimpleNamespace(**kw)
        sharded = run(rank * B_loc, (rank + 1) * B_loc, True, args)
        single = run(0, Bg, False, args)            # every rank recomputes the global batch alone
        assert set(sharded) == set(single)
        err = max(abs(sharded[k] - single[k]) / max(abs(single[k]), 1e-6) for k in single)
        t = torch.tensor([err], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"world={world} learn={kw['learn_agreement']} sharded={sharded} single={single} "
                  f"max rel err over ranks {float(t):.2e}")
        assert float(t) < 1e-5, "sharded loss differs from the single-GPU loss of the same global batch"
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
