"""Worker of tests/test_ddp_gpu.py: `steps` TrainStage steps on this rank's shard of a seeded global batch.

Launched by torchrun (one process per rank).  --one-device puts every rank on cuda:0 and uses gloo for the exchange
(NCCL refuses two ranks on one device), so the multi-rank logic can be checked on a single-GPU box; with >= world GPUs
the ranks take their own device and the all-reduce is NCCL over NVLink, exactly as bench.py runs it.
Rank r writes {flat parameters, per-step losses} to <out>.rank<r>.pt."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from rgb_no_more_b200 import train_step as TS  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--arch", default="vits")
    ap.add_argument("--global-batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--one-device", action="store_true")
    ap.add_argument("--mixup", type=float, default=0.0)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = 0 if args.one_device else int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if args.one_device:
            dist.init_process_group("gloo")
        else:
            dist.init_process_group("nccl", device_id=dev)
    B = args.global_batch // world
    st = TS.TrainStage(dev, arch=args.arch, batch=B, world=world, warmup_steps=10, total_steps=1000, mixup_alpha=0.0,
                       use_graph=True, rank=rank)
    g = torch.Generator().manual_seed(1234)
    losses = []
    for s in range(args.steps):
        xg = (torch.randn((args.global_batch, 196, 384), generator=g) * 0.4).to(torch.bfloat16)
        yg = torch.randint(0, 1000, (args.global_batch,), generator=g)
        # contiguous shards: with mixup the roll-by-one partner must stay inside the rank's shard, so mixup stays off here
        x, y = xg[rank * B:(rank + 1) * B].to(dev), yg[rank * B:(rank + 1) * B].to(dev)
        if args.mixup > 0:
            st.lam_override = args.mixup
        losses.append(float(st.step(x, y)))
    torch.cuda.synchronize()
    torch.save({"flat": st.eng.flat.detach().cpu(), "losses": losses}, f"{args.out}.rank{rank}.pt")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
