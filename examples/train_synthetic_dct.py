#!/usr/bin/env python
"""The whole B200 DCT path in one loop, from JPEG byte strings to optimiser steps -- the shape of the reference's
`python train.py --domain=dct --train` hot loop (train.py:146-176) with the data side of datasets.py / pipeline_utils.py
replaced by JpegFeeder + FusedDCT:

    JPEG bytes -> Huffman decode (host threads) -> pinned ring -> H2D -> K0 (crop / resize / flip / RandAugment / ToRange /
    sub-block) -> mixup -> ViT forward / backward -> [NCCL all-reduce] -> clip + AdamW + decay

Data: synthetic 512x512 4:2:0 JPEGs whose mean colour encodes the class, so the loss visibly falls within a few dozen
steps.  Single process:   python examples/train_synthetic_dct.py --steps 60
Data parallel:            torchrun --nproc-per-node 8 examples/train_synthetic_dct.py"""
from __future__ import annotations

import argparse
import io
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from rgb_no_more_b200 import ddp, feeder as FD, plan as P, train_step as TS, transforms as TF  # noqa: E402


def make_dataset(n: int, n_classes: int, seed: int):
    from PIL import Image
    rng = np.random.default_rng(seed)
    palette = rng.integers(40, 216, size=(n_classes, 3))
    files, labels = [], []
    for i in range(n):
        k = i % n_classes
        low = np.clip(palette[k] + rng.integers(-30, 31, size=(16, 16, 3)), 0, 255).astype(np.uint8)
        img = Image.fromarray(low).resize((512, 512), Image.BICUBIC)
        b = io.BytesIO()
        img.save(b, "JPEG", quality=75, subsampling=2)       # what the reference's resizer writes (mp_scripts.py:74-81)
        files.append(b.getvalue())
        labels.append(k)
    return files, torch.tensor(labels)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="vitti", choices=list(TS.ARCHS))
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--files", type=int, default=256)
    ap.add_argument("--classes", type=int, default=8)
    ap.add_argument("--lr", type=float, default=1e-3)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    files, labels = make_dataset(args.files, args.classes, seed=1)
    tf = TF.get_transform("imagenet_dct", "train", ops_list=P.AUGLIST_VITS if args.arch != "vitti" else P.AUGLIST_VITTI,
                          num_ops=2, ops_magnitude=3, dtype=torch.bfloat16, device=dev)
    stage = TS.TrainStage(dev, arch=args.arch, batch=args.batch, world=world, lr=args.lr, warmup_steps=10, total_steps=args.steps, rank=rank)
    fd = FD.JpegFeeder(dev, args.batch)
    torch.manual_seed(11997733 + rank)                         # SEED + rank, like the reference's workers

    def batch_indices(step):
        idx = ddp.shard_indices(args.files, rank, world, train=True, epoch=step, seed=7)     # reshuffled every step
        return (idx * (args.batch // len(idx) + 1))[:args.batch]

    queue = [batch_indices(0), batch_indices(1)]
    for q in queue:
        fd.submit([files[i] for i in q])
    t0, losses = time.perf_counter(), []
    for step in range(args.steps):
        idx = queue.pop(0)
        y, c, q, flags, slot = fd.get()
        plans = tf.sample_plans(args.batch)
        x = tf.run(y, c, q, plans, clamp_in=flags, out=stage.x_static)
        fd.release(slot)
        loss = stage.step(x, labels[idx].to(dev))
        nxt = batch_indices(step + 2)
        queue.append(nxt)
        fd.submit([files[i] for i in nxt])
        if step % 10 == 0 or step == args.steps - 1:
            losses.append(float(loss))
            if rank == 0:
                print(f"step {step:4d}  loss {losses[-1]:.4f}  {(step + 1) * args.batch * world / (time.perf_counter() - t0):.0f} img/s", flush=True)
    while fd.pending:
        fd.release(fd.get()[4])
    torch.cuda.synchronize()
    fd.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return losses


if __name__ == "__main__":
    main()
