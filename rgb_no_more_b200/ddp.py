"""Data-parallel plumbing of the DCT ViT path: one process per GPU, images sharded across ranks,
ONE all-reduce of the flat fp32 gradient buffer per step.

Reference behaviour mirrored here:
  * train shards  = torch DistributedSampler (datasets.py:533-541, train.py:143 `set_epoch`): seeded permutation,
    padded by wrapping around so every rank gets ceil(n / world) indices, rank takes every world-th index;
  * eval shards   = DistributedEvalSampler (utils/custom_sampler.py:88): strided, NOT padded (no image counted twice);
  * gradients     = DistributedDataParallel mean over ranks (train.py:137).  DDP's bucketed all-reduces collapse into
    one SUM all-reduce of the flat buffer; the 1/world factor is folded into the optimiser kernel (hyper[7]).
torch.distributed is plumbing only: NCCL over NVLink on the GPUs, gloo in the CPU tests."""
from __future__ import annotations

from typing import List

import torch


def shard_indices(n: int, rank: int, world: int, train: bool, epoch: int = 0, seed: int = 0, shuffle: bool = True) -> List[int]:
    if not (0 <= rank < world):
        raise ValueError("rank must be in [0, world)")
    if train:
        if shuffle:
            g = torch.Generator()
            g.manual_seed(seed + epoch)
            idx = torch.randperm(n, generator=g).tolist()
        else:
            idx = list(range(n))
        total = -(-n // world) * world
        pad = total - len(idx)
        if pad > 0:
            idx += (idx * (-(-pad // max(1, len(idx)))))[:pad]
        return idx[rank:total:world]
    return list(range(rank, n, world))


def allreduce_flat(flat_grad: torch.Tensor, world: int) -> float:
    """SUM all-reduce of the flat gradient buffer in place; returns the factor the optimiser must apply (1 / world)."""
    if world > 1:
        import torch.distributed as dist
        if flat_grad.is_cuda and dist.get_backend() != "nccl":
            # gloo (CPU tests, single-GPU multi-rank tests): stage through the host; the product path is NCCL over NVLink
            host = flat_grad.cpu()
            dist.all_reduce(host)
            flat_grad.copy_(host)
        else:
            dist.all_reduce(flat_grad)
    return 1.0 / world
