"""N > 1 host logic on CPU (gloo, world_size 2): index sharding and the single flat-gradient all-reduce
(reference train.py:137,143, datasets.py:533-541, utils/custom_sampler.py:88)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rgb_no_more_b200 import ddp


def test_train_shards_match_distributed_sampler():
    from torch.utils.data.distributed import DistributedSampler
    data = list(range(103))
    for world in (2, 8):
        for rank in range(world):
            ref = DistributedSampler(data, num_replicas=world, rank=rank, shuffle=True, seed=11997733)
            ref.set_epoch(3)
            assert ddp.shard_indices(len(data), rank, world, train=True, epoch=3, seed=11997733) == list(iter(ref))


def test_eval_shards_are_a_partition():
    n, world = 50001, 8
    shards = [ddp.shard_indices(n, r, world, train=False) for r in range(world)]
    assert sorted(i for s in shards for i in s) == list(range(n))           # every image exactly once, no padding
    assert max(map(len, shards)) - min(map(len, shards)) <= 1


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(rank)
        flat = torch.randn(1000)
        mine = flat.clone()
        scale = ddp.allreduce_flat(flat, world)
        gathered = [torch.zeros(1000) for _ in range(world)]
        dist.all_gather(gathered, mine)
        mean = torch.stack(gathered).mean(0)
        ok = torch.allclose(flat * scale, mean, atol=1e-6) and scale == 1.0 / world
        out.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_is_the_ddp_mean():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == {0: True, 1: True}
