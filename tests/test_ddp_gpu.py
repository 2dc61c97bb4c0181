"""Row a32 / SURVEY.md section 4 (4): multi-rank correctness of the data-parallel step ON the GPU.

After N steps (i) every rank holds bit-identical parameters (DDP's invariant, train.py:137), (ii) the 2-rank run at
global batch 16 follows the 1-rank run at batch 16: per-step loss (mean of the rank losses) and parameter deltas agree
to bf16-gradient accuracy.  Runs with NCCL on 2 GPUs when the box has them (`gpurun --gpus 2`), else with both ranks on
cuda:0 and gloo carrying the all-reduce (same TrainStage / ddp.allreduce_flat code path)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "harness", "ddp_gpu_worker.py")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(world, out, one_device, steps=5):
    if world == 1:
        cmd = [sys.executable, WORKER, "--out", out, "--steps", str(steps)]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", str(_free_port()), WORKER, "--out", out, "--steps", str(steps)] + (["--one-device"] if one_device else [])
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    env.setdefault("GLOO_SOCKET_IFNAME", "lo")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    return [torch.load(f"{out}.rank{k}.pt") for k in range(world)]


def test_two_ranks_identical_replicas_and_match_one_rank(tmp_path):
    one_device = torch.cuda.device_count() < 2
    two = _run(2, str(tmp_path / "w2"), one_device)
    one = _run(1, str(tmp_path / "w1"), False)[0]
    # (i) replicas stay bit-identical
    assert torch.equal(two[0]["flat"], two[1]["flat"])
    # (ii) same trajectory as one rank at the same global batch
    l2 = np.mean([two[0]["losses"], two[1]["losses"]], axis=0)
    l1 = np.array(one["losses"])
    assert np.abs(l2 - l1).max() < 2e-2 * np.abs(l1).max(), (l2, l1)          # bf16 forward, different batch split
    st0 = _initial_flat()
    d2, d1 = (two[0]["flat"] - st0).double(), (one["flat"] - st0).double()
    cos = float((d2 * d1).sum() / (d2.norm() * d1.norm()))
    # Adam's first steps are sign-like (m / sqrt(v) = +-1): elements whose tiny gradients differ in the last bf16 bits flip
    assert cos > 0.9, cos
    assert abs(float(d2.norm() / d1.norm()) - 1.0) < 0.05


def _initial_flat():
    from rgb_no_more_b200 import train_step as TS
    st = TS.TrainStage("cuda:0", arch="vits", batch=8, warmup_steps=10, total_steps=1000, mixup_alpha=0.0, use_graph=False)
    return st.eng.flat.detach().cpu().clone()
