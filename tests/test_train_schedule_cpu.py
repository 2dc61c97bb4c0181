"""Row a31 host logic: the learning-rate / decoupled-decay trace of TrainStage equals the reference loop's
(train.py:146-176: `current_itr += 1` before use, warm-up `LR * (current_itr + 1) / WARMUP`, CosineAnnealingLR stepped
after the optimiser from current_itr >= WARMUP on; pipeline_utils.py:536-538; custom_optims.py:37-43)."""
import math

import torch

from rgb_no_more_b200.train_step import TrainStage


def _reference_trace(base_lr, warmup, maxiters, n):
    """The reference's loop body with a real torch optimiser + scheduler; returns the lr in force at every optimiser step."""
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([p], lr=base_lr, weight_decay=0, eps=1e-8)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=maxiters - warmup, eta_min=0)
    current_itr, lrs = 0, []
    for _ in range(n):
        current_itr += 1
        if current_itr < warmup:
            for g in opt.param_groups:
                g["lr"] = base_lr * (current_itr + 1) / warmup
        p.grad = torch.ones(1)
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        if current_itr >= warmup:
            sched.step()
    return lrs


def test_lr_trace_matches_reference_loop():
    st = TrainStage.__new__(TrainStage)
    st.base_lr, st.warmup_steps, st.total_steps = 3e-3, 7, 40
    ref = _reference_trace(3e-3, 7, 40, 40)
    ours = [st._lr(i) for i in range(40)]
    assert ours[0] == 3e-3 * 2 / 7                      # first step: current_itr = 1 -> LR * 2 / WARMUP
    for a, b in zip(ours, ref):
        assert math.isclose(a, b, rel_tol=1e-6, abs_tol=1e-12), (ours, ref)
