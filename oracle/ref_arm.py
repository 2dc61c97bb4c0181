"""Baseline arms that run the REFERENCE ITSELF (oracle/_ref, staged by oracle/make_ref.py).  BASELINE INFRASTRUCTURE ONLY:
imported by bench.py's `--impl reference`, `cpu_baseline` and `torch_b200_baseline` legs and by tests/, never by the product.

What is the reference's and what is ours in these arms:
  * the reference's own code, unmodified: `datasets.dataset_selector` (DataLoader worker processes, `imagenet_dataset_indexing`
    .__getitem__ with its dequantise + clamp, `get_transform('imagenet_dct', 'train', ...)`: RandomResizedCrop_DCT / RandomFlip_DCT /
    RandAugment_dct / ToRange), `utils.get_mixup` (RandomMixup_DCT), `utils.get_model` (models.plainvit.ViT),
    `utils.get_optim_and_criterion` (CrossEntropyLoss, AdamW, custom_optims.WeightDecay, CosineAnnealingLR), `utils.unpack_data`;
  * restated here (15 lines): the loop body of train.py:146-176 -- `traineval` itself needs NCCL + DDP + a GPU per rank;
  * ours: `dct_manip.read_coefficients` -- the reference's native module needs jpeglib.h and cannot be built in this image
    (DESIGN.md section 2), so its Huffman-decode leg is served by this repository's host decoder through the B1 drop-in;
    stand-ins for yacs / timm / torchinfo / sysrsync (rgb_no_more_b200/compat/shims)."""
from __future__ import annotations

import os
import sys
import tempfile
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "utils", "pipeline_utils.py"))


def _activate():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from rgb_no_more_b200.compat import env
    env.activate(REF)
    import datasets as ds
    import utils.configs as configs
    import utils.pipeline_utils as U
    return configs, U, ds


def _config(configs, U, arch, batch, rank, threads):
    cfg = configs.generate_config(modelarch=arch, domain="dct", modelver=1, subblock=True, batchsize=batch, num_ops=2,
                                  ops_magnitude=9, warmup_steps=10000)
    return U.update_config(cfg, rank, 1, threads, True, False, False)


def write_dataset(workdir: str, n_files: int, n_rows: int) -> str:
    """`n_files` synthetic 512x512 Q75 4:2:0 baseline JPEGs (rgb_no_more_b200.synth) + a `Filepath,Label` index of
    `n_rows` rows cycling over them (assets/indexbase_val.csv format)."""
    from rgb_no_more_b200 import synth
    os.makedirs(workdir, exist_ok=True)
    paths = []
    for i in range(n_files):
        p = os.path.join(workdir, f"img_{i:05d}.JPEG")
        if not os.path.exists(p):
            with open(p, "wb") as f:
                f.write(synth.synth_jpeg(i))
        paths.append(p)
    index = os.path.join(workdir, f"index_{n_rows}.csv")
    with open(index, "w") as f:
        f.write("Filepath,Label\n")
        for r in range(n_rows):
            f.write(f"{paths[r % n_files]},{r % 1000}\n")
    return index


def train_loop_body(cfg, U, model, data, criterion, optimizer, weight_decayer, cosinescheduler, mixup, state):
    """train.py:146-176, the non-AMP branch (the reference's ViT-S / ViT-Ti configs train with AMP off, configs.py:86-103)."""
    inputs, labels = U.unpack_data(data, cfg.TRAIN.DATASET, cfg.RANK, mixup)
    optimizer.zero_grad()
    weight_decayer.zero_grad()
    state["current_itr"] += 1
    if state["current_itr"] < cfg.TRAIN.WARMUP:
        U.adjust_lr(optimizer, cfg.TRAIN.LR * (state["current_itr"] + 1) / cfg.TRAIN.WARMUP)
        U.copy_lr(optimizer, weight_decayer)
    outputs = model(inputs[0], inputs[1])
    loss = criterion(outputs, labels)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1)
    optimizer.step()
    weight_decayer.step()
    if state["current_itr"] >= cfg.TRAIN.WARMUP:
        cosinescheduler.step()
        U.copy_lr(optimizer, weight_decayer)
    return loss


def cpu_train_arm(arch: str, images_per_step: int, steps: int, warmup: int, cores: int, n_files: int = 64):
    """`warmup + steps` training steps of the reference on the host: its DataLoader (`cores` worker processes: decode ->
    dequantise -> DCT transforms, one torch thread each, datasets.py:542-556) feeding its ViT train step on `cores` torch threads.
    Returns seconds per timed step and the split (waiting for the loader, model step)."""
    configs, U, ds = _activate()
    torch.set_num_threads(cores)
    cfg = _config(configs, U, arch, images_per_step, "cpu", cores)
    work = tempfile.mkdtemp(prefix="rgbnm_ref_")
    index = write_dataset(work, min(n_files, images_per_step), images_per_step * (warmup + steps))
    # trainval_split just above 0 selects the branch real training uses (the one that forwards ops_magnitude, datasets.py:506-560)
    # with an empty validation split
    loaders = ds.dataset_selector(dataset=cfg.TRAIN.DATASET, type="train", indexpath=index, basepath="", batch_size=images_per_step,
                                  num_workers=cores, shuffle=True, trainval_split=1e-9, distributed=True, rank=0, world_size=1,
                                  seed=cfg.SEED, ops_list=cfg.TRAIN.AUGLIST, num_ops=cfg.TRAIN.NUMOPS, ops_magnitude=cfg.TRAIN.AUGSTR)
    trainloader = loaders[0]
    torch.manual_seed(cfg.SEED)
    model = U.get_model(cfg, report=False)
    criterion, optimizer, weight_decayer, cosinescheduler, _ = U.get_optim_and_criterion(cfg, model, trainloader)
    mixup = U.get_mixup(cfg)
    model.train()
    state = {"current_itr": 0}
    t_wait, t_model, losses = [], [], []
    it = iter(trainloader)
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        data = next(it)
        t1 = time.perf_counter()
        loss = train_loop_body(cfg, U, model, data, criterion, optimizer, weight_decayer, cosinescheduler, mixup, state)
        losses.append(float(loss.detach()))
        t2 = time.perf_counter()
        if s >= warmup:
            t_wait.append(t1 - t0)
            t_model.append(t2 - t1)
    del it
    import shutil
    shutil.rmtree(work, ignore_errors=True)
    return {"sec_per_step": float(np.mean(t_wait) + np.mean(t_model)), "sec_loader_wait": float(np.mean(t_wait)),
            "sec_model": float(np.mean(t_model)), "losses": losses, "amp": bool(cfg.TRAIN.AMP)}


def calibrate_images_per_step(arch: str, cores: int, want: int, steps_total: int, budget_s: float) -> int:
    """Largest multiple of 32 (<= want) whose `steps_total` CPU steps fit `budget_s`, from one timed 16-image model step."""
    configs, U, ds = _activate()
    torch.set_num_threads(cores)
    cfg = _config(configs, U, arch, 16, "cpu", cores)
    model = U.get_model(cfg, report=False).train()
    y, c = torch.rand(16, 1, 28, 28, 8, 8), torch.rand(16, 2, 14, 14, 8, 8)
    for k in range(2):
        t0 = time.perf_counter()
        model(y, c).square().mean().backward()
        dt = time.perf_counter() - t0
    per_img = dt / 16 * 1.15                      # + optimiser / loader share
    fit = int(budget_s / max(1e-9, steps_total * per_img))
    return max(32, min(want, fit // 32 * 32))


def torch_b200_arm(arch: str, batch: int, steps: int, warmup: int, device: int = 0):
    """BASELINE.md section 4 item 5 / SURVEY.md 8d (iv): the reference's own model and optimiser under plain PyTorch on the B200
    (what the hand-written kernels must beat): train step (fwd + CE + bwd + clip + AdamW + WeightDecay) and eval forward, eager fp32
    and bf16 autocast, inputs resident on the device, CUDA events with synchronisation (benchmark.py:125-197 omits the sync)."""
    configs, U, ds = _activate()
    torch.cuda.set_device(device)
    cfg = _config(configs, U, arch, batch, device, 1)
    torch.manual_seed(cfg.SEED)
    model = U.get_model(cfg, report=False).to(device)
    criterion, optimizer, weight_decayer, cosinescheduler, _ = U.get_optim_and_criterion(cfg, model, range(1000))
    mixup = U.get_mixup(cfg)
    g = torch.Generator(device="cuda").manual_seed(1)
    y = torch.rand((batch, 1, 28, 28, 8, 8), device="cuda", generator=g) * 2 - 1
    c = torch.rand((batch, 2, 14, 14, 8, 8), device="cuda", generator=g) * 2 - 1
    labels = torch.randint(0, 1000, (batch,), device="cuda", generator=g)
    out = {}
    for name, amp in (("fp32", False), ("bf16_autocast", True)):
        state = {"current_itr": 0}

        def step():
            (yy, cc), soft = mixup((y, c), labels)
            optimizer.zero_grad()
            weight_decayer.zero_grad()
            state["current_itr"] += 1
            U.adjust_lr(optimizer, cfg.TRAIN.LR * (state["current_itr"] + 1) / cfg.TRAIN.WARMUP)
            U.copy_lr(optimizer, weight_decayer)
            with torch.autocast("cuda", enabled=amp, dtype=torch.bfloat16):
                loss = criterion(model(yy, cc), soft)
            loss.backward()                                     # bf16 autocast needs no GradScaler
            torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1)
            optimizer.step()
            weight_decayer.step()
            return loss

        def fwd():
            with torch.no_grad(), torch.autocast("cuda", enabled=amp, dtype=torch.bfloat16):
                return model(y, c)
        for fn, key in ((step, "train"), (fwd, "eval_forward")):
            model.train() if key == "train" else model.eval()
            for _ in range(warmup):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[f"{key}_{name}"] = {"ms_per_step": ms, "images_per_s": batch / (ms * 1e-3)}
    out["what"] = (f"reference models.plainvit.ViT ({arch}, DCT, embed_type 1) + reference optimiser objects under plain PyTorch "
                   f"{torch.__version__} on one B200, batch {batch}, inputs (ToRange'd planes) resident on the device -- no data path")
    return out
