"""One training step of the DCT ViT on one B200 (SURVEY.md 3.1 hot loop, train.py:146-176):

    mixup (cls_transforms.py:135-182, on the embed input -- mixup is a convex combination and the embed tail is
    affine, so it commutes) -> ViT forward -> soft-label cross entropy -> backward into ONE flat fp32 gradient
    buffer -> [single NCCL allreduce of that buffer over NVLink, train.py:137's DDP buckets collapsed into one]
    -> clip_grad_norm_(1.0) + AdamW + decoupled WeightDecay in one fused kernel (train.py:163-172) -> bf16 weight refresh.

The forward/backward part and the optimiser part are captured in CUDA graphs (launch-bound otherwise: ~600 kernel
launches per step); the learning rate and the mixup lambda live in device memory so replays see fresh values."""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import ops as K
from . import vit as V

ARCHS = {"vitti": dict(emb_size=192, depth=12, num_heads=3, wd=1e-4), "vits": dict(emb_size=384, depth=12, num_heads=6, wd=3e-4),
         "vitb": dict(emb_size=768, depth=12, num_heads=12, wd=3e-4),
         # SwinV2-T DCT, window 8 (utils/configs.py:123-138; TRAIN.WD default :26): same loop, swin_train.SwinFlatEngine underneath
         "swinv2t": dict(swin=True, wd=3e-4)}


class TrainStage:
    RING = 8

    def __init__(self, device, arch: str = "vits", batch: int = 256, dtype: str = "bf16", world: int = 1, lr: float = 3e-3,
                 warmup_steps: int = 10000, total_steps: int = 112590, mixup_alpha: float = 0.2, use_graph: bool = True,
                 seed: int = 11997733, attention: str = "auto", rank: int = 0):
        if dtype != "bf16":
            raise NotImplementedError("rgbnm TrainStage: the tcgen05 path computes in bf16 (fp32 accumulation)")
        cfg = ARCHS[arch]
        self.dev = torch.device(device)
        self.B, self.world = batch, world
        # same initial replica on every rank (DDP's broadcast, train.py:137) WITHOUT touching the global RNG streams:
        # the augmentation plans are drawn from torch's global CPU generator, seeded SEED + rank by the caller (train.py:119)
        with torch.random.fork_rng(devices=[self.dev] if self.dev.type == "cuda" else []):
            torch.manual_seed(seed)
            if cfg.get("swin"):
                from . import swin as S
                self.model = S.SwinTransformerV2(img_size=256, patch_size=4, embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24],
                                                 window_size=8, mlp_ratio=4, drop_path_rate=0.2, pretrained_window_sizes=[0, 0, 0, 0],
                                                 device=self.dev, pixel_space="dct").train()
            else:
                self.model = V.ViT(patch_size=16, emb_size=cfg["emb_size"], depth=cfg["depth"], n_classes=1000, drop_p=0.0,
                                   pixel_space="DCT", ver=1, use_subblock=True, device=self.dev, num_heads=cfg["num_heads"],
                                   head_size=64, attention=attention)
        if cfg.get("swin"):
            from . import swin_train as ST
            self.eng = ST.SwinFlatEngine(self.model, self.dev)
            tokens, feat = 4096, 24
        else:
            self.eng = self.model.prepare(self.dev)
            tokens, feat = V.TOKENS, V.IN_FEAT
        self.base_lr, self.wd = lr, cfg["wd"]
        self.warmup_steps, self.total_steps = warmup_steps, total_steps
        self.m = torch.zeros_like(self.eng.flat)
        self.v = torch.zeros_like(self.eng.flat)
        self.gnorm = torch.zeros(1, dtype=torch.float32, device=self.dev)
        if self.dev.type == "cuda":
            K.sumsq(self.eng.flat_grad, self.gnorm)        # first call allocates the reduction scratch: must not happen under graph capture
            self.gnorm.zero_()
        # per-step scalars (9 optimiser hyper-parameters + 2 mixup weights) travel through a ring of pinned staging
        # buffers, each guarded by an event recorded after its H2D copy: the host may run many steps ahead of the GPU
        # (graph replays are asynchronous) and must not rewrite a buffer whose copy is still queued
        self._scalars_dev = torch.zeros(11, dtype=torch.float32, device=self.dev)
        self.hyper, self.lam = self._scalars_dev[:9], self._scalars_dev[9:]
        pin = self.dev.type == "cuda"
        self._ring = [torch.zeros(11, dtype=torch.float32).pin_memory() if pin else torch.zeros(11) for _ in range(self.RING)]
        self._ring_ev = [None] * self.RING
        self.mixup_alpha = mixup_alpha
        self.step_no = 0
        self.use_graph = use_graph
        self.g_fb: Optional[torch.cuda.CUDAGraph] = None
        self.g_opt: Optional[torch.cuda.CUDAGraph] = None
        self.x_static = torch.zeros((batch, tokens, feat), dtype=torch.bfloat16, device=self.dev)
        self.x_mixed = torch.zeros_like(self.x_static)
        self.y_static = torch.zeros((batch,), dtype=torch.int64, device=self.dev)
        self.loss = torch.zeros((), dtype=torch.float32, device=self.dev)
        self.launches_per_step = 0
        self._rng = torch.Generator().manual_seed(seed + 17 + rank)     # per-rank mixup stream (reference: SEED + rank)
        self.lam_override = None                                           # tests: inject (lam, 1 - lam) instead of drawing

    # ---- pieces ---------------------------------------------------------------------------------------------
    def _lr(self, step_no: int) -> float:
        """Learning rate the reference uses for its (0-based) `step_no`-th optimiser step (train.py:149-152, 174-176):
        `current_itr` is incremented BEFORE use, so it = step_no + 1; warm-up sets LR * (it + 1) / WARMUP while
        it < WARMUP; from it = WARMUP on the per-iteration CosineAnnealingLR (T_max = maxiters - WARMUP,
        pipeline_utils.py:538), stepped after the optimiser, has been stepped it - WARMUP times."""
        it = step_no + 1
        if it < self.warmup_steps:
            return self.base_lr * (it + 1) / self.warmup_steps
        t = (it - self.warmup_steps) / max(1, self.total_steps - self.warmup_steps)
        return 0.5 * self.base_lr * (1.0 + math.cos(math.pi * min(1.0, t)))

    def _fwd_bwd(self):
        x, y, lam = self.x_static, self.y_static, self.lam
        # RandomMixup_DCT: batch rolled by one, lambda from a sorted Dirichlet(alpha, alpha) draw
        xm = self.x_mixed
        K.mixup(x, xm, lam)
        onehot = torch.nn.functional.one_hot(y, 1000).float()
        soft = onehot * lam[0] + onehot.roll(1, 0) * lam[1]
        logits = self.eng.forward(xm)
        logp = torch.log_softmax(logits, dim=1)
        self.loss.copy_(-(soft * logp).sum(1).mean())
        dlogits = (torch.exp(logp) * soft.sum(1, keepdim=True) - soft) / self.B
        self.eng.backward(dlogits)

    def _opt(self):
        self.gnorm.zero_()
        K.sumsq(self.eng.flat_grad, self.gnorm)
        K.adamw_step(self.eng.flat, self.eng.flat_grad, self.m, self.v, self.eng.n_decay, self.gnorm, self.hyper)
        self.eng.refresh_weights()

    def _set_hyper(self):
        t = self.step_no + 1
        lr = self._lr(self.step_no)
        if self.lam_override is not None:
            lam = torch.tensor([float(self.lam_override), 1.0 - float(self.lam_override)])
        elif self.mixup_alpha > 0:
            lam = torch._sample_dirichlet(torch.tensor([self.mixup_alpha, self.mixup_alpha]), generator=self._rng)
            lam, _ = lam.sort(descending=True)
        else:
            lam = torch.tensor([1.0, 0.0])
        # sumsq runs after the SUM allreduce, so the norm and the gradient both carry the factor `world` (h[7])
        vals = torch.tensor([lr, 0.9, 0.999, 1e-8, 1.0 - 0.9 ** t, 1.0 - 0.999 ** t, lr / self.base_lr * self.wd,
                             1.0 / self.world, 1.0, float(lam[0]), float(lam[1])], dtype=torch.float32)
        k = self.step_no % self.RING
        if self._ring_ev[k] is not None:
            self._ring_ev[k].synchronize()                 # the copy that last read this staging buffer has completed
        self._ring[k].copy_(vals)
        self._scalars_dev.copy_(self._ring[k], non_blocking=True)
        if self.dev.type == "cuda":
            if self._ring_ev[k] is None:
                self._ring_ev[k] = torch.cuda.Event()
            self._ring_ev[k].record()

    # ---- public ---------------------------------------------------------------------------------------------
    def step(self, x: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        """x: (B,196,384) bf16 from FusedDCT ((B,4096,24) for swinv2t), labels (B,) int64.  Returns the device scalar loss of this step."""
        self._set_hyper()
        if x.data_ptr() != self.x_static.data_ptr():          # FusedDCT can write straight into x_static (out=stage.x_static)
            self.x_static.copy_(x, non_blocking=True)
        self.y_static.copy_(labels, non_blocking=True)
        if not self.use_graph:
            l0 = self.eng.launches
            self._fwd_bwd()
            self._allreduce()
            self._opt()
            self.launches_per_step = self.eng.launches - l0 + 2
        else:
            if self.g_fb is None:
                self._capture()
            self.g_fb.replay()
            self._allreduce()
            self.g_opt.replay()
        self.step_no += 1
        return self.loss

    def _allreduce(self):
        if self.world > 1:
            from . import ddp
            ddp.allreduce_flat(self.eng.flat_grad, self.world)   # one flat SUM allreduce; 1/world folded into the optimiser kernel (hyper[7])

    def _capture(self):
        # warm up on a side stream (lazy allocations, cuTensorMap encodes, cuBLAS handles of the tiny head ops)
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                self._fwd_bwd()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        l0 = self.eng.launches
        self.g_fb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_fb):
            self._fwd_bwd()
        l1 = self.eng.launches
        self.g_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_opt):
            self._opt()
        self.launches_per_step = (l1 - l0) + (self.eng.launches - l1) + 2
