"""Host side of the fused DCT data path (K0) -- the `get_transform` seam (boundary B2).

The reference builds a per-image `transforms.Compose` that runs on CPU DataLoader workers
(/root/reference/datasets.py:305-390).  Here the same `dataset`/`type`/`ops_list`/`num_ops`/
`ops_magnitude` arguments configure a *batch* transform: the host only draws the
augmentation plan (rgb_no_more_b200/plan.py, same RNG call sequence), everything else runs
in one fused CUDA kernel through the C-ABI (include/rgbnm_b200.h: rgbnm_k0_dcstats +
rgbnm_k0_fused).  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import lib as _lib
from . import plan as P

OUT_F32, OUT_BF16, OUT_INT16_PLANES = 0, 1, 2
LAYOUT_VIT16, LAYOUT_SWIN4, LAYOUT_VIT16_NOSUB = 0, 1, 2
PLANE_ELEMS = (28 * 28 + 2 * 14 * 14) * 64
# out_size (luma blocks per side after the resize) -> (layout, tokens, features per token)
GEOMETRY = {28: (LAYOUT_VIT16, 196, 384), 32: (LAYOUT_SWIN4, 4096, 24)}


class FusedDCT:
    """Batch DCT transform on one GPU.

    >>> tf = FusedDCT(device, kind="train", ops_list=plan.AUGLIST_VITS, num_ops=2, ops_magnitude=9)
    >>> x = tf(y_q, c_q, quant)            # (B,196,384) operand of the patch-projection Linear

    `subblock=False` (ViT only) writes the `--no_subblock` operand: the same token layout with the un-converted 16 x 16 luma tile
    (four 8 x 8 blocks, no A16 products).  `out_size=32` selects the SwinV2 data path (datasets.py:370-382, models/swinv2.py:505-576): planes resized to
    32 x 32 luma blocks, every 8 x 8 block decomposed into 4 x 4 (Y) / 2 x 2 (CbCr) sub-blocks -> (B, 4096, 24).
    """

    def __init__(self, device, kind: str = "test", ops_list: Optional[Sequence[str]] = None, num_ops: int = 2,
                 ops_magnitude: int = 10, out_dtype: torch.dtype = torch.float32, out_size: int = 28, subblock: bool = True):
        if out_size not in GEOMETRY:
            raise NotImplementedError("rgbnm: only the 28-block (ViT, patch 16) and 32-block (SwinV2, patch 4) geometries "
                                      "are on the hot path")
        self.out_size = out_size
        self.layout, self.tokens, self.feat = GEOMETRY[out_size]
        if not subblock:
            # `--no_subblock` (use_subblock = False, plainvit.py:173-216): the luma tile of a token stays un-converted
            if out_size != 28:
                raise NotImplementedError("rgbnm: SwinV2's patch-4 embedding always decomposes blocks (models/swinv2.py:505-576)")
            self.layout = LAYOUT_VIT16_NOSUB
        self.plane_elems = (out_size * out_size + 2 * (out_size // 2) ** 2) * 64
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.RgbnmError("rgbnm: FusedDCT needs a CUDA device; there is no CPU fallback")
        self.kind = kind
        self.ops_list = list(ops_list) if ops_list is not None else None
        self.num_ops = num_ops
        self.ops_magnitude = ops_magnitude
        self.out_dtype = out_dtype
        self.bank = P.FilterBank()
        self._lib = _lib.load()
        self._filters_dev = None
        self._filters_n = -1
        self._lut = torch.from_numpy(P.posterize_lut()).to(self.device)
        self._tables = _lib.K0Tables()
        self._eq_lut = None
        self._stats = None
        self._sync_tables()

    # -- tables ---------------------------------------------------------------------------
    def _sync_tables(self):
        if self._filters_n != self.bank._n:
            self._filters_dev = torch.from_numpy(self.bank.table.reshape(-1, 64).copy()).to(self.device)
            self._filters_n = self.bank._n
        self._tables.filters = self._filters_dev.data_ptr()
        self._tables.posterize_lut = self._lut.data_ptr()

    # -- plans ----------------------------------------------------------------------------
    def sample_plans(self, n: int, hb: int = 64, wb: int = 64) -> List[P.Plan]:
        """Draw n plans from the torch global CPU RNG in the reference's call order."""
        if self.kind == "train":
            return [P.sample_train_plan(hb, wb, self.ops_list, self.num_ops, self.ops_magnitude, self.bank, size=self.out_size)
                    for _ in range(n)]
        return [P.eval_plan(hb, wb) if self.out_size == 28 else P.eval_plan_swin(hb, wb, self.out_size)] * n

    def sample_plans_packed(self, n: int, hb: int = 64, wb: int = 64, clamp_in=None, generator=None) -> np.ndarray:
        """n plans, packed (`struct rgbnm_plan[n]`), drawn by plan.BatchedSampler: the same distributions as `sample_plans`, one
        vectorised draw per batch instead of the reference's per-image RNG call sequence (~0.3 ms instead of ~15 ms per 256)."""
        if self.kind != "train":
            return P.pack_plans(self.sample_plans(n, hb, wb), clamp_in, out_size=self.out_size)
        key = (hb, wb)
        if getattr(self, "_batched", None) is None or self._batched[0] != key:
            self._batched = (key, P.BatchedSampler(hb, wb, self.ops_list if self.ops_list is not None else P.AUGLIST_VITS,
                                                   self.num_ops, self.ops_magnitude, self.bank, size=self.out_size))
        return self._batched[1].sample(n, clamp_in=clamp_in, generator=generator)

    # -- launch ---------------------------------------------------------------------------
    def run(self, y_q: torch.Tensor, c_q: torch.Tensor, quant: torch.Tensor, plans, clamp_in=None,
            out_mode: Optional[int] = None, out: Optional[torch.Tensor] = None,
            plans_dev: Optional[torch.Tensor] = None, needs_stats: Optional[bool] = None, timing: Optional[list] = None) -> torch.Tensor:
        """y_q int16 [B,hb,wb,64] (or [B,1,hb,wb,8,8]), c_q int16 [B,2,hb/2,wb/2,64], quant int16 [B,3,64],
        all on `device`.  `plans`: list[Plan] or packed numpy array (or pass `plans_dev`).
        `needs_stats=False` (known on the host: no plan of the batch holds a statistics op, e.g. the eval transform) skips the
        DC-statistics pre-pass launch; None = decide from `plans` when they are given as a list, else launch it.
        `timing`: a list that receives (before, between, after) CUDA events recorded around the two launches (bench.py)."""
        B = y_q.shape[0]
        if y_q.dim() == 6:
            hb, wb = y_q.shape[2], y_q.shape[3]
        else:
            hb, wb = y_q.shape[1], y_q.shape[2]
        for t in (y_q, c_q, quant):
            if t.device != self.device or t.dtype != torch.int16 or not t.is_contiguous():
                raise ValueError("rgbnm: coefficient tensors must be contiguous int16 on the transform's device")
        if c_q.numel() != B * 2 * (hb // 2) * (wb // 2) * 64 or quant.numel() != B * 192:
            raise ValueError("rgbnm: coefficient batch shapes are inconsistent (4:2:0 layout expected)")
        if needs_stats is None:
            needs_stats = (self.kind == "train") if (plans_dev is not None or isinstance(plans, np.ndarray)) \
                else any(pl.needs_stats for pl in plans)
        if plans_dev is None:
            if not isinstance(plans, np.ndarray):
                for pl in plans:
                    if pl.crop_i + pl.crop_size > hb or pl.crop_j + pl.crop_size > wb or pl.crop_i < 0 or pl.crop_j < 0:
                        raise ValueError("rgbnm: crop window outside the image")
                plans = P.pack_plans(plans, clamp_in, out_size=self.out_size)
            if len(plans) != B:
                raise ValueError("rgbnm: one plan per image required")
            # filter ops carry an index into THIS transform's filter bank (Sharpness / MidfreqAug): plans sampled with
            # another FusedDCT would silently read other filters
            codes, first = plans["ops"]["code"], plans["ops"]["p"][..., 0]
            filt = (codes == P.OP_SHARPNESS) | (codes == P.OP_MIDFREQ)
            used = np.arange(codes.shape[1])[None, :] < plans["n_ops"][:, None]
            if bool((filt & used & (first >= self.bank._n)).any()):
                raise ValueError("rgbnm: a plan refers to a filter this transform's bank does not hold "
                                 "(plans must be sampled by the FusedDCT that runs them)")
            plans_dev = torch.from_numpy(plans.view(np.uint8).reshape(B, -1)).to(self.device, non_blocking=True)
        self._sync_tables()
        if out_mode is None:
            out_mode = OUT_BF16 if self.out_dtype == torch.bfloat16 else OUT_F32
        if out is None:
            if out_mode == OUT_INT16_PLANES:
                out = torch.empty((B, self.plane_elems), dtype=torch.int16, device=self.device)
            else:
                out = torch.empty((B, self.tokens, self.feat), dtype=torch.bfloat16 if out_mode == OUT_BF16 else torch.float32,
                                  device=self.device)
        # scratch of the statistics pre-pass: persistent (no allocation / memset per call).  Every slot the fused kernel reads
        # (ops of plans with needs_stats) is rewritten by the pre-pass of the same call.
        if self._stats is None or self._stats.shape[0] < B:
            self._stats = torch.zeros((B, P.MAX_OPS, 2), dtype=torch.float32, device=self.device)
        stats = self._stats
        if self._eq_lut is None or self._eq_lut.shape[0] < B:      # scratch of the Equalize op (per image and op slot)
            self._eq_lut = torch.empty((B, P.MAX_OPS, 2048), dtype=torch.int16, device=self.device)
        self._tables.equalize_lut = self._eq_lut.data_ptr()
        st = _lib.stream_ptr()
        L = self._lib
        if timing is not None:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
        if needs_stats:
            _lib.check(L.rgbnm_k0_dcstats_ex(y_q.data_ptr(), c_q.data_ptr(), quant.data_ptr(), plans_dev.data_ptr(),
                                             C.byref(self._tables), stats.data_ptr(), B, hb, wb, self.layout, st), "rgbnm_k0_dcstats")
        if timing is not None:
            ev[1].record()
        _lib.check(L.rgbnm_k0_fused_ex(y_q.data_ptr(), c_q.data_ptr(), quant.data_ptr(), plans_dev.data_ptr(),
                                       C.byref(self._tables), stats.data_ptr(), out.data_ptr(), out_mode, self.layout, B, hb, wb, st),
                   "rgbnm_k0_fused")
        if timing is not None:
            ev[2].record()
            timing.append(tuple(ev))
        self.last_stats = stats
        return out

    def __call__(self, y_q, c_q, quant, clamp_in=None):
        plans = self.sample_plans(y_q.shape[0], *(y_q.shape[2:4] if y_q.dim() == 6 else y_q.shape[1:3]))
        return self.run(y_q, c_q, quant, plans, clamp_in)


def split_planes(planes: torch.Tensor, out_size: int = 28):
    """INT16_PLANES output -> (Y [B,1,S,S,8,8], CbCr [B,2,S/2,S/2,8,8]) in the reference layout (S = 28 or 32)."""
    B = planes.shape[0]
    S, H = out_size, out_size // 2
    y = planes[:, : S * S * 64].reshape(B, 1, S, S, 8, 8)
    c = planes[:, S * S * 64:].reshape(B, 2, H, H, 8, 8)
    return y, c


def get_transform(dataset: str = "imagenet_dct", type: str = "train", ops_list=None, num_ops: int = 2,
                  ops_magnitude: int = 10, dtype=torch.float32, device="cuda", subblock: bool = True):
    """Same selector as the reference's datasets.get_transform (datasets.py:305-390) for the
    DCT datasets; returns a batch transform bound to `device`."""
    if dataset not in ("imagenet_dct", "imagenet_dct_swin"):
        raise NotImplementedError(f"rgbnm: dataset '{dataset}' is outside the B200 hot path (SURVEY.md 8f)")
    size = 28 if dataset == "imagenet_dct" else 32          # datasets.py:355-366 / :370-382
    if type == "train":
        return FusedDCT(device, "train", ops_list, num_ops, ops_magnitude, dtype, out_size=size, subblock=subblock)
    if type in ("val", "test"):
        return FusedDCT(device, "test", None, 0, 0, dtype, out_size=size, subblock=subblock)
    print("Unrecognized dataset type! Returning 'None' transform")
    return None
