"""Thin host wrappers of the memory-bound ViT kernels (include/rgbnm_b200.h, csrc/vit_kernels.cu).
Every function launches on torch's current stream and raises if the CUDA library is missing:
there is no CPU fallback on the product path."""
from __future__ import annotations

from typing import Optional

import torch

from . import lib as _lib


def _L():
    return _lib.load()


def layernorm_fwd(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, y: torch.Tensor, mean: torch.Tensor,
                  rstd: torch.Tensor, eps: float = 1e-5) -> None:
    rows, emb = x.shape
    _lib.check(_L().rgbnm_layernorm_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), mean.data_ptr(),
                                        rstd.data_ptr(), rows, emb, eps, _lib.stream_ptr()), "rgbnm_layernorm_fwd")


def layernorm_bwd(dy: torch.Tensor, x: torch.Tensor, mean: torch.Tensor, rstd: torch.Tensor, gamma: torch.Tensor,
                  dres: Optional[torch.Tensor], dx: torch.Tensor, dgamma: torch.Tensor, dbeta: torch.Tensor,
                  dxsum: Optional[torch.Tensor] = None, rows_per_dy_row: int = 1) -> None:
    """dxsum (fp32 [emb], optional) += column sums of dx: the bias gradient of the Linear feeding this residual stream.
    rows_per_dy_row > 1: `dy` has rows / rows_per_dy_row rows, each shared by that many consecutive rows of x."""
    rows, emb = x.shape
    if dy.shape[0] * rows_per_dy_row != rows or not dy.is_contiguous():
        raise ValueError("rgbnm layernorm_bwd: dy must hold rows / rows_per_dy_row contiguous rows")
    _lib.check(_L().rgbnm_layernorm_bwd_ex(dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                                           None if dres is None else dres.data_ptr(), dx.data_ptr(), dgamma.data_ptr(),
                                           dbeta.data_ptr(), None if dxsum is None else dxsum.data_ptr(), rows, emb,
                                           rows_per_dy_row, _lib.stream_ptr()), "rgbnm_layernorm_bwd")


def mixup(x: torch.Tensor, out: torch.Tensor, lam: torch.Tensor) -> None:
    """out[b] = lam[0] * x[b] + lam[1] * x[b-1] on a contiguous bf16 batch (RandomMixup_DCT, cls_transforms.py:135-182)."""
    if x.dtype != torch.bfloat16 or not x.is_contiguous() or not out.is_contiguous() or out.shape != x.shape:
        raise ValueError("rgbnm mixup: contiguous bf16 tensors of equal shape expected")
    _lib.check(_L().rgbnm_mixup_bf16(x.data_ptr(), out.data_ptr(), lam.data_ptr(), x.shape[0], x[0].numel(), _lib.stream_ptr()),
               "rgbnm_mixup_bf16")


def colsum(a: torch.Tensor, out: torch.Tensor, qkv_heads: int = 0, head_dim: int = 0) -> None:
    """out += column sums of a.  qkv_heads > 0: columns of `a` are in the kernel's q|k|v order, `out` in reference order."""
    rows, cols = a.shape
    _lib.check(_L().rgbnm_colsum_bf16(a.data_ptr(), a.stride(0), rows, cols, out.data_ptr(), qkv_heads, head_dim,
                                      _lib.stream_ptr()), "rgbnm_colsum_bf16")


def weight_prep(w: torch.Tensor, wb: torch.Tensor, wt: Optional[torch.Tensor], qkv_heads: int = 0, head_dim: int = 0) -> None:
    n, k = w.shape
    _lib.check(_L().rgbnm_weight_prep(w.data_ptr(), n, k, qkv_heads, head_dim, wb.data_ptr(),
                                      None if wt is None else wt.data_ptr(), _lib.stream_ptr()), "rgbnm_weight_prep")


def qkv_perm_vec(src: torch.Tensor, dst: torch.Tensor, heads: int, head_dim: int, inverse: bool) -> None:
    _lib.check(_L().rgbnm_qkv_perm_vec(src.data_ptr(), dst.data_ptr(), src.numel(), heads, head_dim, int(inverse),
                                       _lib.stream_ptr()), "rgbnm_qkv_perm_vec")


def qkv_unperm_rows_add(src: torch.Tensor, dst: torch.Tensor, heads: int, head_dim: int) -> None:
    n, k = src.shape
    _lib.check(_L().rgbnm_qkv_unperm_rows_add(src.data_ptr(), dst.data_ptr(), n, k, heads, head_dim, _lib.stream_ptr()),
               "rgbnm_qkv_unperm_rows_add")


def sumsq(g: torch.Tensor, out: torch.Tensor) -> None:
    _lib.check(_L().rgbnm_sumsq_f32(g.data_ptr(), g.numel(), out.data_ptr(), _lib.stream_ptr()), "rgbnm_sumsq_f32")


def adamw_step(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, n_decay: int, gnorm_sq: torch.Tensor,
               hyper: torch.Tensor) -> None:
    """hyper: 9 fp32 on the device -- lr, beta1, beta2, eps, 1-beta1^t, 1-beta2^t, decay, grad_scale, max_norm."""
    _lib.check(_L().rgbnm_adamw_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), n_decay,
                                     gnorm_sq.data_ptr(), hyper.data_ptr(), _lib.stream_ptr()), "rgbnm_adamw_step")
