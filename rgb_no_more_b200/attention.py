"""Attention core of the DCT ViT: O = softmax(Q K^T / sqrt(emb_size)) V per (image, head)
(/root/reference/models/plainvit.py:450-461 -- note the reference scales by sqrt(emb_size), not sqrt(head_dim)).

Input is the fused projection output in the kernel-side layout qkv[B*N, 3*H*D] = [q | k | v], each H*D wide and
head-major (the "(h d qkv)" interleave of plainvit.py:447 is undone by regrouping the weight rows once per
step, rgbnm_weight_prep).  Output o[B*N, H*D] is already 'b n (h d)' (plainvit.py:461).

Backends:
  "b200"   hand-written tcgen05 kernels (csrc/attention_tc.cu) through the C-ABI
  "torch"  torch SDPA on the same layout -- library baseline kept for A/B numerics tests and as the stepping stone
           BASELINE.json config 2 names ("fused DCT kernel + torch attention")
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn.functional as F

from . import lib as _lib


def _have_b200() -> bool:
    L = _lib.load()
    return hasattr(L, "rgbnm_attention_fwd")


def _split(qkv: torch.Tensor, B: int, H: int, D: int):
    N = qkv.shape[0] // B
    v = qkv.view(B, N, 3, H, D).permute(2, 0, 3, 1, 4)       # (3, B, H, N, D) strided views
    return v[0], v[1], v[2]


def forward(qkv, o, lse, B, H, D, scale, backend="auto"):
    if backend == "auto":
        backend = "b200" if _have_b200() else "torch"
    if backend == "b200":
        L = _lib.load()
        _lib.check(L.rgbnm_attention_fwd(qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), B, qkv.shape[0] // B, H, D,
                                         C.c_float(scale), _lib.stream_ptr()), "rgbnm_attention_fwd")
        return
    q, k, v = _split(qkv, B, H, D)
    out = F.scaled_dot_product_attention(q, k, v, scale=scale)           # (B, H, N, D)
    o.view(B, -1, H, D).copy_(out.transpose(1, 2))


def backward(do, qkv, o, lse, dqkv, B, H, D, scale, backend="auto", dvec=None):
    if backend == "auto":
        backend = "b200" if _have_b200() else "torch"
    if backend == "b200":
        L = _lib.load()
        if dvec is None:
            dvec = torch.empty((B, H, qkv.shape[0] // B), dtype=torch.float32, device=qkv.device)
        _lib.check(L.rgbnm_attention_bwd(do.data_ptr(), qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), dqkv.data_ptr(),
                                         dvec.data_ptr(), B, qkv.shape[0] // B, H, D, C.c_float(scale), _lib.stream_ptr()),
                   "rgbnm_attention_bwd")
        return
    with torch.enable_grad():
        leaf = qkv.detach().requires_grad_(True)
        q, k, v = _split(leaf, B, H, D)
        out = F.scaled_dot_product_attention(q, k, v, scale=scale).transpose(1, 2).reshape(do.shape)
        (g,) = torch.autograd.grad(out, leaf, do)
    dqkv.copy_(g)
