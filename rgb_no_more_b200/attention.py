"""Attention core of the DCT ViT: O = softmax(Q K^T / sqrt(emb_size)) V per (image, head)
(/root/reference/models/plainvit.py:450-461 -- note the reference scales by sqrt(emb_size), not sqrt(head_dim)).

Input is the fused projection output in the kernel-side layout qkv[B*N, 3*H*D] = [q | k | v], each H*D wide and
head-major (the "(h d qkv)" interleave of plainvit.py:447 is undone by regrouping the weight rows once per
step, rgbnm_weight_prep).  Output o[B*N, H*D] is already 'b n (h d)' (plainvit.py:461).

One implementation: the hand-written tcgen05 kernels (csrc/attention_tc.cu) through the C-ABI.  There is no library
(SDPA) path in the product; the torch formulation the kernels are checked against lives in tests/test_vit_gpu.py.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import lib as _lib


def forward(qkv, o, lse, B, H, D, scale, backend="b200"):
    if backend not in ("b200", "auto"):
        raise ValueError("rgbnm attention: the only backend is 'b200' (tcgen05 kernels); library fallbacks were removed")
    L = _lib.load()
    _lib.check(L.rgbnm_attention_fwd(qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), B, qkv.shape[0] // B, H, D,
                                     C.c_float(scale), _lib.stream_ptr()), "rgbnm_attention_fwd")


def backward(do, qkv, o, lse, dqkv, B, H, D, scale, backend="b200", dvec=None):
    if backend not in ("b200", "auto"):
        raise ValueError("rgbnm attention: the only backend is 'b200' (tcgen05 kernels); library fallbacks were removed")
    L = _lib.load()
    if dvec is None:
        dvec = torch.empty((B, H, qkv.shape[0] // B), dtype=torch.float32, device=qkv.device)
    _lib.check(L.rgbnm_attention_bwd(do.data_ptr(), qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), dqkv.data_ptr(),
                                     dvec.data_ptr(), B, qkv.shape[0] // B, H, D, C.c_float(scale), _lib.stream_ptr()),
               "rgbnm_attention_bwd")
