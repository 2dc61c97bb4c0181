"""Host binding of the tcgen05 GEMM entry point (include/rgbnm_b200.h: rgbnm_gemm_bf16).

C[M,N] = sum_k A[m,k] * B[n,k] with bf16 operands, fp32 accumulation in TMEM, and the fused
epilogues the DCT ViT needs (models/plainvit.py:194-198, 441-443, 475-491).  No CPU fallback."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import lib as _lib

EPI_STORE, EPI_RESIDUAL, EPI_GELU, EPI_DGELU, EPI_POSEMB, EPI_WGRAD_ATOMIC, EPI_F32, EPI_GELU_ACT, EPI_LNRES, EPI_LN = range(10)


class GemmArgs(C.Structure):
    _fields_ = [("A", C.c_void_p), ("B", C.c_void_p), ("C", C.c_void_p), ("C2", C.c_void_p), ("aux", C.c_void_p),
                ("bias", C.c_void_p), ("posemb", C.c_void_p), ("out_f32", C.c_void_p),
                ("lda", C.c_longlong), ("ldb", C.c_longlong), ("ldc", C.c_longlong), ("ldaux", C.c_longlong),
                ("ldo", C.c_longlong), ("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("epilogue", C.c_int),
                ("pos_period", C.c_int), ("splits", C.c_int), ("alpha", C.c_float), ("trans_out", C.c_int),
                ("perm_heads", C.c_int), ("perm_head_dim", C.c_int),
                ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p), ("ln_eps", C.c_float)]


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _check_bf16(t, name):
    if t.dtype != torch.bfloat16 or not t.is_cuda or t.stride(-1) != 1:
        raise ValueError(f"rgbnm gemm: {name} must be a CUDA bf16 tensor with unit inner stride")


def gemm(a: torch.Tensor, b: torch.Tensor, epilogue: int = EPI_STORE, bias: Optional[torch.Tensor] = None,
         aux: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, out2: Optional[torch.Tensor] = None,
         posemb: Optional[torch.Tensor] = None, out_f32: Optional[torch.Tensor] = None, splits: int = 0,
         alpha: float = 1.0, trans_out: bool = False, perm_heads: int = 0, perm_head_dim: int = 0,
         ln: Optional[tuple] = None, ln_eps: float = 1e-5):
    """Forward / dgrad form: a [M,K], b [N,K] (both row-major, K contiguous) -> out [M,N].
    For EPI_WGRAD_ATOMIC: a [T,M] and b [T,N] (T = reduction/token index) -> out_f32 [M,N] += alpha * a^T b
    (trans_out: out_f32 [N,M] += alpha * b^T a, so the caller can put the longer side on the 256-row tile axis);
    splits = 0 lets the library pick the split count over T."""
    L = _lib.load()
    _check_bf16(a, "a")
    _check_bf16(b, "b")
    args = GemmArgs()
    if epilogue == EPI_WGRAD_ATOMIC:
        K, M = a.shape
        K2, N = b.shape
    else:
        M, K = a.shape
        N, K2 = b.shape
    if K != K2:
        raise ValueError("rgbnm gemm: reduction lengths differ")
    args.A, args.B = a.data_ptr(), b.data_ptr()
    args.lda, args.ldb = a.stride(0), b.stride(0)
    args.M, args.N, args.K = M, N, K
    args.epilogue = epilogue
    args.splits = splits
    args.alpha = alpha
    args.trans_out = 1 if trans_out else 0
    args.perm_heads, args.perm_head_dim = perm_heads, perm_head_dim     # WGRAD_ATOMIC: M index kernel-qkv order -> reference rows
    ret = None
    if epilogue in (EPI_WGRAD_ATOMIC, EPI_F32):
        if out_f32 is None:
            out_f32 = torch.zeros((N, M) if trans_out else (M, N), dtype=torch.float32, device=a.device)
        if out_f32.dtype != torch.float32 or out_f32.stride(-1) != 1:
            raise ValueError("rgbnm gemm: out_f32 must be fp32 with unit inner stride")
        args.out_f32, args.ldo = out_f32.data_ptr(), out_f32.stride(0)
        ret = out_f32
    else:
        if out is None:
            out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
        _check_bf16(out, "out")
        args.C, args.ldc = out.data_ptr(), out.stride(0)
        ret = out
        if epilogue == EPI_GELU:
            if out2 is None:
                out2 = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
            _check_bf16(out2, "out2")
            if out2.stride(0) != out.stride(0):
                raise ValueError("rgbnm gemm: out and out2 must share the leading dimension")
            args.C2 = out2.data_ptr()
            ret = (out, out2)
        if epilogue in (EPI_RESIDUAL, EPI_DGELU, EPI_LNRES):
            _check_bf16(aux, "aux")
            args.aux, args.ldaux = aux.data_ptr(), aux.stride(0)
        if epilogue in (EPI_LNRES, EPI_LN):       # out = (aux +) LayerNorm(a b^T + bias) * gamma + beta; ln = (gamma, beta) fp32 [N]
            gamma, beta = ln
            if gamma.dtype != torch.float32 or beta.dtype != torch.float32 or gamma.numel() != N or beta.numel() != N:
                raise ValueError("rgbnm gemm: ln = (gamma, beta) must be fp32 [N]")
            args.ln_gamma, args.ln_beta, args.ln_eps = gamma.data_ptr(), beta.data_ptr(), ln_eps
    if bias is not None:
        if bias.dtype != torch.float32 or bias.numel() != N:
            raise ValueError("rgbnm gemm: bias must be fp32 [N]")
        args.bias = bias.data_ptr()
    if epilogue == EPI_POSEMB:
        if posemb.dtype != torch.float32 or posemb.shape[1] != N or not posemb.is_contiguous():
            raise ValueError("rgbnm gemm: posemb must be contiguous fp32 [period, N]")
        args.posemb, args.pos_period = posemb.data_ptr(), posemb.shape[0]
    _lib.check(L.rgbnm_gemm_bf16(C.byref(args), _lib.stream_ptr()), "rgbnm_gemm_bf16")
    return ret
