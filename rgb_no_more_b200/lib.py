"""ctypes binding of librgbnm_b200.so (the C-ABI declared in include/rgbnm_b200.h).

The product path has no CPU fallback: if the shared library is missing, or a CUDA
entry point is called without a GPU, the call raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RGBNM_LIB") or os.path.join(_HERE, "librgbnm_b200.so")     # RGBNM_LIB: A/B builds of the library
_lib = None


class RgbnmError(RuntimeError):
    pass


class JpegInfo(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("ncomp", C.c_int32), ("progressive", C.c_int32),
                ("hb", C.c_int32 * 3), ("wb", C.c_int32 * 3), ("dsh", C.c_int32 * 3), ("dsw", C.c_int32 * 3),
                ("hsamp", C.c_int32 * 3), ("vsamp", C.c_int32 * 3)]


class WPrepDesc(C.Structure):
    _fields_ = [("w", C.c_void_p), ("w_bf16", C.c_void_p), ("wt_bf16", C.c_void_p), ("bias", C.c_void_p), ("bias_k", C.c_void_p),
                ("n", C.c_int32), ("k", C.c_int32), ("qkv_heads", C.c_int32), ("head_dim", C.c_int32), ("first_tile", C.c_int32),
                ("pad", C.c_int32)]


class K0Tables(C.Structure):
    _fields_ = [("filters", C.c_void_p), ("posterize_lut", C.c_void_p), ("equalize_lut", C.c_void_p)]


# name -> (restype, argtypes); must list every symbol include/rgbnm_b200.h declares.
_vp, _i, _sz = C.c_void_p, C.c_int, C.c_size_t
SIGNATURES = {
    "rgbnm_strerror": (C.c_char_p, [_i]),
    "rgbnm_last_cuda_error": (C.c_char_p, []),
    "rgbnm_abi_version": (_i, []),
    "rgbnm_jpeg_info_from_memory": (_i, [_vp, _sz, C.POINTER(JpegInfo)]),
    "rgbnm_jpeg_read_coefficients": (_i, [_vp, _sz, _vp, _sz, _vp, _sz, _vp, _vp, _vp]),
    "rgbnm_jpeg_decode_batch": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i]),
    "rgbnm_jpeg_decode_batch_rows": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "rgbnm_jpeg_read_file": (_i, [C.c_char_p, C.POINTER(_vp), C.POINTER(_sz)]),
    "rgbnm_free": (None, [_vp]),
    "rgbnm_jpeg_write_coefficients": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, C.POINTER(_vp), C.POINTER(_sz)]),
    "rgbnm_k0_dcstats": (_i, [_vp, _vp, _vp, _vp, C.POINTER(K0Tables), _vp, _i, _i, _i, _vp]),
    "rgbnm_k0_fused": (_i, [_vp, _vp, _vp, _vp, C.POINTER(K0Tables), _vp, _vp, _i, _i, _i, _i, _vp]),
    "rgbnm_k0_launch_count": (_i, []),
    "rgbnm_k0_dcstats_ex": (_i, [_vp, _vp, _vp, _vp, C.POINTER(K0Tables), _vp, _i, _i, _i, _i, _vp]),
    "rgbnm_k0_fused_ex": (_i, [_vp, _vp, _vp, _vp, C.POINTER(K0Tables), _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "rgbnm_gemm_bf16": (_i, [_vp, _vp]),
    "rgbnm_attention_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, C.c_float, _vp]),
    "rgbnm_attention_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, C.c_float, _vp]),
    "rgbnm_layernorm_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, C.c_float, _vp]),
    "rgbnm_layernorm_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "rgbnm_layernorm_bwd_ex": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "rgbnm_mixup_bf16": (_i, [_vp, _vp, _vp, _i, C.c_longlong, _vp]),
    "rgbnm_colsum_bf16": (_i, [_vp, C.c_longlong, _i, _i, _vp, _i, _i, _vp]),
    "rgbnm_weight_prep": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "rgbnm_weight_prep_batch": (_i, [_vp, _i, _i, _vp]),
    "rgbnm_qkv_perm_vec": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "rgbnm_qkv_unperm_rows_add": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "rgbnm_sumsq_f32": (_i, [_vp, C.c_longlong, _vp, _vp]),
    "rgbnm_layernorm_res_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, C.c_float, _vp]),
    "rgbnm_window_attention_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "rgbnm_patch_merge_gather": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "rgbnm_token_mean_bf16": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "rgbnm_layernorm_res_scaled_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _i, C.c_float, _vp]),
    "rgbnm_layernorm_res_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, C.c_float, _vp]),
    "rgbnm_layernorm_res_bwd_ex": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, C.c_float, _vp]),
    "rgbnm_window_attention_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "rgbnm_patch_merge_scatter": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "rgbnm_adamw_step": (_i, [_vp, _vp, _vp, _vp, C.c_longlong, C.c_longlong, _vp, _vp, _vp]),
}


def load():
    """Load the shared library (once) and type every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RgbnmError(f"{LIB_PATH} is missing: run `python __graft_entry__.py` (build()) first; "
                         "there is no CPU fallback for the B200 path")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        lib = load()
        msg = lib.rgbnm_strerror(rc).decode()
        if rc == 7:
            msg += ": " + lib.rgbnm_last_cuda_error().decode()
        raise RgbnmError(f"{what}: {msg}" if what else msg)


def stream_ptr():
    """Raw cudaStream_t of torch's current stream."""
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def no_autocast(fn):
    """The engines compute in their own precision (bf16 operands, fp32 accumulation and statistics): the torch ops inside them
    (heads, attention tables, losses) must not be re-cast by a caller's `torch.autocast` region (the reference enables AMP for
    vitb / swinv2, utils/configs.py:110, 137)."""
    import functools
    import torch

    @functools.wraps(fn)
    def wrapped(*a, **kw):
        with torch.autocast(device_type="cuda", enabled=False):
            return fn(*a, **kw)
    return wrapped
