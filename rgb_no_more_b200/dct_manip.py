"""Drop-in for the reference's `dct_manip` extension module (boundary B1).

`read_coefficients(path)` has the reference's return contract
(/root/reference/dct_manip/dct_manip.cpp:152-178, bound at :578-606):
    (dimensions int32 (C,2), quantization int16 (C,8,8), Y int16 (1,Hb,Wb,8,8),
     CbCr int16 (2,Hb/2,Wb/2,8,8) or None)
Errors surface as RuntimeError, like the pybind11-translated C++ exceptions of the
reference (file open failure :155-159, libjpeg error_exit :24-41).

To use it from the reference's datasets.py (`import dct_manip as dm`, datasets.py:10) put
a one-line shim module named `dct_manip` on sys.path -- see INTEGRATION.md.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import lib as _lib


def _info(buf: bytes) -> _lib.JpegInfo:
    L = _lib.load()
    info = _lib.JpegInfo()
    rc = L.rgbnm_jpeg_info_from_memory(buf, len(buf), C.byref(info))
    if rc != 0:
        raise RuntimeError(L.rgbnm_strerror(rc).decode())
    return info


def read_coefficients_from_bytes(buf: bytes, return_clamp_flag: bool = False):
    L = _lib.load()
    info = _info(buf)
    hb, wb = info.hb[0], info.wb[0]
    y = torch.empty((1, hb, wb, 8, 8), dtype=torch.int16)
    quant = torch.empty((info.ncomp, 8, 8), dtype=torch.int16)
    dims = torch.empty((info.ncomp, 2), dtype=torch.int32)
    cbcr = None
    c_ptr, c_cap = None, 0
    if info.ncomp > 1:
        cbcr = torch.empty((2, info.hb[1], info.wb[1], 8, 8), dtype=torch.int16)
        c_ptr, c_cap = cbcr.data_ptr(), cbcr.numel()
    flag = C.c_int32(1)
    rc = L.rgbnm_jpeg_read_coefficients(buf, len(buf), y.data_ptr(), y.numel(), c_ptr, c_cap,
                                        quant.data_ptr(), dims.data_ptr(), C.addressof(flag))
    if rc != 0:
        raise RuntimeError(L.rgbnm_strerror(rc).decode())
    if return_clamp_flag:
        return dims, quant, y, cbcr, bool(flag.value)
    return dims, quant, y, cbcr


def read_coefficients(path: str):
    """dct_manip.read_coefficients(path) -> (dimensions, quantization, Y, CbCr|None)."""
    try:
        with open(path, "rb") as f:
            buf = f.read()
    except OSError:
        raise RuntimeError(f"Unable to open file for reading: {path}")
    return read_coefficients_from_bytes(buf)


def decode_batch(jpegs: Sequence[bytes], hb: int = 64, wb: int = 64, nthreads: int = 0, pin: bool = False, out=None, last_rows=None):
    """Multithreaded batch decode into the layout the fused kernel reads.
    Returns (y [n,hb,wb,64], cbcr [n,2,hb/2,wb/2,64], quant [n,3,64], clamp_flags uint8 [n]).
    `out` = (y, cbcr, quant) of a previous call (e.g. one slot of a pinned staging ring) is written in place.
    `last_rows` (plan-first decoding): per image the last luma block row its crop window needs (crop_i + crop_size - 1); the scan
    is abandoned after the MCU row that holds it and the block rows below keep the buffers' previous contents."""
    L = _lib.load()
    n = len(jpegs)
    pin = pin and torch.cuda.is_available()
    if out is not None:
        y, c, q = out
        if (y.shape != (n, hb, wb, 64) or c.shape != (n, 2, hb // 2, wb // 2, 64) or q.shape != (n, 3, 64)
                or any(t.dtype != torch.int16 or not t.is_contiguous() or t.is_cuda for t in (y, c, q))):
            raise ValueError("rgbnm decode_batch: `out` must be the (y, cbcr, quant) host tensors of a same-sized batch")
    else:
        y = torch.empty((n, hb, wb, 64), dtype=torch.int16, pin_memory=pin)
        c = torch.empty((n, 2, hb // 2, wb // 2, 64), dtype=torch.int16, pin_memory=pin)
        q = torch.empty((n, 3, 64), dtype=torch.int16, pin_memory=pin)
    flags = torch.empty((n,), dtype=torch.uint8)
    status = torch.zeros((n,), dtype=torch.int32)
    ptrs = (C.c_char_p * n)(*jpegs)
    sizes = (C.c_size_t * n)(*[len(j) for j in jpegs])
    rows = None
    if last_rows is not None:
        rows = np.ascontiguousarray(np.asarray(last_rows, dtype=np.int32))
        if rows.shape != (n,):
            raise ValueError("rgbnm decode_batch: one last_rows entry per image")
    rc = L.rgbnm_jpeg_decode_batch_rows(C.cast(ptrs, C.c_void_p), C.cast(sizes, C.c_void_p), n, hb, wb, y.data_ptr(),
                                        c.data_ptr(), q.data_ptr(), flags.data_ptr(), status.data_ptr(), nthreads,
                                        None if rows is None else rows.ctypes.data)
    if rc != 0:
        bad = int(torch.nonzero(status)[0]) if status.any() else -1
        raise RuntimeError(f"image {bad}: {L.rgbnm_strerror(rc).decode()}")
    return y, c, q, flags


def write_coefficients(width: int, height: int, y: torch.Tensor, cbcr: Optional[torch.Tensor],
                       quant: torch.Tensor, chroma: Tuple[int, int] = (2, 2)) -> bytes:
    """Baseline JPEG bytes from quantised coefficient planes (fixture writer)."""
    L = _lib.load()
    out = C.c_void_p()
    size = C.c_size_t()
    ncomp = 1 if cbcr is None else 3
    y = y.contiguous()
    quant = quant.contiguous()
    cb = cbcr.contiguous() if cbcr is not None else None
    rc = L.rgbnm_jpeg_write_coefficients(width, height, ncomp, chroma[0], chroma[1], y.data_ptr(),
                                         cb.data_ptr() if cb is not None else None, quant.data_ptr(),
                                         C.byref(out), C.byref(size))
    if rc != 0:
        raise RuntimeError(L.rgbnm_strerror(rc).decode())
    data = C.string_at(out, size.value)
    L.rgbnm_free(out)
    return data


# ---- B1 "for completeness": coefficient <-> pixel conversions (dct_manip.cpp:315-375, 485-576) -----------------------------------
# The reference implements both by a round trip through libjpeg's own compressor / decompressor; neither is on the training path
# (their only callers, custom_transforms.ycbcr_to_rgb / rgb_to_dct at :1140-1196, are not referenced by any transform recipe).
# Here the libjpeg half of the round trip is Pillow's bundled libjpeg(-turbo) -- the same library family, default settings
# (ISLOW DCT, fancy up-sampling, 4:2:0) -- and the coefficient half is this repository's reader / writer.
_ANNEX_K_LUMA = [16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
                 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112,
                 100, 103, 99]
_ANNEX_K_CHROMA = [17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99] + \
                  [99] * 32


def quality_tables(quality: int) -> torch.Tensor:
    """jpeg_set_quality(quality, force_baseline=TRUE): the Annex K tables scaled libjpeg's way -> int16 (2, 8, 8), natural order."""
    q = min(max(int(quality), 1), 100)
    scale = 5000 // q if q < 50 else 200 - 2 * q
    tabs = [[min(max((v * scale + 50) // 100, 1), 255) for v in base] for base in (_ANNEX_K_LUMA, _ANNEX_K_CHROMA)]
    return torch.tensor(tabs, dtype=torch.int16).view(2, 8, 8)


def decode_coeff(dimensions: torch.Tensor, quantization: torch.Tensor, Y_coefficients: torch.Tensor,
                 CrCb_coefficients: Optional[torch.Tensor] = None, quality: int = -1) -> torch.Tensor:
    """Quantised coefficients -> pixels, uint8 (C, H, W) (RGB, or one channel for grayscale): dct_manip.cpp:485-576.
    `quality` > 0 replaces `quantization` by the standard tables of that quality, as the reference does."""
    import io
    from PIL import Image
    h, w = int(dimensions[0][0]), int(dimensions[0][1])
    qt = quality_tables(quality) if quality > 0 else quantization.to(torch.int16).reshape(-1, 8, 8)
    y = Y_coefficients.to(torch.int16).reshape(-1, 64)
    if CrCb_coefficients is None:
        buf = write_coefficients(w, h, y, None, qt[:1].reshape(1, 64))
    else:
        q3 = torch.stack([qt[0], qt[1], qt[1]]).reshape(3, 64)         # set_quantization: table 0 = luma, table 1 = both chroma planes
        buf = write_coefficients(w, h, y, CrCb_coefficients.to(torch.int16).reshape(-1, 64), q3, chroma=(2, 2))
    im = Image.open(io.BytesIO(buf))
    arr = np.asarray(im.convert("RGB") if CrCb_coefficients is not None else im.convert("L"))
    t = torch.from_numpy(arr.copy())
    return t.permute(2, 0, 1).contiguous() if t.dim() == 3 else t.unsqueeze(0)


def quantize_at_quality(pixels: torch.Tensor, quality: int, baseline: bool = True):
    """uint8 pixels (C, H, W) -> (dimensions, quantization, Y, CbCr | None) of the JPEG libjpeg writes at `quality` with its default
    settings (4:2:0 for colour): dct_manip.cpp:315-375."""
    import io
    from PIL import Image
    if pixels.dtype != torch.uint8 or pixels.dim() != 3 or pixels.shape[0] not in (1, 3):
        raise RuntimeError("quantize_at_quality: uint8 pixels of shape (1 | 3, H, W) expected")
    if not baseline:
        raise RuntimeError("quantize_at_quality: only baseline tables (8-bit) are supported")
    arr = pixels.permute(1, 2, 0).contiguous().numpy()
    im = Image.fromarray(arr[..., 0], "L") if pixels.shape[0] == 1 else Image.fromarray(arr, "RGB")
    out = io.BytesIO()
    im.save(out, "JPEG", quality=int(quality), subsampling=2 if pixels.shape[0] == 3 else -1, optimize=False)
    return read_coefficients_from_bytes(out.getvalue())
