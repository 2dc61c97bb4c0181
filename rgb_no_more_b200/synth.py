"""Deterministic synthetic inputs (no dataset is available offline).

Images follow SURVEY.md 8(d): a low-resolution uniform-random RGB field, bicubically
up-sampled to 512x512, plus +-12 uniform pixel noise, saved by PIL as baseline JPEG,
quality 75, 4:2:0 -- the format the reference's own 512x512 resizer writes
(/root/reference/utils/mp_scripts.py:74-81).  Seed base = the reference's CFG.SEED
(/root/reference/utils/configs.py:7).
"""
from __future__ import annotations

import io
from typing import List

import numpy as np

SEED = 11997733


def synth_rgb(index: int, size: int = 512, seed: int = SEED) -> np.ndarray:
    """uint8 (size, size, 3) natural-ish image, deterministic in (seed, index)."""
    from PIL import Image
    rng = np.random.default_rng([seed, index])
    low = rng.integers(0, 256, size=(size // 16, size // 16, 3), dtype=np.uint8)
    img = Image.fromarray(low, "RGB").resize((size, size), Image.BICUBIC)
    arr = np.asarray(img).astype(np.int16)
    arr = arr + rng.integers(-12, 13, size=arr.shape, dtype=np.int16)
    return np.clip(arr, 0, 255).astype(np.uint8)


def synth_jpeg(index: int, size: int = 512, quality: int = 75, subsampling: str = "4:2:0",
               seed: int = SEED, restart_blocks: int = 0) -> bytes:
    """Baseline-JPEG bytes of synth_rgb(index)."""
    from PIL import Image
    buf = io.BytesIO()
    kw = {}
    if restart_blocks:
        kw["restart_marker_blocks"] = restart_blocks
    Image.fromarray(synth_rgb(index, size, seed), "RGB").save(
        buf, format="JPEG", quality=quality, subsampling=subsampling, optimize=False, progressive=False, **kw)
    return buf.getvalue()


def synth_jpeg_set(n: int, size: int = 512) -> List[bytes]:
    return [synth_jpeg(i, size) for i in range(n)]


def synth_coefficients(batch: int, hb: int = 64, wb: int = 64, seed: int = SEED, dense: bool = False):
    """Synthetic *quantised* coefficient planes with JPEG-like statistics, without
    running an encoder: Laplacian magnitudes decaying with frequency, Q75-like tables.
    Returns (y int16 [B,hb,wb,64], c int16 [B,2,hb/2,wb/2,64], q int16 [B,3,64]).
    `dense=True` gives the adversarial uniform set used for parity only."""
    rng = np.random.default_rng([seed, batch, hb, wb, int(dense)])
    u = np.arange(8)
    freq = (u[:, None] + u[None, :]).reshape(64).astype(np.float32)
    q_luma = np.clip(np.round(4 + 3.2 * freq), 1, 255).astype(np.int16)
    q_chroma = np.clip(np.round(6 + 7.0 * freq), 1, 255).astype(np.int16)
    q = np.stack([q_luma, q_chroma, q_chroma])[None].repeat(batch, 0)
    if dense:
        y = rng.integers(-1024, 1017, size=(batch, hb, wb, 64)).astype(np.int16)
        c = rng.integers(-1024, 1017, size=(batch, 2, hb // 2, wb // 2, 64)).astype(np.int16)
        q[:] = 1
        q[:, :, 5] = 3          # exercise the dequant clamp on one coefficient
        return y, c, q

    def plane(shape, qtab, dc_sigma):
        scale = (220.0 / (1.0 + freq) ** 1.6) / qtab
        x = rng.laplace(0.0, 1.0, size=shape + (64,)).astype(np.float32) * scale
        x[..., 0] = rng.normal(0.0, dc_sigma, size=shape) / qtab[0]
        return np.round(x).astype(np.int16)

    y = plane((batch, hb, wb), q_luma.astype(np.float32), 420.0)
    c = plane((batch, 2, hb // 2, wb // 2), q_chroma.astype(np.float32), 200.0)
    y = np.clip(y.astype(np.int32) * q_luma, -1024, 1016) // q_luma
    c = np.clip(c.astype(np.int32) * q_chroma, -1024, 1016) // q_chroma
    return y.astype(np.int16), c.astype(np.int16), q


def dequant_clamp_flags(y: np.ndarray, c: np.ndarray, q: np.ndarray) -> np.ndarray:
    """Per image: 1 iff some dequantised coefficient leaves [-1024, 1016], i.e. the clamp of datasets.py:288-290 is
    live (what rgbnm_jpeg_decode_batch reports for decoded files; here for generated coefficients)."""
    n = y.shape[0]
    yq = y.reshape(n, -1, 64).astype(np.int32) * q[:, 0:1, :].astype(np.int32)
    cq = c.reshape(n, 2, -1, 64).astype(np.int32) * q[:, 1:3, None, :].astype(np.int32)
    bad = (yq.min(axis=(1, 2)) < -1024) | (yq.max(axis=(1, 2)) > 1016) | (cq.min(axis=(1, 2, 3)) < -1024) | (cq.max(axis=(1, 2, 3)) > 1016)
    return bad.astype(np.uint8)
