"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not product code.

A CPU restatement (torch-CPU / numpy, the reference's own arithmetic library) of
the reference's DCT-domain data path, stage by stage, driven by a *resolved*
augmentation plan instead of live RNG draws.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module; the product path (rgb_no_more_b200/) never does.

Parity pinning: the reference has no tests or golden vectors of its own
(SURVEY.md 4, 8c).  This restatement is pinned against outputs of the *reference
itself*, generated in the build container by `tools/make_golden.py` (which imports
/root/reference) and committed under tests/golden/ -- see tests/test_oracle_golden.py.

Each function cites the reference file:line (relative to /root/reference) it follows.
"""
from __future__ import annotations

import math

import numpy as np
import torch

CLAMP_MIN = -2 ** 10
CLAMP_MAX = 2 ** 10 - 8

# ---------------------------------------------------------------------------
# Basis / conversion matrices            utils/dct_ops.py:150-208
# ---------------------------------------------------------------------------
_BASIS_CACHE = {}


def basis_matrix(length: int) -> torch.Tensor:
    """Orthonormal DCT-II basis D_L (dct_ops.py:150-169), fp32 op-for-op."""
    h = torch.arange(length, dtype=torch.float32).unsqueeze(1)
    w = torch.arange(length, dtype=torch.float32).unsqueeze(0)
    w = w + 0.5
    b = h.mm(w)
    b = b * torch.pi / length
    b = b.cos()
    b[0] *= 1 / (2 ** 0.5)
    b *= (2 / length) ** 0.5
    return b


def conversion_matrix(mult: int, small: int = 8) -> torch.Tensor:
    """A_mult = D_{small*mult} @ blockdiag(D_small x mult)^T (dct_ops.py:180-208)."""
    key = (mult, small)
    if key not in _BASIS_CACHE:
        if mult == 1:
            m = torch.eye(small, dtype=torch.float32)
        else:
            large = basis_matrix(small * mult)
            blocks = torch.block_diag(*basis_matrix(small).unsqueeze(0).repeat(mult, 1, 1))
            m = large.mm(blocks.T)
        _BASIS_CACHE[key] = m
    return _BASIS_CACHE[key]


# ---------------------------------------------------------------------------
# Dequantise                               datasets.py:286-293
# ---------------------------------------------------------------------------
def dequantize(y_q: torch.Tensor, c_q, quant: torch.Tensor):
    """y_q int16 (1,H,W,8,8); c_q int16 (2,H/2,W/2,8,8) or None; quant int16 (C,8,8)."""
    y = torch.clamp(y_q * quant[0], min=CLAMP_MIN, max=CLAMP_MAX)
    if c_q is not None:
        c = torch.clamp(c_q * quant[1:3].unsqueeze(1).unsqueeze(1), min=CLAMP_MIN, max=CLAMP_MAX)
    else:
        _, h, w, kh, kw = y.shape
        c = torch.zeros((2, h // 2, w // 2, kh, kw), dtype=y.dtype)
    return y, c


# ---------------------------------------------------------------------------
# Resize                                   utils/dct_ops.py:436-580
# ---------------------------------------------------------------------------
def _upsample(x: torch.Tensor, L: int) -> torch.Tensor:
    # dct_ops.py:457-482 (square factors, KH == KW == 8)
    if L == 1:
        return x.to(torch.float32)
    A = conversion_matrix(L)
    shape = list(x.shape)
    shape[-2] = L * 8
    shape[-1] = L * 8
    z = torch.zeros(tuple(shape), dtype=torch.float32)
    z[..., 0:8, 0:8] = x.to(torch.float32) * (L * L) ** 0.5
    z = torch.einsum("l o, c h w o m -> c h w l m", A.T, z)
    z = torch.einsum("c h w l o, o m -> c h w l m", z, A)
    c, h, w = z.shape[:3]
    z = z.reshape(c, h, w, L, 8, L, 8).permute(0, 1, 3, 2, 5, 4, 6)   # c h l w m kh kw
    return z.reshape(c, h * L, w * L, 8, 8)


def _downsample(x: torch.Tensor, L: int) -> torch.Tensor:
    # dct_ops.py:505-527
    if L == 1:
        return x.to(torch.float32)
    A = conversion_matrix(L)
    c, H, W = x.shape[:3]
    z = x.to(torch.float32).reshape(c, H // L, L, W // L, L, 8, 8).permute(0, 1, 3, 2, 5, 4, 6)
    z = z.reshape(c, H // L, W // L, L * 8, L * 8)
    z = torch.einsum("l o, c h w o m -> c h w l m", A, z)
    z = torch.einsum("c h w l o, o m -> c h w l m", z, A.T)
    return z[..., 0:8, 0:8] / (L * L) ** 0.5


def resize_blocks(x: torch.Tensor, size: int) -> torch.Tensor:
    """resize_dct for square block grids (dct_ops.py:541-578): upsample by
    size/gcd, downsample by H/gcd, round-half-even back to the input dtype."""
    _, H, W, _, _ = x.shape
    assert H == W
    g = math.gcd(H, size)
    up, down = size // g, H // g
    z = _upsample(x, up)
    z = _downsample(z, down)
    return torch.round(z).to(x.dtype)


def crop_blocks(x: torch.Tensor, top: int, left: int, height: int, width: int) -> torch.Tensor:
    """crop_dct in-bounds branch (dct_ops.py:584-599)."""
    return x[:, top:top + height, left:left + width]


# ---------------------------------------------------------------------------
# Geometric ops (integer exact)
# ---------------------------------------------------------------------------
def flip_blocks(x: torch.Tensor, direction: str = "horizontal", fixed_pos: bool = False) -> torch.Tensor:
    # dct_ops.py:601-621
    x = x.clone()
    if direction == "horizontal":
        if not fixed_pos:
            x = x.flip(dims=(2,))
        x[:, :, :, :, 1::2] *= -1
    else:
        if not fixed_pos:
            x = x.flip(dims=(1,))
        x[:, :, :, 1::2, :] *= -1
    return x


def rot90_blocks(x: torch.Tensor, direction: int) -> torch.Tensor:
    # dct_ops.py:99-130 restricted to rotate = +-1
    x = x.clone()
    if direction > 0:     # counter-clockwise
        x = torch.rot90(x, k=1, dims=(1, 2)).transpose(-2, -1)
        return flip_blocks(x, "vertical", fixed_pos=True)
    x = torch.rot90(x, k=-1, dims=(1, 2)).transpose(-2, -1)
    return flip_blocks(x, "horizontal", fixed_pos=True)


def translate_blocks(x: torch.Tensor, t: int, axis: str) -> torch.Tensor:
    # dct_ops.py:748-774
    x = x.clone()
    dim = 1 if axis == "H" else 2
    x = torch.roll(x, t, dims=(dim,))
    idx = [slice(None)] * 5
    idx[dim] = slice(None, t) if t >= 0 else slice(t, None)
    x[tuple(idx)] = 0
    return x


def zero_rect(x: torch.Tensor, r0: int, r1: int, c0: int, c1: int) -> torch.Tensor:
    # effect of cutout_dct's mask (dct_ops.py:801-814) given its resolved rectangle
    x = x.clone()
    x[:, r0:r1, c0:c1] = 0
    return x


# ---------------------------------------------------------------------------
# DC / per-coefficient ops
# ---------------------------------------------------------------------------
def _put_dc(x: torch.Tensor, dc: torch.Tensor) -> torch.Tensor:
    x = x.clone()
    x[:, :, :, 0, 0] = torch.round(dc).to(x.dtype)
    return x


def brightness(x: torch.Tensor, m_f32: float) -> torch.Tensor:
    # dct_ops.py:829-836; factor - 1 == m
    dc = x[:, :, :, 0, 0].to(torch.float32)
    dc = dc + torch.mean(torch.abs(dc)) * m_f32
    return _put_dc(x, dc)


def dc_scale(x: torch.Tensor, factor_f32: float) -> torch.Tensor:
    # contrast_dct, dct_ops.py:852-859
    dc = x[:, :, :, 0, 0].to(torch.float32)
    dc = dc * factor_f32
    return _put_dc(x, dc)


def autocontrast(x: torch.Tensor) -> torch.Tensor:
    # dct_ops.py:873-886.  The flat non-zero DC case divides by zero in the reference
    # (undefined int16 cast); the B200 path defines it as "leave unchanged".
    dc = x[:, :, :, 0, 0].to(torch.float32)
    lo, hi = dc.min(), dc.max()
    if lo == hi:
        return x.clone()
    dc = (dc - lo) / (hi - lo)
    dc = CLAMP_MIN + (dc * (CLAMP_MAX - CLAMP_MIN))
    return _put_dc(x, dc)


def posterize(x: torch.Tensor, bits: int) -> torch.Tensor:
    # dct_ops.py:900-913
    dc = x[:, :, :, 0, 0].to(torch.float32)
    dc = dc - CLAMP_MIN
    dc = dc / 2 ** bits
    idx = torch.round(dc).to(torch.int64)
    table = torch.linspace(CLAMP_MIN, CLAMP_MAX, round((CLAMP_MAX - CLAMP_MIN) / (2 ** bits)) + 1)
    return _put_dc(x, table[idx])


def solarize_add(x: torch.Tensor, addition: int) -> torch.Tensor:
    # dct_ops.py:668-679
    x = x.clone()
    dc = x[:, :, :, 0, 0]
    mask = dc < 0
    dc[mask] += addition
    x[:, :, :, 0, 0] = dc
    return torch.clamp(x, CLAMP_MIN, CLAMP_MAX)


def block_filter(x: torch.Tensor, filt: torch.Tensor) -> torch.Tensor:
    """sharpblur_dct / midfreqaug_dct given their 8x8 filter expressed in
    un-shifted block coordinates (dct_ops.py:696-708, 725-746)."""
    z = x * filt.reshape(1, 1, 1, 8, 8)
    z = z.clamp(min=CLAMP_MIN, max=CLAMP_MAX)
    return torch.round(z).to(x.dtype)


def solarize(y: torch.Tensor, c: torch.Tensor, threshold_f32: float):
    """solarize_dct on Y and, with the luma mask sub-sampled [::2, ::2], on CbCr (dct_ops.py:631-651,
    custom_transforms.py:980-982): every coefficient of a block is negated where the luma DC term exceeds the threshold."""
    y, c = y.clone(), c.clone()
    mask = y[:, :, :, 0, 0] > threshold_f32
    y[mask] *= -1
    c[mask[:, ::2, ::2].repeat(2, 1, 1)] *= -1
    return y, c


def equalize(x: torch.Tensor) -> torch.Tensor:
    """equalize_dct / scale_channel_dct (dct_ops.py:916-955), CPU branch (torch.bincount): histogram equalisation of the DC
    plane of every channel.  One distinct DC value divides by zero in the reference (NaN -> undefined int16 cast); the B200
    path defines that case as "leave unchanged"."""
    x = x.clone()
    for ch in range(x.shape[0]):
        dc = x[ch, :, :, 0, 0].clone() - CLAMP_MIN
        hist = torch.bincount(dc.reshape(-1).to(torch.int64), minlength=(CLAMP_MAX - CLAMP_MIN) + 1)
        nz = hist[hist != 0]
        if nz.numel() < 2:
            continue
        mn_minus_min = nz[1:].sum()
        cdf = torch.cumsum(hist, 0)
        eq = torch.round((cdf - nz[0]) / mn_minus_min * (CLAMP_MAX - CLAMP_MIN - 1))
        x[ch, :, :, 0, 0] = eq[dc.to(torch.int64)].to(x.dtype) + CLAMP_MIN
    return x


def freq_enhance(x: torch.Tensor, factor_f32: float) -> torch.Tensor:
    # freq_enhance_dct, dct_ops.py:1015-1034: every coefficient except DCT[0,0] times `magnitude` in fp32, rounded
    z = x.to(torch.float32).reshape(*x.shape[:3], 64).clone()
    z[:, :, :, 1:] *= factor_f32
    return torch.round(z).reshape(x.shape).to(x.dtype)


def to_range(x: torch.Tensor) -> torch.Tensor:
    # ToRange(-1, 1, -1024, 1016)  custom_transforms.py:448-452
    z = x.to(torch.float32)
    z = (z - (-1024)) / (1016 - (-1024))
    return -1 + (z * (1 - (-1)))


# ---------------------------------------------------------------------------
# Whole per-image transform driven by a resolved plan
# ---------------------------------------------------------------------------
def _clamp(x):
    return x.clamp(min=CLAMP_MIN, max=CLAMP_MAX)


def apply_op(y: torch.Tensor, c: torch.Tensor, op, filters: np.ndarray):
    """_apply_op_dct for one resolved op (custom_transforms.py:944-1021), including
    the trailing clamp on both planes."""
    name = op.name
    p = op.p
    if name == "TranslateX":
        y, c = translate_blocks(y, p[0], "W"), translate_blocks(c, p[1], "W")
    elif name == "TranslateY":
        y, c = translate_blocks(y, p[0], "H"), translate_blocks(c, p[1], "H")
    elif name == "Rotate90":
        y, c = rot90_blocks(y, p[0]), rot90_blocks(c, p[0])
    elif name == "Cutout":
        y, c = zero_rect(y, *p[0:4]), zero_rect(c, *p[4:8])
    elif name == "Brightness":
        y = brightness(y, op.f)
    elif name == "Contrast":
        y = dc_scale(y, op.f)
    elif name == "Color":
        c = dc_scale(c, op.f)
    elif name == "AutoContrast":
        y = autocontrast(y)
    elif name == "AutoSaturation":
        c = autocontrast(c)
    elif name == "Posterize":
        y, c = posterize(y, p[0]), posterize(c, p[0])
    elif name in ("Sharpness", "MidfreqAug"):
        y = block_filter(y, torch.from_numpy(np.asarray(filters[p[0]], dtype=np.float32)))
    elif name == "Grayscale":
        c = c * 0
    elif name == "ChromaDrop":
        c = c.clone()
        c[p[0]] *= 0
    elif name == "SolarizeAdd":
        y = solarize_add(y, p[0])
    elif name == "Invert":
        y, c = y * -1, c * -1
    elif name == "Solarize":
        y, c = solarize(y, c, op.f)
    elif name == "Equalize":
        y = equalize(y)
    elif name == "FreqEnhance":
        y, c = freq_enhance(y, op.f), freq_enhance(c, op.f)
    elif name == "Identity":
        pass
    else:
        raise ValueError(name)
    return _clamp(y).contiguous(), _clamp(c).contiguous()


def resized_planes(y_q, c_q, quant, plan, out_size: int = 28):
    """dequantise -> crop -> resize: the int16 planes RandomFlip_DCT receives."""
    y, c = dequantize(y_q, c_q, quant)
    s = plan.crop_size
    y = resize_blocks(crop_blocks(y, plan.crop_i, plan.crop_j, s, s), out_size)
    c = resize_blocks(crop_blocks(c, plan.crop_i // 2, plan.crop_j // 2, max(1, s // 2), max(1, s // 2)),
                      math.ceil(out_size / 2))
    return y, c


def resized_planes_exact(y_q, c_q, quant, plan, out_size: int = 28):
    """The same resize evaluated in float64 with float64 basis matrices and NOT rounded.
    Used by the parity tests to decide whether an int16 mismatch sits on an exact .5 tie
    of the real-valued result (where the reference's own answer depends on the BLAS
    summation order, SURVEY.md 7 hard part 2)."""
    def basis64(n):
        h = torch.arange(n, dtype=torch.float64).unsqueeze(1)
        w = torch.arange(n, dtype=torch.float64).unsqueeze(0) + 0.5
        b = (h @ w * math.pi / n).cos()
        b[0] *= 1 / math.sqrt(2)
        return b * math.sqrt(2 / n)

    def conv64(mult):
        return basis64(8 * mult) @ torch.block_diag(*[basis64(8)] * mult).T

    def resize64(x, size):
        _, H, _, _, _ = x.shape
        g = math.gcd(H, size)
        up, down = size // g, H // g
        z = x.to(torch.float64)
        if up > 1:
            A = conv64(up)
            c, h, w = z.shape[:3]
            big = torch.zeros((c, h, w, 8 * up, 8 * up), dtype=torch.float64)
            big[..., :8, :8] = z * up
            big = A.T @ big @ A
            z = big.reshape(c, h, w, up, 8, up, 8).permute(0, 1, 3, 2, 5, 4, 6).reshape(c, h * up, w * up, 8, 8)
        if down > 1:
            A = conv64(down)
            c, H2, W2 = z.shape[:3]
            z = z.reshape(c, H2 // down, down, W2 // down, down, 8, 8).permute(0, 1, 3, 2, 5, 4, 6)
            z = z.reshape(c, H2 // down, W2 // down, 8 * down, 8 * down)
            z = (A @ z @ A.T)[..., :8, :8] / down
        return z

    y, c = dequantize(y_q, c_q, quant)
    s = plan.crop_size
    return (resize64(crop_blocks(y, plan.crop_i, plan.crop_j, s, s), out_size),
            resize64(crop_blocks(c, plan.crop_i // 2, plan.crop_j // 2, max(1, s // 2), max(1, s // 2)),
                     math.ceil(out_size / 2)))


def transform_from_resized(y, c, plan, filters):
    """flip -> RandAugment ops on already-resized int16 planes (integer / DC arithmetic only)."""
    if plan.flip:
        y, c = flip_blocks(y), flip_blocks(c)
    if plan.train:
        y, c = _clamp(y), _clamp(c)                    # custom_transforms.py:1107-1108
        for op in plan.ops:
            y, c = apply_op(y, c, op, filters)
    return y, c


def transform_int16(y_q, c_q, quant, plan, filters, out_size: int = 28):
    """dequantise -> crop -> resize -> flip -> RandAugment ops; returns the int16
    planes the reference hands to ToRange: Y (1,S,S,8,8), CbCr (2,S/2,S/2,8,8)."""
    y, c = resized_planes(y_q, c_q, quant, plan, out_size)
    if plan.flip:
        y, c = flip_blocks(y), flip_blocks(c)
    if plan.train:
        y, c = _clamp(y), _clamp(c)                    # custom_transforms.py:1107-1108
        for op in plan.ops:
            y, c = apply_op(y, c, op, filters)
    return y, c


def embed_input(yf: torch.Tensor, cf: torch.Tensor, subblock: bool = True) -> torch.Tensor:
    """PatchEmbedding_DCT_Group.forward up to (not including) the Linear
    (models/plainvit.py:200-216) for patch 16: rearrange 2x2 luma blocks into a
    16x16 tile, A16 . X . A16^T, collapse to [Y 256 | Cb 64 | Cr 64].
    yf (B,1,28,28,8,8), cf (B,2,14,14,8,8) fp32 -> (B,14,14,384).
    subblock = False (`--no_subblock`: patch2subblock returns no matrix, plainvit.py:33-38; apply_subblock is the identity, :59-60):
    the rearrange 'b c (h pdh) (w pdw) p1 p2 -> b c h w (pdh p1) (pdw p2)' (:83) alone, i.e. the un-converted 16 x 16 tile."""
    b, _, H, W, _, _ = yf.shape
    if not subblock:
        y = yf.reshape(b, 1, H // 2, 2, W // 2, 2, 8, 8).permute(0, 1, 2, 4, 3, 6, 5, 7).reshape(b, H // 2, W // 2, 256)
        c = cf.permute(0, 2, 3, 1, 4, 5).reshape(b, H // 2, W // 2, 128)
        return torch.cat([y, c], dim=3)
    A = conversion_matrix(2)
    y = yf.reshape(b, 1, H // 2, 2, W // 2, 2, 8, 8).permute(0, 1, 2, 4, 3, 6, 5, 7)
    y = y.reshape(b, 1, H // 2, W // 2, 16, 16)
    y = torch.einsum("i o, b c h w o j -> b c h w i j", A, y)
    y = torch.einsum("b c h w i o, o j -> b c h w i j", y, A.T)
    y = y.permute(0, 2, 3, 1, 4, 5).reshape(b, H // 2, W // 2, 256)
    c = cf.permute(0, 2, 3, 1, 4, 5).reshape(b, H // 2, W // 2, 128)
    return torch.cat([y, c], dim=3)


def transform_embed(y_q, c_q, quant, plan, filters) -> torch.Tensor:
    """Full K0 semantics for one image: (196, 384) fp32."""
    y, c = transform_int16(y_q, c_q, quant, plan, filters)
    out = embed_input(to_range(y).unsqueeze(0), to_range(c).unsqueeze(0))
    return out.reshape(196, 384)


# ---------------------------------------------------------------------------
# SwinV2 data path: 32-block geometry, patch 4
# ---------------------------------------------------------------------------
def embed_input_swin(yf: torch.Tensor, cf: torch.Tensor) -> torch.Tensor:
    """swinv2.PatchEmbedding_DCT_Group.forward up to (not including) the Linear(24, E)
    (models/swinv2.py:553-571 with patch 4: combine_Y = combine_C = False): every 8x8 block is DECOMPOSED,
    D = A^T . X . A (apply_subblock combine=False, plainvit.py:65-68) with A = A(4,2) for luma and A(2,4)
    for chroma (patch2subblock, plainvit.py:43-47), then the reference's rearrange
    'b c h w (p1 pdh) (p2 pdw) -> b c (h pdh) (w pdw) p1 p2' (plainvit.py:86) -- note the INTERLEAVED split
    of the 8 indices (i = p1*pd + pdh) -- collapse 'b c h w i j -> b h w (c i j)' and concat [Y 16 | Cb 4 | Cr 4].
    yf (B,1,32,32,8,8), cf (B,2,16,16,8,8) fp32 -> (B,64,64,24)."""
    def decompose(x, small, pd):
        A = conversion_matrix(pd, small)
        x = torch.einsum("i o, b c h w o j -> b c h w i j", A.T, x)
        x = torch.einsum("b c h w i o, o j -> b c h w i j", x, A)
        b, c, h, w, _, _ = x.shape
        x = x.reshape(b, c, h, w, small, pd, small, pd)            # (p1 pdh) (p2 pdw)
        x = x.permute(0, 1, 2, 5, 3, 7, 4, 6)                      # b c h pdh w pdw p1 p2
        return x.reshape(b, c, h * pd, w * pd, small, small)

    y = decompose(yf, 4, 2)
    c = decompose(cf, 2, 4)
    b, _, H, W, _, _ = y.shape
    y = y.permute(0, 2, 3, 1, 4, 5).reshape(b, H, W, 16)
    c = c.permute(0, 2, 3, 1, 4, 5).reshape(b, H, W, 8)
    return torch.cat([y, c], dim=3)


def transform_embed_swin(y_q, c_q, quant, plan, filters) -> torch.Tensor:
    """Full K0 semantics for one image in the SwinV2 layout: (4096, 24) fp32."""
    y, c = transform_int16(y_q, c_q, quant, plan, filters, out_size=32)
    out = embed_input_swin(to_range(y).unsqueeze(0), to_range(c).unsqueeze(0))
    return out.reshape(4096, 24)
