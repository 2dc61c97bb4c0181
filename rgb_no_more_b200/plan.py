"""Host-side augmentation-plan sampler for the fused DCT kernel (K0).

The reference draws every random augmentation parameter on the CPU, inside the
transform modules, from the torch *global* CPU generator.  The B200 path keeps
exactly that contract: this module replays the reference's RNG call sequence
(SURVEY.md section 8a "RNG contract") and resolves every draw into a fixed-size
per-image POD record -- the *plan*.  The CUDA kernel is then a pure function of
(quantised coefficients, quantisation tables, plan).

Reference call sites mirrored here (file:line in /root/reference):
  * RandomResizedCrop_DCT.get_params          utils/custom_transforms.py:557-629
  * ResizedCenterCrop_DCT.get_params          utils/custom_transforms.py:850-882
  * RandomFlip_DCT.forward                    utils/custom_transforms.py:926-942
  * RandAugment_dct.forward / _augmentation_space   :1066-1127
  * _apply_op_dct (parameter resolution only) :944-1021
  * cutout_dct centre draws                   utils/dct_ops.py:791-799
  * filters of sharpblur_dct / midfreqaug_dct utils/dct_ops.py:696-698,728-737
  * posterize_dct quantisation table          utils/dct_ops.py:903-912

Nothing here touches the oracle; the module is product code and is covered by
CPU tests against golden plans dumped from the reference itself.
"""
from __future__ import annotations

import functools
import itertools
import math
import struct
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

# --------------------------------------------------------------------------
# Op codes shared with include/rgbnm_b200.h (enum rgbnm_op) -- keep in sync.
# --------------------------------------------------------------------------
OP_NOP = 0
OP_TRANSLATE_X = 1
OP_TRANSLATE_Y = 2
OP_ROT90 = 3
OP_CUTOUT = 4
OP_BRIGHTNESS = 5
OP_CONTRAST = 6
OP_COLOR = 7
OP_AUTOCONTRAST = 8
OP_AUTOSATURATION = 9
OP_POSTERIZE = 10
OP_SHARPNESS = 11
OP_MIDFREQ = 12
OP_GRAYSCALE = 13
OP_CHROMADROP = 14
OP_SOLARIZE_ADD = 15
OP_INVERT = 16
OP_FREQ_ENHANCE = 17
OP_EQUALIZE = 18
OP_SOLARIZE = 19

OP_NAMES = {
    "Identity": OP_NOP, "TranslateX": OP_TRANSLATE_X, "TranslateY": OP_TRANSLATE_Y,
    "Rotate90": OP_ROT90, "Cutout": OP_CUTOUT, "Brightness": OP_BRIGHTNESS,
    "Contrast": OP_CONTRAST, "Color": OP_COLOR, "AutoContrast": OP_AUTOCONTRAST,
    "AutoSaturation": OP_AUTOSATURATION, "Posterize": OP_POSTERIZE,
    "Sharpness": OP_SHARPNESS, "MidfreqAug": OP_MIDFREQ, "Grayscale": OP_GRAYSCALE,
    "ChromaDrop": OP_CHROMADROP, "SolarizeAdd": OP_SOLARIZE_ADD, "Invert": OP_INVERT,
    "FreqEnhance": OP_FREQ_ENHANCE, "Equalize": OP_EQUALIZE, "Solarize": OP_SOLARIZE,
}
# Dispatchable in the reference but outside every default DCT AUGLIST
# (utils/configs.py:29,93): arbitrary-angle DCT->DFT warps and histogram ops.
UNSUPPORTED_OPS = ("Rotate", "ShearX", "ShearY")

MAX_OPS = 4          # plan slots per image (reference default num_ops = 2)
N_FILTER_SLOTS = 48  # distinct 8x8 multiplicative filters per launch
CLAMP_MIN = -1024
CLAMP_MAX = 1016

# Default DCT op lists (utils/configs.py:29 and :93).
AUGLIST_VITS = ("AutoContrast,Posterize,Color,Contrast,Brightness,Sharpness,Cutout,TranslateX,"
                "TranslateY,Rotate90,AutoSaturation,Grayscale,MidfreqAug,ChromaDrop").split(",")
AUGLIST_VITTI = ("AutoContrast,Posterize,SolarizeAdd,Color,Contrast,Brightness,MidfreqAug,Cutout,"
                 "TranslateX,TranslateY,Rotate90,AutoSaturation,Grayscale,ChromaDrop").split(",")

# numpy mirror of `struct rgbnm_plan` (include/rgbnm_b200.h); 96 bytes, 16-byte aligned.
PLAN_OP_DTYPE = np.dtype([("code", np.int16), ("p", np.int16, (8,)), ("pad", np.int16),
                          ("f", np.float32)], align=False)          # 2+16+2+4 = 24 B? -> see assert
PLAN_DTYPE = np.dtype([
    ("crop_i", np.int16), ("crop_j", np.int16), ("crop_size", np.int16),
    ("flip", np.int16), ("n_ops", np.int16), ("clamp_in", np.int16),
    ("needs_stats", np.int16), ("train", np.int16),
    ("ops", PLAN_OP_DTYPE, (MAX_OPS,)),
])
assert PLAN_OP_DTYPE.itemsize == 24, PLAN_OP_DTYPE.itemsize
assert PLAN_DTYPE.itemsize == 16 + 24 * MAX_OPS, PLAN_DTYPE.itemsize


@dataclass
class PlanOp:
    code: int
    p: List[int] = field(default_factory=lambda: [0] * 8)   # integer parameters
    f: float = 0.0                                          # float parameter (fp32 exact)
    name: str = ""


@dataclass
class Plan:
    """Resolved per-image plan.  Crop geometry is in Y blocks; chroma uses //2."""
    crop_i: int
    crop_j: int
    crop_size: int
    flip: bool = False
    train: bool = False          # RandAugment stage present -> entry clamp applies
    ops: List[PlanOp] = field(default_factory=list)

    @property
    def needs_stats(self) -> bool:
        return any(o.code in (OP_BRIGHTNESS, OP_AUTOCONTRAST, OP_AUTOSATURATION, OP_EQUALIZE, OP_SOLARIZE) for o in self.ops)


# --------------------------------------------------------------------------
# Filter / LUT tables (computed with the reference's own torch/scipy calls so
# that the constants are bit-identical).
# --------------------------------------------------------------------------
class FilterBank:
    """De-duplicating store of 8x8 multiplicative filters used by Sharpness and
    MidfreqAug.  Slot 0 is the all-ones filter."""

    def __init__(self, slots: int = N_FILTER_SLOTS):
        self.table = np.ones((slots, 8, 8), dtype=np.float32)
        self._index = {}
        self._n = 1

    def _put(self, key, mat: torch.Tensor) -> int:
        if key in self._index:
            return self._index[key]
        if self._n >= self.table.shape[0]:
            raise RuntimeError("rgbnm: FilterBank full (too many distinct filter magnitudes)")
        self.table[self._n] = mat.numpy()
        self._index[key] = self._n
        self._n += 1
        return self._index[key]

    def sharpness(self, intensity: float) -> int:
        # dct_ops.py:696-698 -- outer product of two clamped linspace ramps.
        if ("sharp", float(intensity)) in self._index:
            return self._index[("sharp", float(intensity))]
        f_h = torch.linspace(1, (1 + 2 * intensity), 8, dtype=torch.float32).unsqueeze(1).clamp(min=0)
        f_w = torch.linspace(1, (1 + 2 * intensity), 8, dtype=torch.float32).unsqueeze(0).clamp(min=0)
        return self._put(("sharp", float(intensity)), f_h.mm(f_w))

    def midfreq(self, intensity: float) -> int:
        # dct_ops.py:725-741 -- gaussian window applied in block-shifted coordinates;
        # blockshift(x)[i] = x[(i-4) % 8], so in unshifted coordinates the filter is
        # F[(i+4)%8][(j+4)%8].
        if ("mid", float(intensity)) in self._index:
            return self._index[("mid", float(intensity))]
        import scipy.signal
        sig = 8 // 2 - (8 // 8 * 2.2) * abs(intensity)
        f_h = torch.tensor(scipy.signal.windows.gaussian(8, sig), dtype=torch.float32).unsqueeze(1)
        f_w = torch.tensor(scipy.signal.windows.gaussian(8, sig), dtype=torch.float32).unsqueeze(0)
        mat = f_h.mm(f_w)
        if intensity >= 0:
            mat = 1 / mat
        mat = torch.roll(mat, shifts=(4, 4), dims=(0, 1))
        return self._put(("mid", float(intensity)), mat)


def posterize_lut() -> np.ndarray:
    """int16 LUT [6][2048]: DC value (index = dc + 1024) -> posterised DC, for
    bit offsets 0..5 (dct_ops.py:903-912, evaluated with the same torch ops)."""
    lut = np.zeros((6, 2048), dtype=np.int16)
    dc = torch.arange(CLAMP_MIN, CLAMP_MAX + 1, dtype=torch.int16).to(torch.float32)
    for b in range(6):
        x = dc - CLAMP_MIN
        x = x / 2 ** b
        idx = torch.round(x).to(torch.int64)
        table = torch.linspace(CLAMP_MIN, CLAMP_MAX, round((CLAMP_MAX - CLAMP_MIN) / (2 ** b)) + 1)
        out = torch.round(table[idx]).to(torch.int16)
        lut[b, : out.numel()] = out.numpy()
    return lut


# --------------------------------------------------------------------------
# Crop geometry
# --------------------------------------------------------------------------
def _factors(n: int) -> List[int]:
    return list(itertools.chain.from_iterable((i, n // i) for i in range(1, int(n ** 0.5) + 1) if n % i == 0))


@functools.lru_cache(maxsize=None)
def _even_choices(size: int) -> Tuple[int, ...]:
    # pure function of `size`, no RNG draw: the sorted even factors (custom_transforms.py:549-555) as Python ints
    return tuple(sorted(c for c in set(_factors(size)) if c % 2 == 0))


def _choose_closest(val, choices: Sequence[int], maxval: int):
    """custom_transforms.py:571-578 / :860-867 without tensor ops (this ran 25 % of the per-plan host time):
    val <= largest choice: the closest choice, first one on ties (torch.argmin returns the first minimum);
    else the closest multiple of the largest choice, minus one multiple if beyond maxval.  The reference's quotient is
    `int / int64 tensor` = Tensor.__rdiv__ = reciprocal() * val in float32 (91 / 14 gives 6.5000005, not 6.5), rounded
    half-even by torch.round: reproduced with numpy float32 scalars."""
    top = choices[-1]
    if val <= top:
        return min(choices, key=lambda c: abs(c - val))
    closest = int(np.round(np.float32(val) * (np.float32(1.0) / np.float32(top)))) * top
    if closest > maxval:
        closest -= top
    return closest


def eval_crop(height: int, width: int, size_resize: int = 32, size_crop: int = 28) -> Tuple[int, int, int, int]:
    """ResizedCenterCrop_DCT.get_params for the luma plane (custom_transforms.py:850-882)."""
    choices = _even_choices(size_crop)
    ratio = size_crop / size_resize
    w = round(ratio * width)
    h = round(ratio * height)
    w = _choose_closest(w, choices, width)
    h = _choose_closest(h, choices, height)
    i = int(torch.div((height - h), 2, rounding_mode="floor"))
    j = int(torch.div((width - w), 2, rounding_mode="floor"))
    i = i // 2 * 2
    j = j // 2 * 2
    return i, j, int(max(1, h)), int(max(1, w))


def train_crop(height: int, width: int, size: int = 28, scale=(0.05, 1.0)) -> Tuple[int, int, int, int]:
    """RandomResizedCrop_DCT.get_params with ratio=(1,1) (custom_transforms.py:580-629).
    Issues the same torch global-RNG calls in the same order."""
    choices = _even_choices(size)
    area = height * width
    for _ in range(10):
        target_area = area * torch.empty(1).uniform_(scale[0], scale[1]).item()
        w = int(round(math.sqrt(target_area)))
        w = _choose_closest(w, choices, width)
        h = w
        w = int(max(2, w))
        h = int(max(2, h))
        if w <= width and h <= height:
            i = int(torch.randint(0, height - h + 1, size=(1,)).item() // 2 * 2)
            j = int(torch.randint(0, width - w + 1, size=(1,)).item() // 2 * 2)
            return i, j, h, w
    # central-crop fallback (:612-629); ratio == (1, 1)
    in_ratio = float(width) / float(height)
    if in_ratio < 1:
        w = width
        h = int(round(w / 1))
    elif in_ratio > 1:
        h = height
        w = int(round(h * 1))
    else:
        w, h = width, height
    h = _choose_closest(h, choices, height)
    w = _choose_closest(w, choices, width)
    i = int(torch.div((height - h), 2, rounding_mode="floor").div(2, rounding_mode="floor") * 2)
    j = int(torch.div((width - w), 2, rounding_mode="floor").div(2, rounding_mode="floor") * 2)
    return i, j, int(max(1, h)), int(max(1, w))


# --------------------------------------------------------------------------
# RandAugment parameter resolution
# --------------------------------------------------------------------------
@functools.lru_cache(maxsize=None)
def _augmentation_space(num_bins: int, image_size: Tuple[int, int]):
    # custom_transforms.py:1066-1092 (only magnitudes + signedness are needed).  The reference rebuilds this table of
    # linspaces for every image; it is a pure function of its arguments and draws no random numbers, so it is cached
    # (the consumers only read it).
    table = {
        "Identity": (torch.tensor(0.0), False),
        "AutoContrast": (torch.tensor(0.0), False),
        "Equalize": (torch.tensor(0.0), False),
        "Invert": (torch.tensor(0.0), False),
        "Rotate": (torch.linspace(0.0, 30.0, num_bins), True),
        "Posterize": (torch.linspace(0.0, 5.0, num_bins).round().int(), False),
        "Solarize": (torch.linspace(818, -818, num_bins), False),
        "SolarizeAdd": (torch.linspace(0, 883, num_bins), False),
        "Color": (torch.linspace(0.0, 0.9, num_bins), True),
        "Contrast": (torch.linspace(0.0, 0.9, num_bins), True),
        "Brightness": (torch.linspace(0.0, 0.9, num_bins), True),
        "Sharpness": (torch.linspace(0.0, 0.9, num_bins), True),
        "ShearX": (torch.linspace(0.0, 17.0, num_bins), True),
        "ShearY": (torch.linspace(0.0, 17.0, num_bins), True),
        "Cutout": (torch.linspace(0, 6, num_bins), False),
        "TranslateX": (torch.linspace(0.0, 150.0 / 336.0 * image_size[1], num_bins), True),
        "TranslateY": (torch.linspace(0.0, 150.0 / 336.0 * image_size[0], num_bins), True),
        "Rotate90": (torch.tensor(1), True),
        "AutoSaturation": (torch.tensor(0.0), False),
        "Grayscale": (torch.tensor(0.0), False),
        "MidfreqAug": (torch.linspace(0.0, 0.9, num_bins), True),
        "FreqEnhance": (torch.linspace(0.0, 0.9, num_bins), True),
        "ChromaDrop": (torch.tensor(0.0), False),
    }
    # the consumers only read single bins: hand out Python numbers (float(t[bin].item()) / t.item() of the reference, once)
    return {k: ([float(x) for x in m.tolist()] if m.ndim > 0 else m.item(), sgn) for k, (m, sgn) in table.items()}


def cutout_rect(size: int, centre_h: int, centre_w: int, H: int, W: int) -> Tuple[int, int, int, int]:
    """Zeroed block rectangle rows [r0, r1) x cols [c0, c1) of cutout_dct
    (dct_ops.py:796-807).  Note the rows are mirrored about the image: the mask is
    built with F.pad((left, right, upper, lower)) where `upper` = H - centre - pad."""
    lower_pad = max(0, centre_h - size)
    upper_pad = max(0, H - centre_h - size)
    left_pad = max(0, centre_w - size)
    right_pad = max(0, W - centre_w - size)
    return upper_pad, H - lower_pad, left_pad, W - right_pad


def resolve_op(op_name: str, magnitude: float, grid: int, bank: FilterBank) -> PlanOp:
    """Turn (op name, signed magnitude) into a resolved PlanOp, drawing the
    op-internal random numbers exactly where _apply_op_dct would
    (custom_transforms.py:944-1021)."""
    if op_name in UNSUPPORTED_OPS:
        raise NotImplementedError(
            f"rgbnm: DCT op '{op_name}' is outside the B200 hot path (SURVEY.md 8a row a21); "
            f"it is not in any default DCT AUGLIST")
    if op_name not in OP_NAMES:
        raise ValueError(f"The provided operator {op_name} is not recognized.")
    code = OP_NAMES[op_name]
    op = PlanOp(code=code, name=op_name)
    H = W = grid
    if code in (OP_TRANSLATE_X, OP_TRANSLATE_Y):
        t = int(magnitude - (magnitude % 2))           # python float %, :958
        op.p[0] = t
        op.p[1] = t // 2
    elif code == OP_ROT90:
        # rotate_dct_90deg(rotate=magnitude), magnitude = +-1 (dct_ops.py:111-128)
        op.p[0] = 1 if magnitude > 0 else -1
    elif code == OP_CUTOUT:
        size = round(magnitude)
        size = int(size - (size % 2))
        ch = (torch.randint(low=0, high=H, size=(1,)).item()) // 2 * 2
        cw = (torch.randint(low=0, high=W, size=(1,)).item()) // 2 * 2
        op.p[0:4] = list(cutout_rect(size, ch, cw, H, W))
        op.p[4:8] = list(cutout_rect(size // 2, ch // 2, cw // 2, H // 2, W // 2))
    elif code == OP_BRIGHTNESS:
        # coeff_dc += mean(|dc|) * ((1 + m) - 1)   (dct_ops.py:832); (1+m)-1 == m in double
        op.f = float(np.float32((1.0 + magnitude) - 1.0))
    elif code in (OP_CONTRAST, OP_COLOR):
        op.f = float(np.float32(1.0 + magnitude))     # fp32 scalar multiply (dct_ops.py:856)
    elif code == OP_POSTERIZE:
        op.p[0] = int(magnitude)
    elif code == OP_SHARPNESS:
        op.p[0] = bank.sharpness(magnitude)
    elif code == OP_MIDFREQ:
        op.p[0] = bank.midfreq(magnitude)
    elif code == OP_CHROMADROP:
        op.p[0] = 0 if torch.rand(1).item() > 0.5 else 1   # >0.5 drops Cb, else Cr (:1012-1015)
    elif code == OP_SOLARIZE_ADD:
        op.p[0] = int(magnitude)
    elif code == OP_SOLARIZE:
        op.f = float(np.float32(magnitude))           # threshold on the luma DC term: `dcBlocks > threshold` (dct_ops.py:646)
    elif code == OP_FREQ_ENHANCE:
        op.f = float(np.float32(1.0 + magnitude))     # freq_enhance_dct(coeff, 1.0 + magnitude): fp32 multiply (dct_ops.py:1029)
    return op


def sample_randaugment(ops_list: Sequence[str], num_ops: int, magnitude_bin: int, grid: int,
                       bank: FilterBank, num_bins: int = 11) -> List[PlanOp]:
    """RandAugment_dct.forward's draw sequence (custom_transforms.py:1109-1124)."""
    if len(ops_list) == 0:
        return []
    ops_list = list(ops_list).copy()
    op_meta = _augmentation_space(num_bins, (grid, grid))
    chromas = {"Grayscale", "Color", "AutoSaturation", "ChromaDrop"}
    out: List[PlanOp] = []
    for _ in range(num_ops):
        op_index = int(torch.randint(len(ops_list), (1,)).item())
        op_name = ops_list[op_index]
        if op_name in chromas:
            # NB: list(set(..)) ordering is hash dependent, exactly as in the reference.
            if op_name == "Grayscale":
                ops_list = list(set(ops_list).difference(chromas))
            else:
                ops_list = list(set(ops_list).difference({"Grayscale"}))
        magnitudes, signed = op_meta[op_name]
        magnitude = magnitudes[magnitude_bin] if isinstance(magnitudes, list) else magnitudes
        if signed and torch.randint(2, (1,)):
            magnitude *= -1.0
        out.append(resolve_op(op_name, magnitude, grid, bank))
    return out


def sample_train_plan(height: int, width: int, ops_list: Optional[Sequence[str]], num_ops: int,
                      magnitude_bin: int, bank: FilterBank, size: int = 28) -> Plan:
    """One image's worth of `get_transform('imagenet_dct','train')` draws
    (datasets.py:355-361): RandomResizedCrop_DCT -> RandomFlip_DCT -> RandAugment_dct."""
    if len(ops_list if ops_list is not None else AUGLIST_VITS) > 0 and num_ops > MAX_OPS:
        raise ValueError(f"rgbnm: num_ops={num_ops} exceeds plan capacity {MAX_OPS}")
    i, j, h, w = train_crop(height, width, size)
    if h != w:
        raise NotImplementedError("rgbnm: non-square DCT crops are outside the hot path")
    flip = not (torch.rand(1) > 0.5)                  # custom_transforms.py:934
    if ops_list is None:
        ops_list = ["AutoContrast", "Equalize", "Invert", "Rotate", "Posterize", "Solarize", "SolarizeAdd",
                    "Color", "Contrast", "Brightness", "Sharpness", "ShearX", "ShearY", "Cutout",
                    "TranslateX", "TranslateY"]
    ops = sample_randaugment(ops_list, num_ops, magnitude_bin, size, bank)
    return Plan(crop_i=i, crop_j=j, crop_size=h, flip=bool(flip), train=len(ops_list) > 0, ops=ops)


def eval_plan(height: int, width: int, size_resize: int = 32, size_crop: int = 28) -> Plan:
    """`get_transform('imagenet_dct','test')` (datasets.py:362-366): deterministic."""
    i, j, h, w = eval_crop(height, width, size_resize, size_crop)
    if h != w:
        raise NotImplementedError("rgbnm: non-square DCT crops are outside the hot path")
    return Plan(crop_i=i, crop_j=j, crop_size=h, flip=False, train=False, ops=[])


def eval_plan_swin(height: int, width: int, size: int = 32) -> Plan:
    """`get_transform('imagenet_dct_swin','test')` (datasets.py:378-382): Resize_DCT(32) of the whole image
    (custom_transforms.py:468-513) = a crop window covering every block, resized to `size`."""
    if height != width:
        raise NotImplementedError("rgbnm: non-square DCT resizes are outside the hot path")
    return Plan(crop_i=0, crop_j=0, crop_size=height, flip=False, train=False, ops=[])


# crop sides the fused kernel resizes (x2 down / identity / x2 up) per output grid.  28: ViT (datasets.py:355-366),
# 32: SwinV2 (datasets.py:370-382).  Smaller sides are unreachable for 64x64-block inputs with scale=(0.05, 1):
# sqrt(0.05 * 4096) = 14.3 snaps to 14 (factors of 28) / 16 (factors of 32).
SUPPORTED_CROPS = {28: (14, 28, 56), 32: (16, 32, 64)}


_PLAN_HEAD = struct.Struct("<8h")
_PLAN_OP = struct.Struct("<h8hhf")
assert _PLAN_HEAD.size + MAX_OPS * _PLAN_OP.size == PLAN_DTYPE.itemsize


def pack_plans(plans: Sequence[Plan], clamp_in: Optional[Sequence[bool]] = None, out_size: int = 28) -> np.ndarray:
    """Pack plans into the `struct rgbnm_plan` array the C-ABI takes (struct.pack_into: ~10x faster than per-field
    assignments into a numpy structured array)."""
    item = PLAN_DTYPE.itemsize
    buf = bytearray(len(plans) * item)
    allowed = SUPPORTED_CROPS[out_size]
    for n, pl in enumerate(plans):
        if pl.crop_size not in allowed:
            raise ValueError(f"rgbnm: unsupported crop size {pl.crop_size} -> {out_size} blocks")
        if len(pl.ops) > MAX_OPS:
            raise ValueError("rgbnm: too many ops in plan")
        base = n * item
        _PLAN_HEAD.pack_into(buf, base, pl.crop_i, pl.crop_j, pl.crop_size, int(pl.flip), len(pl.ops),
                             1 if clamp_in is None else int(bool(clamp_in[n])), int(pl.needs_stats), int(pl.train))
        for k, op in enumerate(pl.ops):
            _PLAN_OP.pack_into(buf, base + _PLAN_HEAD.size + k * _PLAN_OP.size, op.code, *[int(v) for v in op.p], 0, float(op.f))
    return np.frombuffer(buf, dtype=PLAN_DTYPE).copy()


# --------------------------------------------------------------------------
# Batched sampler: the same DISTRIBUTION as `sample_train_plan`, a whole batch per call
# --------------------------------------------------------------------------
_CHROMAS = ("Grayscale", "Color", "AutoSaturation", "ChromaDrop")
_STATS_CODES = (OP_BRIGHTNESS, OP_AUTOCONTRAST, OP_AUTOSATURATION, OP_EQUALIZE, OP_SOLARIZE)


class BatchedSampler:
    """Draws `n` training plans at once and returns them PACKED (`struct rgbnm_plan[n]`, the array `FusedDCT.run` takes).

    `sample_train_plan` replays the reference's torch-global-RNG call sequence image by image (the RNG contract of SURVEY.md
    8a: seed-for-seed identical plans, ~60 us of Python per image -- as much host time per batch as the GPU step takes).  This
    sampler draws from the same distributions -- RandomResizedCrop_DCT (custom_transforms.py:580-629), RandomFlip_DCT (:934),
    RandAugment_dct incl. the chroma-op exclusion rules and the op-internal draws (:944-1021, 1109-1124) -- but takes all
    uniforms of a batch from one `torch.Generator` call and resolves them with numpy, ~0.3 ms per 256 images.  The STREAM
    differs from the reference's (same seed -> different but identically distributed plans); parity tests use the per-image
    sampler, throughput paths (bench from-JPEG arm, compat loader) this one."""

    def __init__(self, height: int, width: int, ops_list: Optional[Sequence[str]], num_ops: int, magnitude_bin: int,
                 bank: FilterBank, size: int = 28, num_bins: int = 11, scale=(0.05, 1.0)):
        if ops_list is None:
            raise NotImplementedError("rgbnm: the batched sampler needs an explicit DCT ops_list")
        if len(ops_list) > 0 and num_ops > MAX_OPS:
            raise ValueError(f"rgbnm: num_ops={num_ops} exceeds plan capacity {MAX_OPS}")
        self.h, self.w, self.size, self.scale = height, width, size, scale
        self.num_ops = num_ops if len(ops_list) > 0 else 0
        self.train = len(ops_list) > 0
        choices = _even_choices(size)
        vmax = int(round(math.sqrt(height * width))) + 2
        # crop side for every possible rounded sqrt(target area): the scalar rule, tabulated
        self.side_of = np.array([max(2, _choose_closest(v, choices, width)) for v in range(vmax + 1)], dtype=np.int64)
        if self.side_of.max() > min(height, width):
            raise NotImplementedError("rgbnm: batched sampler expects crops that always fit (square inputs)")
        allowed = SUPPORTED_CROPS[size]
        lo = int(round(math.sqrt(height * width * scale[0])))
        if any(int(s) not in allowed for s in self.side_of[lo:]):
            raise ValueError(f"rgbnm: a {height}x{width}-block input can draw crop sides outside {allowed}")
        # the three op lists RandAugment_dct can be in (full -> minus Grayscale after a chroma op / minus all chroma ops after Grayscale)
        full = list(ops_list)
        for name in full:
            if name in UNSUPPORTED_OPS:
                raise NotImplementedError(f"rgbnm: DCT op '{name}' is outside the B200 hot path (SURVEY.md 8a row a21)")
            if name not in OP_NAMES:
                raise ValueError(f"The provided operator {name} is not recognized.")
        self.lists = [full, [o for o in full if o != "Grayscale"], [o for o in full if o not in _CHROMAS]]
        meta = _augmentation_space(num_bins, (size, size))
        self.names = full
        idx = {n: i for i, n in enumerate(full)}
        self.list_idx = [np.array([idx[o] for o in lst], dtype=np.int64) for lst in self.lists]
        # per (op, sign) template of the resolved op; Cutout / ChromaDrop parameters are filled from their own draws
        k = len(full)
        self.t_code = np.zeros(k, dtype=np.int16)
        self.t_signed = np.zeros(k, dtype=bool)
        self.t_p = np.zeros((k, 2, 8), dtype=np.int16)
        self.t_f = np.zeros((k, 2), dtype=np.float32)
        self.cut_size = np.zeros(k, dtype=np.int64)
        state = torch.random.get_rng_state()
        for i, name in enumerate(full):
            mags, signed = meta[name]
            mag = mags[magnitude_bin] if isinstance(mags, list) else mags
            self.t_signed[i] = signed
            for s, m in enumerate((mag, -1.0 * mag)):
                op = resolve_op(name, m, size, bank)          # Cutout / ChromaDrop draw here: state restored below
                self.t_code[i] = op.code
                self.t_p[i, s] = op.p
                self.t_f[i, s] = op.f
            if OP_NAMES[name] == OP_CUTOUT:
                sz = round(mag)
                self.cut_size[i] = int(sz - (sz % 2))
        torch.random.set_rng_state(state)
        self.is_chroma = np.array([n in _CHROMAS for n in full])
        self.is_gray = np.array([n == "Grayscale" for n in full])
        self.needs = np.isin(self.t_code, np.array(_STATS_CODES, dtype=np.int16))

    def sample(self, n: int, clamp_in=None, generator: Optional[torch.Generator] = None) -> np.ndarray:
        """-> packed PLAN_DTYPE array of n plans.  `generator`: a CPU torch.Generator (default: the global one)."""
        K = self.num_ops
        u = torch.rand((n, 4 + 5 * max(K, 1)), dtype=torch.float64, generator=generator).numpy()
        out = np.zeros(n, dtype=PLAN_DTYPE)
        area = self.h * self.w
        target = area * (self.scale[0] + (self.scale[1] - self.scale[0]) * u[:, 0])
        side = self.side_of[np.rint(np.sqrt(target)).astype(np.int64)]
        out["crop_size"] = side
        out["crop_i"] = np.floor(u[:, 1] * (self.h - side + 1)).astype(np.int64) // 2 * 2
        out["crop_j"] = np.floor(u[:, 2] * (self.w - side + 1)).astype(np.int64) // 2 * 2
        out["flip"] = u[:, 3] <= 0.5
        out["train"] = int(self.train)
        out["n_ops"] = K
        out["clamp_in"] = 1 if clamp_in is None else np.asarray(clamp_in).astype(np.int16)
        state = np.zeros(n, dtype=np.int64)                     # index into self.lists
        needs = np.zeros(n, dtype=bool)
        G = self.size
        for k in range(K):
            c = u[:, 4 + 5 * k: 9 + 5 * k]
            lens = np.array([len(x) for x in self.list_idx])[state]
            pick = np.minimum((c[:, 0] * lens).astype(np.int64), lens - 1)
            op = np.empty(n, dtype=np.int64)
            for s in range(3):
                m = state == s
                if m.any():
                    op[m] = self.list_idx[s][pick[m]]
            sign = (self.t_signed[op] & (c[:, 1] >= 0.5)).astype(np.int64)       # torch.randint(2): 1 -> negative magnitude
            code = self.t_code[op]
            p = self.t_p[op, sign].copy()
            f = self.t_f[op, sign]
            cut = code == OP_CUTOUT
            if cut.any():
                # dct_ops.py:796-807 via cutout_rect: centre = randint(0, G) // 2 * 2 per axis, rows mirrored about the image
                ch = np.floor(c[cut, 2] * G).astype(np.int64) // 2 * 2
                cw = np.floor(c[cut, 3] * G).astype(np.int64) // 2 * 2
                sz = self.cut_size[op[cut]]
                for col0, (cen_h, cen_w, sze, H) in ((0, (ch, cw, sz, G)), (4, (ch // 2, cw // 2, sz // 2, G // 2))):
                    p[cut, col0 + 0] = np.maximum(0, H - cen_h - sze)
                    p[cut, col0 + 1] = H - np.maximum(0, cen_h - sze)
                    p[cut, col0 + 2] = np.maximum(0, cen_w - sze)
                    p[cut, col0 + 3] = H - np.maximum(0, H - cen_w - sze)
            drop = code == OP_CHROMADROP
            if drop.any():
                p[drop, 0] = np.where(c[drop, 4] > 0.5, 0, 1)
            out["ops"]["code"][:, k] = code
            out["ops"]["p"][:, k] = p
            out["ops"]["f"][:, k] = f
            needs |= self.needs[op]
            gray, chroma = self.is_gray[op], self.is_chroma[op]
            state = np.where(gray, 2, np.where(chroma & (state == 0), 1, state))
        out["needs_stats"] = needs
        return out
