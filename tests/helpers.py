"""Shared test helpers: golden loading and plan (de)serialisation."""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

from rgb_no_more_b200 import plan as P  # noqa: E402

CODE_TO_NAME = {v: k for k, v in P.OP_NAMES.items()}


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def unpack_plans(arr):
    """numpy PLAN_DTYPE array -> list[Plan] (names restored from op codes)."""
    out = []
    for a in arr:
        ops = []
        for k in range(int(a["n_ops"])):
            o = a["ops"][k]
            ops.append(P.PlanOp(code=int(o["code"]), p=[int(v) for v in o["p"]], f=float(o["f"]),
                                name=CODE_TO_NAME[int(o["code"])]))
        out.append(P.Plan(crop_i=int(a["crop_i"]), crop_j=int(a["crop_j"]), crop_size=int(a["crop_size"]),
                          flip=bool(a["flip"]), train=bool(a["train"]), ops=ops))
    return out


def lsb_report(a, b):
    """(max |diff|, fraction of entries that differ) for two integer arrays."""
    d = np.abs(a.astype(np.int64) - b.astype(np.int64))
    return int(d.max()), float((d != 0).mean())


def reference_available():
    return os.path.isdir("/root/reference/utils")


def import_reference():
    if "/root/reference" not in sys.path:
        sys.path.insert(1, "/root/reference")
    sys.modules.setdefault("dct_manip", types.ModuleType("dct_manip"))
    import utils.custom_transforms as ctrans
    import utils.dct_ops as dops
    return ctrans, dops


TIE_EPS = 2e-3     # |frac(exact) - 0.5| below which the real-valued resize result counts as a .5 tie


def assert_only_tie_mismatches(got, ref, exact, what=""):
    """int16 resize parity: `got` (CUDA) may differ from `ref` (oracle, fp32 torch-CPU) only by
    one LSB and only where the float64 result `exact` sits on a .5 tie -- there the
    reference's own rounding direction is decided by fp32 noise of its BLAS (summation order,
    the ~1e-8 "zeros" of its conversion matrix) and is not reproducible across machines."""
    got = np.asarray(got).astype(np.int64)
    ref = np.asarray(ref).astype(np.int64)
    exact = np.asarray(exact, dtype=np.float64)
    d = np.abs(got - ref)
    assert d.max() <= 1, (what, int(d.max()))
    frac = np.abs(exact - np.floor(exact) - 0.5)
    bad = (d != 0) & (frac > TIE_EPS)
    assert not bad.any(), (what, int(bad.sum()), float(frac[d != 0].max()))
    # and the CUDA value is always one of the two nearest integers of the exact result
    assert np.all(np.abs(got - exact) <= 0.5 + TIE_EPS), what
    return float((d != 0).mean())


def seeded_state_dict(model, seed=11997733):
    """The weight recipe tools/make_golden.py used for the reference ViT: per key, seeded normal."""
    sd = {}
    for k, v in sorted(model.state_dict().items()):
        g = torch.Generator().manual_seed(seed + sum(ord(ch) * (i + 1) for i, ch in enumerate(k)) % 1000003)
        if k.endswith("weight") and v.ndim == 2:
            sd[k] = torch.randn(v.shape, generator=g) * (0.5 / v.shape[1] ** 0.5)
        elif k.endswith("weight"):
            sd[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        else:
            sd[k] = 0.05 * torch.randn(v.shape, generator=g)
    return sd


def golden_vit_inputs(seed):
    g = torch.Generator().manual_seed(int(seed))
    yf = torch.rand((2, 1, 28, 28, 8, 8), generator=g) * 2 - 1
    cf = torch.rand((2, 2, 14, 14, 8, 8), generator=g) * 2 - 1
    return yf, cf


def seeded_swin_state_dict(model, seed=11997733):
    """Weight recipe of tools/make_golden_swin.py for SwinV2: `seeded_state_dict`, except that the constructed
    buffers (relative_coords_table, relative_position_index, attn_mask) keep their values and logit_scale is
    log(10) + 0.3 * seeded normal (so that both sides of the clamp at log(100), swinv2.py:158, are visited)."""
    sd = seeded_state_dict(model, seed)
    own = model.state_dict()
    for k in own:
        if "relative_coords_table" in k or "relative_position_index" in k or "attn_mask" in k:
            sd[k] = own[k]
        if k.endswith("logit_scale"):
            g = torch.Generator().manual_seed(len(k))
            sd[k] = torch.log(10 * torch.ones_like(own[k])) + 0.3 * torch.randn(own[k].shape, generator=g)
    return sd


def golden_swin_inputs(seed, batch=2):
    g = torch.Generator().manual_seed(int(seed))
    yf = torch.rand((batch, 1, 32, 32, 8, 8), generator=g) * 2 - 1
    cf = torch.rand((batch, 2, 16, 16, 8, 8), generator=g) * 2 - 1
    return yf, cf


def golden_vits_inputs(seed, batch=8):
    """Inputs of tests/golden/vit_s.npz (tools/make_golden.py::golden_vits_inputs)."""
    g = torch.Generator().manual_seed(int(seed))
    decay = torch.rand((8, 8), generator=g).pow(2) * 0.9 + 0.1
    yf = (torch.rand((batch, 1, 28, 28, 8, 8), generator=g) * 2 - 1) * decay
    cf = (torch.rand((batch, 2, 14, 14, 8, 8), generator=g) * 2 - 1) * decay
    return yf, cf


VITS_GOLDEN_KEYS = ("patchembed.projection.0.weight", "patchembed.projection.0.bias", "encoder.0.0.fn.eb_mha.qkv.weight",
                    "encoder.0.0.fn.eb_mha.qkv.bias", "encoder.3.0.fn.eb_mha.projection.weight", "encoder.6.1.fn.eb_ffb.0.weight",
                    "encoder.6.1.fn.eb_ffb.0.bias", "encoder.11.1.fn.eb_ffb.3.weight", "encoder.11.1.fn.eb_ffb.3.bias",
                    "encoder.5.0.fn.eb_lrnorm1.weight", "encoder.9.1.fn.eb_lrnorm2.bias", "classhead.ch_lrnorm.weight",
                    "classhead.ch_linear1.weight", "classhead.ch_linear2.weight", "classhead.ch_linear2.bias")


def vits_soft_labels():
    labels = torch.zeros((8, 1000))
    for b in range(8):
        labels[b, (37 * b + 3) % 1000] = 0.75
        labels[b, (91 * b + 500) % 1000] = 0.25
    return labels
