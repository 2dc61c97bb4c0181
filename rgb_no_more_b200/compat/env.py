"""Make a checkout of the reference importable and runnable on the target PyTorch (yacs / timm / torchmetrics /
torchinfo / fvcore / sysrsync are absent, `CosineAnnealingLR(verbose=)` is rejected by torch 2.11, and `dct_manip` is the native
module this repository replaces)."""
from __future__ import annotations

import importlib
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
SHIMS = os.path.join(_HERE, "shims")
B1 = os.path.join(_HERE, "b1")              # holds dct_manip.py, the B1 drop-in under the reference's import name


def activate(reference_root: str) -> str:
    """Idempotent.  Returns the absolute reference root."""
    root = os.path.abspath(reference_root)
    if not os.path.isfile(os.path.join(root, "utils", "pipeline_utils.py")):
        raise FileNotFoundError(f"rgbnm compat: {root} is not a checkout of RGB-no-more (utils/pipeline_utils.py missing)")
    # import order: B1 drop-in (must beat the reference's empty dct_manip/ source directory) -> reference ->
    # everything already on the path -> stand-ins (a real installation of yacs / timm / ... wins over them)
    for p in (root, B1):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    if SHIMS not in sys.path:
        sys.path.append(SHIMS)
    importlib.invalidate_caches()
    _patch_torch()
    return root


def _patch_torch() -> None:
    import inspect

    import torch
    sched = torch.optim.lr_scheduler
    if "verbose" not in inspect.signature(sched.CosineAnnealingLR.__init__).parameters and not getattr(sched.CosineAnnealingLR, "_rgbnm", False):
        base = sched.CosineAnnealingLR

        class CosineAnnealingLR(base):          # pipeline_utils.py:538 passes verbose=False
            _rgbnm = True

            def __init__(self, optimizer, T_max, eta_min=0.0, last_epoch=-1, verbose=False):
                super().__init__(optimizer, T_max, eta_min=eta_min, last_epoch=last_epoch)
        CosineAnnealingLR.__name__ = base.__name__
        CosineAnnealingLR.__qualname__ = base.__qualname__
        sched.CosineAnnealingLR = CosineAnnealingLR
