"""B200-native DCT-domain ViT hot path (drop-in for RGB-no-more's --domain=dct path).

Sub-modules:
  lib         ctypes binding of the C-ABI shared library (include/rgbnm_b200.h)
  dct_manip   drop-in for the reference's `dct_manip` extension (read_coefficients)
  plan        host-side augmentation-plan sampler (replays the reference RNG contract)
  transforms  `get_transform`-compatible surface running the fused CUDA kernel
  synth       deterministic synthetic inputs
"""
from . import lib  # noqa: F401

__all__ = ["lib"]
