"""The drop-in boundary on a machine without a GPU: librgbnm_b200.so loads, exports every entry point that
include/rgbnm_b200.h declares, the ctypes table binds every one of them, and GPU entry points fail loudly (no CPU
fallback, no compute calls here)."""
import ctypes as C
import os
import re

import pytest
import torch

from rgb_no_more_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "rgbnm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rgbnm_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    names = _declared_symbols()
    assert len(names) >= 25
    lib = C.CDLL(L.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rgbnm_b200.h but not exported"
        assert n in L.SIGNATURES, f"{n} has no ctypes signature in rgb_no_more_b200/lib.py"
    assert set(L.SIGNATURES) <= set(names)


def test_struct_sizes_match_the_header():
    from rgb_no_more_b200 import gemm as G, plan as P
    assert P.PLAN_DTYPE.itemsize == 112                      # rgbnm_plan
    assert C.sizeof(L.WPrepDesc) == 64                       # rgbnm_wprep_desc
    # rgbnm_gemm_args: 8 pointers, 5 long long, 9 int + 1 float, then (ABI 4) 2 pointers + 1 float (+ 4 bytes tail padding)
    assert C.sizeof(G.GemmArgs) == 8 * 8 + 5 * 8 + 10 * 4 + 2 * 8 + 8
    assert L.load().rgbnm_abi_version() == 5


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_gpu_entry_points_fail_loudly_without_a_device():
    from rgb_no_more_b200 import transforms as TF, vit as V
    with pytest.raises(Exception):
        TF.FusedDCT("cpu")
    m = V.ViT(patch_size=16, emb_size=192, depth=1, n_classes=10, drop_p=0.0, pixel_space="DCT", num_heads=3, head_size=64)
    with pytest.raises(Exception):
        m(torch.zeros(1, 196, 384))


def test_product_modules_never_import_the_oracle():
    """oracle/ is test infrastructure: only smoke/ (the checkers __graft_entry__.smoke() runs), tests/ and bench.py's CPU
    legs may import it -- never a module on the product path."""
    import ast
    import glob
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    offenders = []
    for path in glob.glob(os.path.join(root, "rgb_no_more_b200", "**", "*.py"), recursive=True):
        tree = ast.parse(open(path).read())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            if any(n == "oracle" or n.startswith("oracle.") for n in names):
                offenders.append(os.path.basename(path))
    assert offenders == []
