"""Recipe for oracle/_ref: the REFERENCE ITSELF, staged so that it can travel to the GPU box.

TEST / BASELINE INFRASTRUCTURE ONLY.  The reference (JeongsooP/RGB-no-more) is a Python program; nothing of it is
compiled.  `/root/reference` exists only in the build container, so `__graft_entry__.build()` runs this script there and it
copies the reference's own, unmodified Python sources into `oracle/_ref/` (git-ignored: never part of the history; NOT
gpurun-ignored: it travels with the snapshot like the built .so files).  Users of oracle/_ref:

  * bench.py --impl reference      the reference's own DataLoader + get_transform + pvit.ViT + optimiser loop on the host cores
  * bench.py (torch_b200_baseline) the reference's own pvit.ViT under plain PyTorch on the B200
  * tests/test_compat_launcher_*   the reference's train.py driven through rgb_no_more_b200.compat (boundary B4)

The native part of the reference (dct_manip/dct_manip.cpp) is NOT built: it needs jpeglib.h, which this image does not have
(DESIGN.md section 2); the staged tree imports `dct_manip` by name and gets the B1 drop-in (oracle/ref_shims/dct_manip.py).
Third-party modules the reference imports but this image lacks (yacs, timm, torchmetrics, torchinfo, fvcore, sysrsync) are
provided by small stand-ins under rgb_no_more_b200/compat/shims (they implement the few calls the reference makes)."""
from __future__ import annotations

import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("RGBNM_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["train.py", "eval.py", "benchmark.py", "datasets.py", "models/plainvit.py", "models/swinv2.py",
         "utils/cls_transforms.py", "utils/configs.py", "utils/custom_optims.py", "utils/custom_sampler.py",
         "utils/custom_transforms.py", "utils/dct_ops.py", "utils/dct_torch_utils.py", "utils/mp_scripts.py",
         "utils/pipeline_utils.py", "LICENSE"]


def main() -> int:
    if not os.path.isdir(REF):
        print(f"oracle/make_ref: {REF} not present (GPU box): using the prebuilt oracle/_ref" if os.path.isdir(DST)
              else f"oracle/make_ref: neither {REF} nor oracle/_ref present", file=sys.stderr)
        return 0
    n = 0
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or not filecmp.cmp(src, dst, shallow=False):
            shutil.copyfile(src, dst)
            n += 1
    with open(os.path.join(DST, "PROVENANCE.txt"), "w") as f:
        f.write(f"verbatim copies of {len(FILES)} files of {REF} made by oracle/make_ref.py; not tracked by git\n")
    if n:
        print(f"oracle/make_ref: staged {n} reference files into {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
