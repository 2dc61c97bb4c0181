"""Compatibility launcher: the reference's `train.py` on the target PyTorch, optionally with the B200 backend.

    python -m rgb_no_more_b200.compat.launch --ref /path/to/RGB-no-more [--backend b200|reference] -- <train.py arguments>

`--backend` defaults to $RGBNM_BACKEND (else "reference").  The reference spawns one process per GPU with
torch.multiprocessing.spawn (train.py:275-297); spawned children start from a fresh interpreter, so the launcher swaps the
spawn target for `_child`, which re-applies `env.activate` (+ `backend.install`) in the child before calling the reference's
own `traineval`."""
from __future__ import annotations

import argparse
import os
import sys


def _child(rank, ref_root, use_b200, *args):
    from rgb_no_more_b200.compat import backend, env
    env.activate(ref_root)
    if os.environ.get("RGBNM_CHILD_LOG"):
        # the reference configures logging in the parent only (train.py:236), so its INFO lines ("Training complete", "Test Acc",
        # the --benchmark FPS table) are dropped in the spawned ranks; opt-in to see them
        import logging
        logging.basicConfig(level=logging.INFO, format="%(message)s")
    import train as ref_train                      # the reference's train.py
    if use_b200:
        backend.install(ref_train.utils)
    return ref_train.traineval(rank, *args)


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    rest = []
    if "--" in argv:
        k = argv.index("--")
        argv, rest = argv[:k], argv[k + 1:]
    ap = argparse.ArgumentParser(prog="rgb_no_more_b200.compat.launch")
    ap.add_argument("--ref", default=os.environ.get("RGBNM_REFERENCE", ""), help="checkout of JeongsooP/RGB-no-more")
    ap.add_argument("--backend", default=os.environ.get("RGBNM_BACKEND", "reference"), choices=["b200", "reference"])
    args = ap.parse_args(argv)
    if not args.ref:
        ap.error("--ref (or $RGBNM_REFERENCE) must point at a checkout of the reference")
    from rgb_no_more_b200.compat import env
    root = env.activate(args.ref)
    use_b200 = args.backend == "b200"
    os.environ["RGBNM_BACKEND"] = args.backend
    os.environ["RGBNM_REFERENCE"] = root
    import train as ref_train
    real_spawn = ref_train.torchmp.spawn

    def spawn(fn, args=(), nprocs=1, **kw):
        return real_spawn(_child, args=(root, use_b200, *args), nprocs=nprocs, **kw)
    ref_train.torchmp.spawn = spawn
    old_argv = sys.argv
    sys.argv = [os.path.join(root, "train.py")] + rest
    try:
        ref_train.main()
    finally:
        sys.argv = old_argv
        ref_train.torchmp.spawn = real_spawn
    return 0


if __name__ == "__main__":
    sys.exit(main())
