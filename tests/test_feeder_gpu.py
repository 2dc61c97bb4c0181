"""JpegFeeder (host decode -> pinned ring -> copy stream -> device slots) delivers exactly what decode_batch decodes,
in submission order, across ring wrap-around."""
import pytest
import torch

from rgb_no_more_b200 import dct_manip as dm
from rgb_no_more_b200 import feeder as FD
from rgb_no_more_b200 import synth

pytestmark = pytest.mark.gpu


def test_feeder_round_robin_matches_direct_decode():
    B = 4
    sets = [synth.synth_jpeg_set(B + k)[k:] for k in range(5)]          # five different batches of 4 files
    ref = [dm.decode_batch(s, 64, 64, nthreads=2) for s in sets]
    fd = FD.JpegFeeder("cuda:0", B, 64, 64, slots=3, nthreads=2)
    got = []
    fd.submit(sets[0])
    fd.submit(sets[1])
    for k in range(5):
        y, c, q, flags, slot = fd.get()
        got.append((y.cpu().clone(), c.cpu().clone(), q.cpu().clone(), flags))
        fd.release(slot)
        if k + 2 < 5:
            fd.submit(sets[k + 2])
    fd.close()
    for (y, c, q, fl), (ry, rc, rq, rfl) in zip(got, ref):
        assert torch.equal(y, ry) and torch.equal(c, rc) and torch.equal(q, rq) and fl == rfl.tolist()


def test_feeder_rejects_wrong_batch_and_corrupt_files():
    fd = FD.JpegFeeder("cuda:0", 2, 64, 64)
    with pytest.raises(ValueError):
        fd.submit(synth.synth_jpeg_set(3))
    fd.submit([b"not a jpeg", synth.synth_jpeg(0)])
    with pytest.raises(RuntimeError):
        fd.get()
    fd.close()


def test_end_to_end_training_from_jpeg_bytes_learns(monkeypatch):
    """examples/train_synthetic_dct.py: JPEG bytes -> feeder -> K0 (train recipe) -> mixup -> ViT-Ti -> AdamW.  The class is
    encoded in the mean colour, so the soft-label cross entropy must fall well below ln(1000) within 40 steps."""
    import importlib.util
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("train_synthetic_dct", os.path.join(root, "examples", "train_synthetic_dct.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    monkeypatch.setattr(sys, "argv", ["train_synthetic_dct.py", "--steps", "40", "--batch", "32", "--files", "64", "--classes", "4"])
    losses = mod.main()
    assert losses[0] > 5.0 and losses[-1] < 0.6 * losses[0], losses
