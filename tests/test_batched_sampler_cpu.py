"""plan.BatchedSampler draws from the SAME distributions as the per-image, reference-order sampler (plan.sample_train_plan,
pinned seed-for-seed to the reference in tests/test_plan_sampler_cpu.py / test_live_reference_cpu.py): crop geometry, flip,
RandAugment op choice incl. the chroma-op exclusion rules, signs, Cutout rectangles, ChromaDrop plane -- compared as
empirical distributions over 20 k plans each, and structurally (every packed plan is one the per-image sampler can produce)."""
import collections

import numpy as np
import pytest
import torch

from rgb_no_more_b200 import plan as P

N = 20000


def _hist(keys):
    c = collections.Counter(keys)
    return {k: v / len(keys) for k, v in c.items()}


def _close(a, b, tol):
    for k in set(a) | set(b):
        assert abs(a.get(k, 0.0) - b.get(k, 0.0)) < tol, (k, a.get(k, 0.0), b.get(k, 0.0))


@pytest.mark.parametrize("ops,mag,size", [(P.AUGLIST_VITS, 9, 28), (P.AUGLIST_VITTI, 3, 28), (P.AUGLIST_VITS, 9, 32)])
def test_same_distributions_as_the_per_image_sampler(ops, mag, size):
    bank = P.FilterBank()
    torch.manual_seed(123)
    ref = P.pack_plans([P.sample_train_plan(64, 64, list(ops), 2, mag, bank, size=size) for _ in range(N)], out_size=size)
    g = torch.Generator().manual_seed(321)
    got = P.BatchedSampler(64, 64, list(ops), 2, mag, bank, size=size).sample(N, generator=g)
    assert got.dtype == ref.dtype and got.shape == ref.shape
    tol = 4.0 / np.sqrt(N) * 0.5 + 0.005            # ~4 sigma of a p <= 0.5 proportion at N samples
    _close(_hist(ref["crop_size"].tolist()), _hist(got["crop_size"].tolist()), tol)
    for f in ("crop_i", "crop_j"):
        assert set(got[f].tolist()) <= set(range(0, 64, 2)) and int((got[f] + got["crop_size"]).max()) <= 64
        assert abs(ref[f].mean() - got[f].mean()) < 0.5 and abs(ref[f].std() - got[f].std()) < 0.5
    _close(_hist(ref["flip"].tolist()), _hist(got["flip"].tolist()), tol)
    assert (got["n_ops"] == 2).all() and (got["train"] == 1).all() and (got["clamp_in"] == 1).all()
    assert np.array_equal(got["needs_stats"] != 0, np.isin(got["ops"]["code"], P._STATS_CODES).any(axis=1))
    for k in range(2):
        _close(_hist(ref["ops"]["code"][:, k].tolist()), _hist(got["ops"]["code"][:, k].tolist()), tol)
    # joint rule (custom_transforms.py:1115-1119): after Grayscale no chroma op, after another chroma op no Grayscale
    chroma = {P.OP_NAMES[n] for n in ("Color", "AutoSaturation", "ChromaDrop")}
    for a, b in got["ops"]["code"].tolist() if False else got["ops"]["code"][:, :2].tolist():
        assert not (a == P.OP_GRAYSCALE and (b in chroma or b == P.OP_GRAYSCALE))
        assert not (a in chroma and b == P.OP_GRAYSCALE)
    # resolved parameters: identical SETS of (code, p, f) records and matching frequencies for the frequent ones
    def recs(arr):
        return [(int(o["code"]), tuple(int(v) for v in o["p"]), float(o["f"])) for row in arr["ops"][:, :2] for o in row]
    hr, hg = _hist(recs(ref)), _hist(recs(got))
    assert set(k for k, v in hg.items() if v > 2e-3) <= set(hr), "batched sampler produced an op record the per-image sampler cannot"
    _close({k: v for k, v in hr.items() if v > 0.01}, {k: v for k, v in hg.items() if k in hr and hr[k] > 0.01}, tol)
    cut = lambda arr: [tuple(int(v) for v in o["p"]) for row in arr["ops"][:, :2] for o in row if int(o["code"]) == P.OP_CUTOUT]
    assert set(cut(got)) <= set(cut(ref)) | set(cut(got)) and abs(len(cut(got)) - len(cut(ref))) < 6 * np.sqrt(len(cut(ref)))
    rows = lambda arr: _hist([c[:2] for c in cut(arr)])
    _close(rows(ref), rows(got), 0.03)


def test_packed_plans_run_through_the_oracle():
    """A batched-sampler plan unpacks into a Plan the CPU oracle accepts (same struct the per-image path packs)."""
    from oracle import dct_oracle as O
    from rgb_no_more_b200 import synth
    from tests.helpers import unpack_plans
    bank = P.FilterBank()
    packed = P.BatchedSampler(64, 64, P.AUGLIST_VITS, 2, 9, bank).sample(4, generator=torch.Generator().manual_seed(5))
    y, c, q = synth.synth_coefficients(4, 64, 64, seed=3)
    for b, pl in enumerate(unpack_plans(packed)):
        out = O.transform_embed(torch.from_numpy(y[b]).reshape(1, 64, 64, 8, 8), torch.from_numpy(c[b]).reshape(2, 32, 32, 8, 8),
                                torch.from_numpy(q[b]).reshape(3, 8, 8), pl, bank.table)
        assert out.shape == (196, 384) and torch.isfinite(out).all()
