"""CPU: the SwinV2 parts of the oracle against outputs of the reference itself
(tests/golden/swin_*.npz, written by tools/make_golden_swin.py from /root/reference)."""
import numpy as np
import torch

from oracle import dct_oracle as O
from rgb_no_more_b200 import dct_manip as dm
from rgb_no_more_b200 import plan as P
from rgb_no_more_b200 import synth
from tests.helpers import load, unpack_plans, lsb_report, golden_swin_inputs


def test_swin_embed_input_matches_reference():
    g = load("swin_embed.npz")
    yf, cf = golden_swin_inputs(g["input_seed"])
    e = O.embed_input_swin(yf, cf).numpy()
    assert e.shape == (2, 64, 64, 24)
    assert np.abs(e - g["embed_in"]).max() < 1e-5
    # the conversion matrices of the restatement are the reference's own
    assert np.abs(O.conversion_matrix(2, 4).numpy() - g["A42"]).max() < 1e-7
    assert np.abs(O.conversion_matrix(4, 2).numpy() - g["A24"]).max() < 1e-7


def test_swin_embed_interleave_is_the_references():
    """The reference splits the 8 decomposed indices as (p1 pdh), i.e. interleaved: sub-block (i % 2, j % 2) takes D[i][j]."""
    yf = torch.zeros((1, 1, 32, 32, 8, 8))
    cf = torch.zeros((1, 2, 16, 16, 8, 8))
    A = O.conversion_matrix(2, 4)
    X = torch.randn(8, 8, generator=torch.Generator().manual_seed(1))
    yf[0, 0, 3, 5] = X
    D = A.T @ X @ A
    e = O.embed_input_swin(yf, cf)[0]
    for i in range(8):
        for j in range(8):
            assert abs(float(e[2 * 3 + i % 2, 2 * 5 + j % 2, (i // 2) * 4 + j // 2]) - float(D[i, j])) < 1e-6


def test_swin_pipeline_matches_reference():
    g = load("swin_pipeline.npz")
    plans = unpack_plans(g["plans"])
    filters = g["filters"]
    images = []
    for i in g["synth_ids"]:
        dims, quant, Y, C = dm.read_coefficients_from_bytes(synth.synth_jpeg(int(i)))
        images.append((Y, C, quant))
    for k, (img, seed, mag) in enumerate(g["cases"]):
        yq, cq, q = images[img]
        oy, oc = O.transform_int16(yq, cq, q, plans[k], filters, out_size=32)
        assert oy.shape == (1, 32, 32, 8, 8) and oc.shape == (2, 16, 16, 8, 8)
        my, fy = lsb_report(oy.numpy(), g[f"case{k}_y"])
        mc, fc = lsb_report(oc.numpy(), g[f"case{k}_c"])
        assert my <= 1 and mc <= 1, (k, my, mc)
        assert fy < 5e-3 and fc < 5e-3, (k, fy, fc)


def test_swin_plan_sampler_crops():
    """RandomResizedCrop_DCT(32) on 64 x 64 blocks only ever yields crop sides 16 / 32 / 64 (x2 up / identity / x2 down)."""
    torch.manual_seed(7)
    bank = P.FilterBank()
    sides = {P.sample_train_plan(64, 64, list(P.AUGLIST_VITS), 2, 9, bank, size=32).crop_size for _ in range(300)}
    assert sides <= {16, 32, 64} and len(sides) == 3
    pl = P.eval_plan_swin(64, 64)
    assert (pl.crop_i, pl.crop_j, pl.crop_size) == (0, 0, 64)
    P.pack_plans([pl], out_size=32)
