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


def _swin_shell():
    from rgb_no_more_b200 import swin as S
    return S.SwinTransformerV2(img_size=256, patch_size=4, embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24],
                               window_size=8, mlp_ratio=4, drop_rate=0, attn_drop_rate=0, drop_path_rate=0.2, qkv_bias=True,
                               ape=False, patch_norm=True, pretrained_window_sizes=[0, 0, 0, 0], device="cpu", pixel_space="dct")


def test_swin_state_dict_matches_reference_keys():
    """The drop-in exposes exactly the reference's state_dict (names, shapes, buffers incl. attn_mask values)."""
    g = load("swin_model.npz")
    m = _swin_shell()
    sd = m.state_dict()
    assert sorted(sd.keys()) == list(g["state_keys"])
    for k, shp in zip(g["state_keys"], g["state_shapes"]):
        assert ",".join(map(str, sd[str(k)].shape)) == str(shp), k
    from oracle import swin_oracle as SO
    assert torch.equal(sd["layers.0.blocks.1.attn_mask"], SO.shift_mask(64, 64, 8, 4))
    t, i = SO.relative_tables(8)
    assert torch.equal(sd["layers.2.blocks.3.attn.relative_position_index"], i)
    assert torch.allclose(sd["layers.2.blocks.3.attn.relative_coords_table"], t)


def test_swin_oracle_matches_reference():
    """oracle/swin_oracle.py against the real swinv2.SwinTransformerV2 outputs in swin_model.npz (fp32 CPU both sides)."""
    from oracle import swin_oracle as SO
    from tests.helpers import seeded_swin_state_dict
    g = load("swin_model.npz")
    sd = seeded_swin_state_dict(_swin_shell())
    yf, cf = golden_swin_inputs(g["input_seed"])
    emb = O.embed_input_swin(yf, cf)
    assert np.abs(SO.tokens(sd, emb)[0].numpy() - g["tokens"]).max() < 2e-5
    acts = []
    logits = SO.forward_from_embed(sd, emb, collect=acts)
    acts = dict(acts)
    for name, rows in (("l0b0", 512), ("l0b1", 512), ("stage0", 256), ("stage1", 256), ("stage2", 256), ("stage3", 256)):
        ref = g["act:" + name]
        got = acts[name][0, :rows].numpy()
        assert np.abs(got - ref).max() < 2e-4 * max(1.0, np.abs(ref).max()), name
    assert np.abs(logits.numpy() - g["logits"]).max() < 2e-4


def test_swin_no_cpu_fallback():
    m = _swin_shell().eval()
    yf, cf = golden_swin_inputs(5, batch=1)
    import pytest
    from rgb_no_more_b200.lib import RgbnmError
    with pytest.raises((RgbnmError, RuntimeError)):
        with torch.no_grad():
            m(yf, cf)
    m.train()
    with pytest.raises((RgbnmError, RuntimeError)):       # the training engine needs a CUDA device as well
        m(yf, cf)


def test_swin_compat_embed_path_matches_oracle():
    """forward(y, cbcr)'s torch-op tail (reference-format inputs) equals the oracle's embed input."""
    from rgb_no_more_b200 import swin as S
    yf, cf = golden_swin_inputs(9, batch=1)
    got = S.swin_embed_input_from_planes(yf, cf)
    ref = O.embed_input_swin(yf, cf).reshape(1, 4096, 24)
    assert float((got - ref).abs().max()) < 1e-5


def test_swin_oracle_training_gradients_match_reference():
    """Groundwork for the SwinV2 backward: autograd through oracle/swin_oracle.py reproduces the loss and
    the gradients of the reference model in train() mode (stochastic depth off), incl. the logit scale, the cpb_mlp behind
    the position bias, q / v biases, post-norm weights and the patch-merging reduction."""
    from oracle import swin_oracle as SO
    from tests.helpers import seeded_swin_state_dict
    g = load("swin_train.npz")
    sd = seeded_swin_state_dict(_swin_shell())
    yf, cf = golden_swin_inputs(g["input_seed"])
    params = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "relative_coords_table" not in k and "attn_mask" not in k
                  else v.clone()) for k, v in sd.items()}
    labels = torch.zeros((2, 1000))
    labels[0, 3], labels[0, 7], labels[1, 999] = 0.7, 0.3, 1.0
    logits = SO.forward(params, yf, cf)
    loss = torch.nn.CrossEntropyLoss()(logits, labels)
    loss.backward()
    assert abs(float(loss.detach()) - float(g["loss"])) < 1e-4
    assert np.abs(logits.detach().numpy() - g["logits"]).max() < 2e-4
    keys = [k[5:] for k in g.files if k.startswith("grad:")]
    assert len(keys) == 10
    for k in keys:
        got = params[k].grad.reshape(-1)[:4096].numpy()
        ref = g["grad:" + k]
        assert np.abs(got - ref).max() < 1e-4 * max(1.0, np.abs(ref).max()) + 2e-6, (k, np.abs(got - ref).max())
        assert abs(float(params[k].grad.norm()) - float(g["gradnorm:" + k])) < 1e-3 * max(1e-3, float(g["gradnorm:" + k])), k
