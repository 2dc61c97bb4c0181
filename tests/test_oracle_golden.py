"""Pin the CPU oracle (oracle/dct_oracle.py) against vectors produced by the reference
itself (tools/make_golden.py).  CPU-only."""
import numpy as np
import pytest
import torch

from oracle import dct_oracle as O
from rgb_no_more_b200 import plan as P
from tests.helpers import load, unpack_plans, lsb_report

torch.set_num_threads(1)


def test_conversion_matrices_bit_equal():
    g = load("resize.npz")
    assert np.array_equal(O.conversion_matrix(2).numpy(), g["A16"])
    assert np.array_equal(O.conversion_matrix(7).numpy(), g["A7"])
    A = O.conversion_matrix(2)
    assert float((A @ A.T - torch.eye(16)).abs().max()) < 2e-6          # orthonormal
    # even rows are 2-sparse up to fp32 noise: A16[2m] = (e_m | (-1)^m e_m)/sqrt(2)
    for m in range(8):
        row = A[2 * m].clone()
        assert abs(float(row[m]) - 2 ** -0.5) < 1e-6 and abs(float(row[8 + m]) - (-1) ** m * 2 ** -0.5) < 1e-6
        row[m] = 0
        row[8 + m] = 0
        assert float(row.abs().max()) < 3e-6


@pytest.mark.parametrize("side", [2, 4, 14, 28, 56])
def test_resize_matches_reference(side):
    g = load("resize.npz")
    out = O.resize_blocks(torch.from_numpy(g[f"in_{side}"]), 28).numpy()
    mx, frac = lsb_report(out, g[f"out_{side}"])
    assert mx <= 1 and frac < 2e-3, (mx, frac)
    outc = O.resize_blocks(torch.from_numpy(g[f"cin_{side}"]), 14).numpy()
    mx, frac = lsb_report(outc, g[f"cout_{side}"])
    assert mx <= 1 and frac < 2e-3, (mx, frac)


def _resolve(name, mag, seed, bank):
    if seed is not None:
        torch.manual_seed(seed)
    return P.resolve_op(name, mag, 8, bank)


def test_every_op_bit_exact_on_small_grid():
    g = load("ops_small.npz")
    y, c = torch.from_numpy(g["y"]), torch.from_numpy(g["c"])
    bank = P.FilterBank()
    for k, (name, mag) in enumerate(zip(g["case_names"], g["case_mags"])):
        name = str(name)
        key = f"{name}_{k}"
        seed = int(g[key + "_seed"]) if key + "_seed" in g.files else None
        op = _resolve(name, float(mag), seed, bank)
        oy, oc = O.apply_op(y.clone(), c.clone(), op, bank.table)
        assert np.array_equal(oy.numpy(), g[key + "_y"]), (name, mag)
        assert np.array_equal(oc.numpy(), g[key + "_c"]), (name, mag)


def test_full_pipeline_matches_reference():
    g = load("pipeline.npz")
    plans = unpack_plans(g["plans"])
    filters = g["filters"]
    worst = 0.0
    for k, (img, seed, mag) in enumerate(g["cases"]):
        yq = torch.from_numpy(g[f"img{img}_y"])
        cq = torch.from_numpy(g[f"img{img}_c"])
        q = torch.from_numpy(g[f"img{img}_q"])
        oy, oc = O.transform_int16(yq, cq, q, plans[k], filters)
        my, fy = lsb_report(oy.numpy(), g[f"case{k}_y"])
        mc, fc = lsb_report(oc.numpy(), g[f"case{k}_c"])
        # same machine + same torch ops => identical; allow resize ties on other CPUs
        assert my <= 1 and mc <= 1, (k, my, mc)
        assert fy < 5e-3 and fc < 5e-3, (k, fy, fc)
        worst = max(worst, fy, fc)
    print("worst mismatch fraction", worst)


def test_to_range_formula():
    x = torch.arange(-1100, 1100, dtype=torch.int16)
    z = O.to_range(x)
    assert z.dtype == torch.float32
    assert float(z[x == -1024]) == -1.0 and float(z[x == 1016]) == 1.0
    ref = -1 + 2 * ((x.double() + 1024) / 2040)
    assert float((z.double() - ref).abs().max()) < 2e-7


def test_embed_input_bit_exact_permutation():
    g = load("embed_vit.npz")
    gen = torch.Generator().manual_seed(int(g["input_seed"]))
    yf = torch.rand((2, 1, 28, 28, 8, 8), generator=gen) * 2 - 1
    cf = torch.rand((2, 2, 14, 14, 8, 8), generator=gen) * 2 - 1
    e = O.embed_input(yf, cf)[0].numpy()
    ref = g["embed_in"]
    # chroma part is a pure permutation -> bit exact; luma part goes through A16 X A16^T
    assert np.array_equal(e[..., 256:], ref[..., 256:])
    assert np.abs(e[..., :256] - ref[..., :256]).max() < 1e-5


def test_vit_oracle_matches_reference():
    """oracle/vit_oracle.py against the real pvit.ViT outputs in embed_vit.npz (fp32 CPU both sides)."""
    from oracle import vit_oracle as VO
    from tests.helpers import seeded_state_dict, golden_vit_inputs
    from rgb_no_more_b200 import vit as V
    g = load("embed_vit.npz")
    shell = V.ViT(patch_size=16, emb_size=192, depth=12, n_classes=1000, drop_p=0.0, num_heads=3, head_size=64,
                  pixel_space="DCT", ver=1, use_subblock=True)           # only used for the key/shape list
    sd = seeded_state_dict(shell)
    yf, cf = golden_vit_inputs(g["input_seed"])
    emb_in = O.embed_input(yf, cf)
    assert np.abs(VO.tokens(sd, emb_in)[0].numpy() - g["tokens"]).max() < 2e-5
    assert np.abs(VO.forward(sd, yf, cf, upto_block=1)[0].numpy() - g["after_block0"]).max() < 5e-5
    logits = VO.forward(sd, yf, cf)
    assert np.abs(logits.numpy() - g["logits_vitti"]).max() < 2e-4
    labels = torch.zeros((2, 1000))
    labels[0, 3], labels[0, 7], labels[1, 999] = 0.7, 0.3, 1.0
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    loss = torch.nn.CrossEntropyLoss()(VO.forward(params, yf, cf), labels)
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) < 1e-4
    for k in ("patchembed.projection.0.weight", "encoder.0.0.fn.eb_mha.qkv.weight", "encoder.11.1.fn.eb_ffb.3.bias",
              "classhead.ch_linear2.weight", "encoder.5.0.fn.eb_lrnorm1.weight"):
        got = params[k].grad.reshape(-1)[:4096].numpy()
        assert np.abs(got - g["grad:" + k]).max() < 1e-4 * max(1.0, np.abs(g["grad:" + k]).max()), k


def test_ops_extra_bit_exact():
    """FreqEnhance (dct_ops.py:1015-1034; dispatchable, outside the default AUGLISTs) against the reference's own output."""
    g = load("ops_extra.npz")
    y, c = torch.from_numpy(g["y"]), torch.from_numpy(g["c"])
    bank = P.FilterBank()
    for k, (name, mag) in enumerate(zip(g["case_names"], g["case_mags"])):
        op = P.resolve_op(str(name), float(mag), 8, bank)
        oy, oc = O.apply_op(y.clone(), c.clone(), op, bank.table)
        assert np.array_equal(oy.numpy(), g[f"{name}_{k}_y"]), (name, mag)
        assert np.array_equal(oc.numpy(), g[f"{name}_{k}_c"]), (name, mag)


def test_equalize_quantised_dc_plane():
    g = load("ops_extra.npz")
    y2, c = torch.from_numpy(g["y2"]), torch.from_numpy(g["c"])
    op = P.resolve_op("Equalize", 0.0, 8, P.FilterBank())
    oy, oc = O.apply_op(y2.clone(), c.clone(), op, None)
    assert np.array_equal(oy.numpy(), g["Equalize_y2_y"]) and np.array_equal(oc.numpy(), g["Equalize_y2_c"])
    # one distinct DC value: the reference divides by zero; here the plane is left unchanged
    flat = y2.clone()
    flat[0, :, :, 0, 0] = 37
    assert torch.equal(O.equalize(flat), flat)


def test_vit_oracle_train_step_matches_reference_loop():
    """oracle/vit_oracle.py (forward, loss, gradients AND the 3-step optimiser loop: clip -> AdamW -> decoupled decay under
    the warm-up schedule) against the reference's own classes run by tools/make_golden.py::gen_vit_s (ViT-S, batch 8)."""
    from oracle import vit_oracle as VO
    from tests.helpers import seeded_state_dict, golden_vits_inputs, vits_soft_labels, VITS_GOLDEN_KEYS
    from rgb_no_more_b200 import vit as V
    torch.set_num_threads(8)
    g = load("vit_s.npz")
    shell = V.ViT(patch_size=16, emb_size=384, depth=12, n_classes=1000, drop_p=0.0, num_heads=6, head_size=64,
                  pixel_space="DCT", ver=1, use_subblock=True)
    sd = seeded_state_dict(shell)
    yf, cf = golden_vits_inputs(g["input_seed"])
    emb_in = O.embed_input(yf, cf)
    with torch.no_grad():
        logits = VO.forward_embedded(sd, emb_in)
    assert np.abs(logits.numpy() - g["logits"]).max() < 2e-4 * max(1.0, np.abs(g["logits"]).max())
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    loss = torch.nn.CrossEntropyLoss()(VO.forward_embedded(params, emb_in), vits_soft_labels())
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) < 1e-4
    for k in VITS_GOLDEN_KEYS:
        got = params[k].grad.reshape(-1)[:4096].numpy()
        assert np.abs(got - g["grad:" + k]).max() < 2e-4 * max(1e-3, np.abs(g["grad:" + k]).max()), k
    # the optimiser loop: same lr trace as train.py:149-152 (it = step + 1 -> LR * (it + 1) / WARMUP)
    onehot = torch.nn.functional.one_hot(torch.from_numpy(g["train_labels"]), 1000).float()
    state, losses = {}, []
    sd = {k: v.clone() for k, v in sd.items()}
    for s_, l0 in enumerate(g["train_lams"]):
        l0 = float(l0)
        x = l0 * emb_in + (1 - l0) * emb_in.roll(1, 0)
        soft = l0 * onehot + (1 - l0) * onehot.roll(1, 0)
        losses.append(VO.train_step(sd, x, soft, state, lr=3e-3 * (s_ + 2) / 10, wd=3e-4))
    assert np.abs(np.array(losses) - g["train_losses"]).max() < 2e-3, (losses, g["train_losses"])
    init = seeded_state_dict(shell)
    for k in VITS_GOLDEN_KEYS:
        ref = g["param3:" + k]
        d_ref = ref - init[k].reshape(-1)[:4096].numpy()
        d_got = sd[k].reshape(-1)[:4096].numpy() - init[k].reshape(-1)[:4096].numpy()
        cos = float((d_ref * d_got).sum() / (np.linalg.norm(d_ref) * np.linalg.norm(d_got) + 1e-30))
        assert cos > 0.999, (k, cos)         # fp32 both sides; Adam's m / sqrt(v) amplifies last-bit gradient differences near 0


@pytest.mark.parametrize("ver,sub", [(2, True), (2, False), (1, False)])
def test_vit_oracle_embedding_variants_match_reference(ver, sub):
    """embed_type 2 (with / without sub-block conversion) and embed_type 1 --no_subblock: oracle tokens / logits / loss / embedding
    gradients vs the reference's own pvit.ViT (tests/golden/embed_variants.npz, tools/make_golden.py::gen_embed_variants)."""
    from oracle import vit_oracle as VO
    from tests.helpers import seeded_state_dict, golden_vits_inputs
    from rgb_no_more_b200 import vit as V
    g = load("embed_variants.npz")
    tag = f"v{ver}{'s' if sub else 'n'}"
    shell = V.ViT(patch_size=16, emb_size=192, depth=2, n_classes=1000, drop_p=0.0, num_heads=3, head_size=64, pixel_space="DCT",
                  ver=ver, use_subblock=sub)
    assert sorted(shell.state_dict().keys()) == list(g[f"{tag}:state_keys"])
    # through load_state_dict, like the generator: PatchEmbedding_DCT_Separate registers its mixing Linear twice (LinearMix and
    # projection.1 are one tensor), so the two seeded entries collapse to whichever is loaded last -- in both module trees
    shell.load_state_dict(seeded_state_dict(shell))
    sd = {k: v.detach().clone() for k, v in shell.state_dict().items()}
    yf, cf = golden_vits_inputs(g["input_seed"], batch=3)
    emb_in = O.embed_input(yf, cf, subblock=sub)
    assert np.abs(VO.tokens(sd, emb_in)[0].numpy() - g[f"{tag}:tokens"]).max() < 5e-5
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    logits = VO.forward_embedded(params, emb_in, depth=2)
    assert np.abs(logits.detach().numpy() - g[f"{tag}:logits"]).max() < 2e-4
    labels = torch.zeros((3, 1000))
    labels[0, 3], labels[0, 7], labels[1, 999], labels[2, 500] = 0.7, 0.3, 1.0, 1.0
    loss = torch.nn.CrossEntropyLoss()(logits, labels)
    loss.backward()
    assert abs(float(loss.detach()) - float(g[f"{tag}:loss"])) < 1e-4
    for key in g.files:
        if key.startswith(f"{tag}:grad:"):
            k = key.split(":", 2)[2]
            ref = g[key]
            got = params[k].grad.reshape(-1)[:ref.size].numpy()
            assert np.abs(got - ref).max() < 2e-4 * max(1e-3, np.abs(ref).max()), k
