"""GPU parity tests for the SwinV2 layout of the fused DCT kernel (datasets.py:370-382 +
swinv2.PatchEmbedding_DCT_Group, models/swinv2.py:505-576), called through the C-ABI (rgbnm_k0_*_ex, layout SWIN4).

Same stage-wise statement as tests/test_k0_gpu.py: (1) dequantise + crop + resize to 32 x 32 blocks: bit-exact without
a resize (crop 32), one LSB only on exact .5 ties for x2 up / x2 down; (2) flip + RandAugment ops bit-exact given the
resized planes; (3) ToRange + block decomposition + interleaved rearrange (fp32): |diff| <= F32_TOL against the oracle
fed with K0's own int16 planes."""
import numpy as np
import pytest
import torch

from oracle import dct_oracle as O
from rgb_no_more_b200 import dct_manip as dm
from rgb_no_more_b200 import plan as P
from rgb_no_more_b200 import synth
from rgb_no_more_b200 import transforms as TF
from tests.helpers import load, unpack_plans, lsb_report, assert_only_tie_mismatches

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
F32_TOL = 2e-5
BF16_TOL = 2 ** -8


def _run_planes(tf, y, c, q, plans):
    out = tf.run(y.to(DEV), c.to(DEV), q.to(DEV), plans, out_mode=TF.OUT_INT16_PLANES)
    torch.cuda.synchronize()
    return TF.split_planes(out.cpu(), 32)


def _views(y, c, q, b):
    return y[b].reshape(1, 64, 64, 8, 8), c[b].reshape(2, 32, 32, 8, 8), q[b].reshape(3, 8, 8)


def _resize_only(pl):
    return P.Plan(crop_i=pl.crop_i, crop_j=pl.crop_j, crop_size=pl.crop_size, flip=False, train=False, ops=[])


def _check_stagewise(tf, y, c, q, plans):
    oy, oc = _run_planes(tf, y, c, q, plans)
    ry, rc = _run_planes(tf, y, c, q, [_resize_only(p) for p in plans])
    fracs = []
    for b, pl in enumerate(plans):
        yq, cq, qq = _views(y, c, q, b)
        ey, ec = O.resized_planes(yq, cq, qq, pl, out_size=32)
        desc = (b, pl.crop_size, [o.name for o in pl.ops])
        if pl.crop_size == 32:
            assert torch.equal(ry[b], ey) and torch.equal(rc[b], ec), desc
        else:
            xy, xc = O.resized_planes_exact(yq, cq, qq, pl, out_size=32)
            fracs.append(assert_only_tie_mismatches(ry[b].numpy(), ey.numpy(), xy.numpy(), desc))
            fracs.append(assert_only_tie_mismatches(rc[b].numpy(), ec.numpy(), xc.numpy(), desc))
        fy, fc = O.transform_from_resized(ry[b].clone(), rc[b].clone(), pl, tf.bank.table)
        assert torch.equal(oy[b], fy), (desc, lsb_report(oy[b].numpy(), fy.numpy()))
        assert torch.equal(oc[b], fc), (desc, lsb_report(oc[b].numpy(), fc.numpy()))
    return fracs


def test_golden_swin_pipeline_cases():
    """Cases produced by the reference's own 'imagenet_dct_swin' transform classes (tools/make_golden_swin.py)."""
    g = load("swin_pipeline.npz")
    plans = unpack_plans(g["plans"])
    tf = TF.FusedDCT(DEV, "train", P.AUGLIST_VITS, 2, 9, out_size=32)
    tf.bank.table[:] = g["filters"]
    tf.bank._n = 47
    tf._filters_n = -1
    images = []
    for i in g["synth_ids"]:
        dims, quant, Y, C = dm.read_coefficients_from_bytes(synth.synth_jpeg(int(i)))
        images.append((Y.reshape(1, 64, 64, 64).contiguous(), C.reshape(1, 2, 32, 32, 64).contiguous(), quant.reshape(1, 3, 64).contiguous()))
    for k, (img, seed, mag) in enumerate(g["cases"]):
        y, c, q = images[img]
        _check_stagewise(tf, y, c, q, [plans[k]])
        oy, oc = _run_planes(tf, y, c, q, [plans[k]])
        if plans[k].crop_size == 32:
            assert np.array_equal(oy[0].numpy(), g[f"case{k}_y"]) and np.array_equal(oc[0].numpy(), g[f"case{k}_c"]), k
        else:
            assert float((oy[0].numpy() != g[f"case{k}_y"]).mean()) < 0.12, k
            assert float((oc[0].numpy() != g[f"case{k}_c"]).mean()) < 0.12, k


def _random_batch(B, seed, dense):
    y, c, q = synth.synth_coefficients(B, 64, 64, seed=seed, dense=dense)
    return torch.from_numpy(y), torch.from_numpy(c), torch.from_numpy(q)


@pytest.mark.parametrize("dense", [False, True])
@pytest.mark.parametrize("mag,ops", [(9, P.AUGLIST_VITS), (3, P.AUGLIST_VITTI)])
def test_swin_random_plans_vs_oracle(dense, mag, ops):
    B = 24
    y, c, q = _random_batch(B, 17 + mag, dense)
    tf = TF.FusedDCT(DEV, "train", ops, 2, mag, out_size=32)
    torch.manual_seed(4321 + mag + int(dense))
    plans = tf.sample_plans(B)
    assert len({p.crop_size for p in plans}) >= 2 and {p.crop_size for p in plans} <= {16, 32, 64}
    fracs = _check_stagewise(tf, y, c, q, plans)
    assert fracs and max(fracs) < 0.15


@pytest.mark.parametrize("crop", [16, 32, 64])
def test_swin_embed_input_f32_and_bf16(crop):
    B = 6
    y, c, q = _random_batch(B, 31 + crop, False)
    plans = [P.Plan(crop_i=0 if crop == 64 else 2 * (b % 3), crop_j=0 if crop == 64 else 4, crop_size=crop, flip=bool(b & 1),
                    train=False, ops=[]) for b in range(B)]
    tf = TF.FusedDCT(DEV, "test", out_size=32)
    out = tf.run(y.to(DEV), c.to(DEV), q.to(DEV), plans, out_mode=TF.OUT_F32).cpu()
    outb = tf.run(y.to(DEV), c.to(DEV), q.to(DEV), plans, out_mode=TF.OUT_BF16).float().cpu()
    assert out.shape == (B, 4096, 24)
    py, pc = _run_planes(tf, y, c, q, plans)
    for b in range(B):
        ref = O.embed_input_swin(O.to_range(py[b]).unsqueeze(0), O.to_range(pc[b]).unsqueeze(0)).reshape(4096, 24)
        d = (out[b] - ref).abs()
        assert float(d.max()) < F32_TOL, (float(d.max()), int(d.argmax()))
        assert float((outb[b] - ref).abs().max()) < BF16_TOL + F32_TOL


def test_swin_index_permutation_exact():
    """The interleaved sub-block read-out is an index permutation: with one non-zero coefficient per block whose
    decomposition is known in closed form (DC only: D = dc/4 * ones for luma 2x2 sub-blocks' DC ...), every token
    receives exactly the value the oracle puts there -- checked on a delta image, position by position."""
    y = torch.zeros((1, 64, 64, 64), dtype=torch.int16)
    c = torch.zeros((1, 2, 32, 32, 64), dtype=torch.int16)
    rng = np.random.default_rng(3)
    y[0, :32, :32, :] = torch.from_numpy(rng.integers(-1024, 1017, (32, 32, 64)).astype(np.int16))
    c[0, :, :16, :16, :] = torch.from_numpy(rng.integers(-1024, 1017, (2, 16, 16, 64)).astype(np.int16))
    q = torch.ones((1, 3, 64), dtype=torch.int16)
    tf = TF.FusedDCT(DEV, "test", out_size=32)
    out = tf.run(y.to(DEV), c.to(DEV), q.to(DEV), [P.Plan(0, 0, 32)], out_mode=TF.OUT_F32).cpu()[0]
    yy = y[0, :32, :32].reshape(1, 32, 32, 8, 8)
    cc = c[0, :, :16, :16].reshape(2, 16, 16, 8, 8)
    ref = O.embed_input_swin(O.to_range(yy).unsqueeze(0), O.to_range(cc).unsqueeze(0)).reshape(4096, 24)
    assert float((out - ref).abs().max()) < F32_TOL
    # a wrong permutation would show up as O(1) errors; make that explicit
    assert float((out - ref.roll(1, 0)).abs().max()) > 0.1


def test_swin_properties_full_batch():
    """Size-independent properties at the benchmark batch size (B = 256)."""
    B = 256
    y, c, q = _random_batch(B, 98, True)
    yd, cd, qd = y.to(DEV), c.to(DEV), q.to(DEV)
    tf = TF.FusedDCT(DEV, "train", P.AUGLIST_VITS, 4, 9, out_size=32)
    op = lambda name, p=None, f=0.0: P.PlanOp(code=P.OP_NAMES[name], p=(p or [0] * 8), f=f, name=name)
    base = [P.Plan(4, 8, 32, False, True, []) for _ in range(B)]
    ref = tf.run(yd, cd, qd, base, out_mode=TF.OUT_INT16_PLANES)
    rot4 = [P.Plan(4, 8, 32, False, True, [op("Rotate90", [1] + [0] * 7)] * 4) for _ in range(B)]
    got = tf.run(yd, cd, qd, rot4, out_mode=TF.OUT_INT16_PLANES)
    assert torch.equal(got.clamp(min=-1016), ref.clamp(min=-1016))
    fl = [P.Plan(4, 8, 32, True, True, []) for _ in range(B)]
    fy, fc = TF.split_planes(tf.run(yd, cd, qd, fl, out_mode=TF.OUT_INT16_PLANES), 32)
    ry, rc = TF.split_planes(ref, 32)
    sign = torch.tensor([1, -1] * 4, dtype=torch.int16, device=DEV)
    assert torch.equal(fy, (ry.flip(3) * sign).clamp(-1024, 1016))
    assert torch.equal(fc, (rc.flip(3) * sign).clamp(-1024, 1016))
    # the decomposition is orthonormal: per image the embed input keeps the energy of the ToRange'd planes
    e = tf.run(yd, cd, qd, base, out_mode=TF.OUT_F32)
    pf = (ref.float() + 1024) / 2040 * 2 - 1
    assert torch.allclose(e.square().sum(dim=(1, 2)), pf.square().sum(dim=1), rtol=1e-4)


def test_swin_bad_arguments_fail_loudly():
    tf = TF.FusedDCT(DEV, "test", out_size=32)
    y, c, q = _random_batch(2, 1, False)
    with pytest.raises(ValueError):
        tf.run(y.to(DEV), c.to(DEV), q.to(DEV), [P.Plan(0, 0, 28)] * 2)           # a ViT crop side in the Swin geometry
    with pytest.raises(ValueError):
        tf.run(y.to(DEV), c.to(DEV), q.to(DEV), [P.Plan(40, 0, 32)] * 2)          # crop outside the image
