"""Generate tests/golden/*.npz by running the REFERENCE ITSELF (imported from
/root/reference) in the build container.  The reference cannot travel to the GPU box, so
the vectors are committed; this script is the committed recipe that made them.

    python tools/make_golden.py            # writes tests/golden/

Vectors:
  ops_small.npz      every supported _apply_op_dct op on a small grid  (utils/custom_transforms.py:944)
  resize.npz         resize_dct for the five crop cases                 (utils/dct_ops.py:529)
  pipeline.npz       full get_transform('imagenet_dct', train/test) runs under fixed seeds,
                     with the plans our sampler resolves under the same seeds
  embed_vit.npz      PatchEmbedding_DCT_Group input tensor + ViT-Ti logits  (models/plainvit.py)
  vit_s.npz          ViT-S (benchmarked config) logits / loss / 15 gradients at batch 8 and 3 steps of the reference's optimiser loop
"""
from __future__ import annotations

import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(1, REF)
if os.environ.get("PYTHONHASHSEED") != "0":
    # RandAugment_dct's list(set(...)) op order depends on the interpreter's string-hash seed (custom_transforms.py:1115-1119):
    # the vectors are made under PYTHONHASHSEED=0, which only takes effect at interpreter start -> re-exec
    os.environ["PYTHONHASHSEED"] = "0"
    os.execv(sys.executable, [sys.executable] + sys.argv)

import numpy as np  # noqa: E402
import torch  # noqa: E402

# `utils.custom_transforms` imports the native `dct_manip` module at import time only.
sys.modules.setdefault("dct_manip", types.ModuleType("dct_manip"))
import utils.custom_transforms as ctrans  # noqa: E402
import utils.dct_ops as dops  # noqa: E402
import models.plainvit as pvit  # noqa: E402

from rgb_no_more_b200 import plan as P  # noqa: E402
from rgb_no_more_b200 import synth  # noqa: E402
from rgb_no_more_b200 import dct_manip as dm  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
torch.set_num_threads(1)


def ref_transform(kind, ops_list, num_ops, mag):
    import torchvision.transforms as T
    if kind == "train":
        return T.Compose([
            ctrans.RandomResizedCrop_DCT(28, scale=(0.05, 1.0), ratio=(1, 1)),
            ctrans.RandomFlip_DCT(p=0.5, direction="horizontal"),
            ctrans.RandAugment_dct(num_ops=num_ops, magnitude=mag, num_magnitude_bins=11, ops_list=ops_list),
        ])  # ToRange is checked separately (it is a pure elementwise map)
    return T.Compose([ctrans.ResizedCenterCrop_DCT(32, 28)])


def dequant(y, c, q):
    y = torch.clamp(y * q[0], min=-2 ** 10, max=2 ** 10 - 8)
    c = torch.clamp(c * q[1:3].unsqueeze(1).unsqueeze(1), min=-2 ** 10, max=2 ** 10 - 8)
    return y, c


def plan_to_arrays(plans):
    return P.pack_plans(plans)


def gen_ops_small():
    g = torch.Generator().manual_seed(1)
    y = torch.randint(-1024, 1017, (1, 8, 8, 8, 8), generator=g, dtype=torch.int16)
    c = torch.randint(-1024, 1017, (2, 4, 4, 8, 8), generator=g, dtype=torch.int16)
    y[0, :, :, 0, 0] = torch.randint(-900, 900, (8, 8), generator=g, dtype=torch.int16)
    out = {"y": y.numpy(), "c": c.numpy()}
    cases = []
    mags = {"TranslateX": [4.0, -4.0, 2.5, -2.5], "TranslateY": [4.0, -4.0, 1.9, -3.7],
            "Rotate90": [1.0, -1.0], "Brightness": [0.81, -0.81, 0.27], "Contrast": [0.81, -0.81],
            "Color": [0.81, -0.27], "AutoContrast": [0.0], "AutoSaturation": [0.0], "Posterize": [0, 1, 2, 3, 4, 5],
            "Sharpness": [0.81, -0.81, 0.27, -0.27], "MidfreqAug": [0.81, -0.81, 0.27, -0.27],
            "Grayscale": [0.0], "SolarizeAdd": [794.0, 264.0], "Invert": [0.0], "Identity": [0.0]}
    lin = torch.linspace(0.0, 0.9, 11)
    for name, ms in mags.items():
        for m in ms:
            if name in ("Brightness", "Contrast", "Color", "Sharpness", "MidfreqAug"):
                # use the exact fp32 bin values the reference would use
                m = float(lin[9].item()) * (1 if m > 0 else -1) if abs(m) > 0.5 else float(lin[3].item()) * (1 if m > 0 else -1)
            coeff = [y.clone(), c.clone()]
            res = ctrans._apply_op_dct(coeff, name, float(m), pad=2 ** 0.5, conv_Ls=[None, None], conv_Ms=[None, None])
            key = f"{name}_{len(cases)}"
            out[key + "_y"] = res[0].numpy()
            out[key + "_c"] = res[1].numpy()
            cases.append((name, float(m)))
    # Cutout / ChromaDrop draw random numbers: replay with a seed
    for seed in range(6):
        for name, m in (("Cutout", 4.0 if seed % 2 else 2.0), ("ChromaDrop", 0.0)):
            torch.manual_seed(seed)
            coeff = [y.clone(), c.clone()]
            res = ctrans._apply_op_dct(coeff, name, float(m), pad=2 ** 0.5, conv_Ls=[None, None], conv_Ms=[None, None])
            key = f"{name}_{len(cases)}"
            out[key + "_y"] = res[0].numpy()
            out[key + "_c"] = res[1].numpy()
            out[key + "_seed"] = np.int64(seed)
            cases.append((name, float(m)))
    out["case_names"] = np.array([n for n, _ in cases])
    out["case_mags"] = np.array([m for _, m in cases], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "ops_small.npz"), **out)
    print("ops_small:", len(cases), "cases")


def gen_ops_extra():
    """Ops added after ops_small.npz was frozen: FreqEnhance (dct_ops.py:1015-1034), same input blocks as gen_ops_small."""
    g = torch.Generator().manual_seed(1)
    y = torch.randint(-1024, 1017, (1, 8, 8, 8, 8), generator=g, dtype=torch.int16)
    c = torch.randint(-1024, 1017, (2, 4, 4, 8, 8), generator=g, dtype=torch.int16)
    y[0, :, :, 0, 0] = torch.randint(-900, 900, (8, 8), generator=g, dtype=torch.int16)
    out = {"y": y.numpy(), "c": c.numpy()}
    lin = torch.linspace(0.0, 0.9, 11)
    cases = []
    for m in (float(lin[9]), -float(lin[9]), float(lin[3]), -float(lin[3]), float(lin[10]), -float(lin[10])):
        res = ctrans._apply_op_dct([y.clone(), c.clone()], "FreqEnhance", m, pad=2 ** 0.5, conv_Ls=[None, None], conv_Ms=[None, None])
        key = f"FreqEnhance_{len(cases)}"
        out[key + "_y"], out[key + "_c"] = res[0].numpy(), res[1].numpy()
        cases.append(("FreqEnhance", m))
    res = ctrans._apply_op_dct([y.clone(), c.clone()], "Equalize", 0.0, pad=2 ** 0.5, conv_Ls=[None, None], conv_Ms=[None, None])
    key = f"Equalize_{len(cases)}"
    out[key + "_y"], out[key + "_c"] = res[0].numpy(), res[1].numpy()
    cases.append(("Equalize", 0.0))
    for m in (654.4, 0.0, -327.2, 818.0, -818.0):
        res = ctrans._apply_op_dct([y.clone(), c.clone()], "Solarize", float(m), pad=2 ** 0.5, conv_Ls=[None, None], conv_Ms=[None, None])
        key = f"Solarize_{len(cases)}"
        out[key + "_y"], out[key + "_c"] = res[0].numpy(), res[1].numpy()
        cases.append(("Solarize", float(m)))
    # a DC plane with repeated values (ties in the histogram) and a few distinct levels only
    y2 = y.clone()
    y2[0, :, :, 0, 0] = (y2[0, :, :, 0, 0] // 200) * 200
    res = ctrans._apply_op_dct([y2.clone(), c.clone()], "Equalize", 0.0, pad=2 ** 0.5, conv_Ls=[None, None], conv_Ms=[None, None])
    out["y2"] = y2.numpy()
    out["Equalize_y2_y"], out["Equalize_y2_c"] = res[0].numpy(), res[1].numpy()
    out["case_names"] = np.array([n for n, _ in cases])
    out["case_mags"] = np.array([m for _, m in cases], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "ops_extra.npz"), **out)
    print("ops_extra:", len(cases), "cases")


def gen_resize():
    g = torch.Generator().manual_seed(2)
    out = {}
    for side in (2, 4, 14, 28, 56):
        x = torch.randint(-1024, 1017, (1, side, side, 8, 8), generator=g, dtype=torch.int16)
        # mix in JPEG-like sparse blocks
        x[:, :, : side // 2] = (x[:, :, : side // 2].float() * torch.rand((8, 8), generator=g).pow(4)).to(torch.int16)
        r = dops.resize_dct(x, 28, conv_mxs={})
        out[f"in_{side}"] = x.numpy()
        out[f"out_{side}"] = r.numpy()
        xc = torch.randint(-1024, 1017, (2, max(1, side // 2), max(1, side // 2), 8, 8), generator=g, dtype=torch.int16)
        rc = dops.resize_dct(xc, 14, conv_mxs={})
        out[f"cin_{side}"] = xc.numpy()
        out[f"cout_{side}"] = rc.numpy()
    out["A16"] = dops.generate_conversion_matrix(8, 2).numpy()
    out["A7"] = dops.generate_conversion_matrix(8, 7).numpy()
    np.savez_compressed(os.path.join(OUT, "resize.npz"), **out)
    print("resize done")


def gen_pipeline():
    jpegs = [synth.synth_jpeg(i) for i in range(2)]
    out = {}
    images = []
    for i, buf in enumerate(jpegs):
        dims, quant, Y, C = dm.read_coefficients_from_bytes(buf)
        images.append((Y, C, quant))
        out[f"img{i}_y"] = Y.numpy()
        out[f"img{i}_c"] = C.numpy()
        out[f"img{i}_q"] = quant.numpy()
    # one dense adversarial image (exercises clamps)
    g = torch.Generator().manual_seed(3)
    Yd = torch.randint(-1024, 1017, (1, 64, 64, 8, 8), generator=g, dtype=torch.int16)
    Cd = torch.randint(-1024, 1017, (2, 32, 32, 8, 8), generator=g, dtype=torch.int16)
    qd = torch.ones((3, 8, 8), dtype=torch.int16)
    qd[:, 0, 5] = 3
    images.append((Yd, Cd, qd))
    out["img2_y"], out["img2_c"], out["img2_q"] = Yd.numpy(), Cd.numpy(), qd.numpy()

    cases = []
    plans = []
    bank = P.FilterBank()
    # eval
    tf = ref_transform("test", None, 0, 0)
    for i, (Y, C, q) in enumerate(images):
        y, c = dequant(Y, C, q)
        ry, rc = tf((y, c))
        k = len(cases)
        out[f"case{k}_y"], out[f"case{k}_c"] = ry.numpy(), rc.numpy()
        cases.append((i, -1, 0))
        plans.append(P.eval_plan(64, 64))
    # train, both default op lists, magnitudes 9 and 3
    seeds = list(range(100, 118))
    for n, seed in enumerate(seeds):
        img = n % 3
        ops = P.AUGLIST_VITS if n % 2 == 0 else P.AUGLIST_VITTI
        mag = 9 if n % 4 < 2 else 3
        Y, C, q = images[img]
        y, c = dequant(Y, C, q)
        tf = ref_transform("train", list(ops), 2, mag)
        torch.manual_seed(seed)
        ry, rc = tf((y, c))
        torch.manual_seed(seed)
        pl = P.sample_train_plan(64, 64, list(ops), 2, mag, bank)
        k = len(cases)
        out[f"case{k}_y"], out[f"case{k}_c"] = ry.numpy(), rc.numpy()
        # the reference's planes right after RandomResizedCrop_DCT (first transform = first RNG draws): the GPU test feeds
        # them back through K0 in identity geometry, so flip + RandAugment are checked bit-exactly against the
        # reference's own final output in the resize cases too
        torch.manual_seed(seed)
        iy, ic = ctrans.RandomResizedCrop_DCT(28, scale=(0.05, 1.0), ratio=(1, 1))((y, c))
        out[f"case{k}_ry"], out[f"case{k}_rc"] = iy.numpy(), ic.numpy()
        cases.append((img, seed, mag))
        plans.append(pl)
        print(f"  case {k}: img {img} seed {seed} crop {pl.crop_i},{pl.crop_j},{pl.crop_size} flip {pl.flip} ops "
              f"{[(o.name, o.p[:4], round(o.f, 3)) for o in pl.ops]}")
    out["cases"] = np.array(cases, dtype=np.int64)
    out["plans"] = P.pack_plans(plans)
    out["plan_op_names"] = np.array([",".join(o.name for o in pl.ops) for pl in plans])
    out["filters"] = bank.table
    out["pythonhashseed"] = np.array(os.environ.get("PYTHONHASHSEED", ""))
    np.savez_compressed(os.path.join(OUT, "pipeline.npz"), **out)
    print("pipeline:", len(cases), "cases")


def seeded_state_dict(model, seed=11997733):
    """Weights by a construction-order independent recipe: per key, seeded normal."""
    sd = {}
    for k, v in sorted(model.state_dict().items()):
        g = torch.Generator().manual_seed(seed + sum(ord(ch) * (i + 1) for i, ch in enumerate(k)) % 1000003)
        if k.endswith("weight") and v.ndim == 2:
            sd[k] = torch.randn(v.shape, generator=g) * (0.5 / v.shape[1] ** 0.5)
        elif k.endswith("weight"):
            sd[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        else:
            sd[k] = 0.05 * torch.randn(v.shape, generator=g)
    return sd


def gen_embed_vit():
    g = torch.Generator().manual_seed(4)
    yf = torch.rand((2, 1, 28, 28, 8, 8), generator=g) * 2 - 1
    cf = torch.rand((2, 2, 14, 14, 8, 8), generator=g) * 2 - 1
    out = {"input_seed": np.int64(4)}   # yf/cf are regenerated from the seed by the tests
    model = pvit.ViT(patch_size=16, emb_size=192, depth=12, n_classes=1000, drop_p=0.0, num_heads=3, head_size=64,
                     pixel_space="DCT", ver=1, use_subblock=True)
    model.load_state_dict(seeded_state_dict(model))
    model.eval()
    pe = model.patchembed
    with torch.no_grad():
        y = pe.rearrange_Y(yf)
        y = pvit.apply_subblock(y, pe.conv_Y, combine=True)
        c = pe.rearrange_C(cf)
        emb_in = torch.cat([pe.collapser(y), pe.collapser(c)], dim=3)
        out["embed_in"] = emb_in[0].numpy()
        out["tokens"] = pe(yf, cf)[0].numpy()
        out["logits_vitti"] = model(yf, cf).numpy()
        x = model.patchembed(yf, cf)
        x = model.encoder[0](x)
        out["after_block0"] = x[0].numpy()
    out["state_keys"] = np.array(sorted(model.state_dict().keys()))
    out["state_shapes"] = np.array([",".join(map(str, model.state_dict()[k].shape)) for k in sorted(model.state_dict().keys())])
    # training-step golden: loss + a few gradients for soft labels
    model.train()
    labels = torch.zeros((2, 1000))
    labels[0, 3], labels[0, 7], labels[1, 999] = 0.7, 0.3, 1.0
    logits = model(yf, cf)
    loss = torch.nn.CrossEntropyLoss()(logits, labels)
    loss.backward()
    out["loss"] = loss.detach().numpy()
    sdg = dict(model.named_parameters())
    for k in ("patchembed.projection.0.weight", "encoder.0.0.fn.eb_mha.qkv.weight", "encoder.11.1.fn.eb_ffb.3.bias",
              "classhead.ch_linear2.weight", "encoder.5.0.fn.eb_lrnorm1.weight"):
        gk = sdg[k].grad
        out["gradnorm:" + k] = gk.norm().numpy()
        out["grad:" + k] = gk.reshape(-1)[:4096].numpy()
    np.savez_compressed(os.path.join(OUT, "embed_vit.npz"), **out)
    print("embed_vit done; loss", float(loss))


def golden_vits_inputs(seed, batch=8):
    """ToRange'd planes with a natural-ish spectrum (decaying with frequency) so activations are not white noise."""
    g = torch.Generator().manual_seed(int(seed))
    decay = torch.rand((8, 8), generator=g).pow(2) * 0.9 + 0.1
    yf = (torch.rand((batch, 1, 28, 28, 8, 8), generator=g) * 2 - 1) * decay
    cf = (torch.rand((batch, 2, 14, 14, 8, 8), generator=g) * 2 - 1) * decay
    return yf, cf


def gen_vit_s():
    """ViT-S (E = 384, 6 heads -- the BENCHMARKED configuration, BASELINE config 3) from the reference's own classes:
    (i) eval logits + training loss and gradients at batch 8 (plainvit.py:559-612),
    (ii) three optimiser steps of the reference loop body (train.py:146-176: CE on mixup soft labels, clip_grad_norm_(1),
         AdamW(wd = 0, eps = 1e-8), custom_optims.WeightDecay on '.weight' names without 'lrnorm', warm-up lr), fp32 CPU,
         with the mixed inputs / soft labels fixed -> parameters after the 3 steps (slices + norms)."""
    import utils.custom_optims as coptim
    torch.set_num_threads(8)
    yf, cf = golden_vits_inputs(5)
    out = {"input_seed": np.int64(5), "batch": np.int64(8)}
    model = pvit.ViT(patch_size=16, emb_size=384, depth=12, n_classes=1000, drop_p=0.0, num_heads=6, head_size=64,
                     pixel_space="DCT", ver=1, use_subblock=True)
    model.load_state_dict(seeded_state_dict(model))
    model.eval()
    with torch.no_grad():
        out["logits"] = model(yf, cf).numpy()
    model.train()
    labels = torch.zeros((8, 1000))
    for b in range(8):
        labels[b, (37 * b + 3) % 1000] = 0.75
        labels[b, (91 * b + 500) % 1000] = 0.25
    loss = torch.nn.CrossEntropyLoss()(model(yf, cf), labels)
    loss.backward()
    out["loss"] = loss.detach().numpy()
    keys = ("patchembed.projection.0.weight", "patchembed.projection.0.bias", "encoder.0.0.fn.eb_mha.qkv.weight",
            "encoder.0.0.fn.eb_mha.qkv.bias", "encoder.3.0.fn.eb_mha.projection.weight", "encoder.6.1.fn.eb_ffb.0.weight",
            "encoder.6.1.fn.eb_ffb.0.bias", "encoder.11.1.fn.eb_ffb.3.weight", "encoder.11.1.fn.eb_ffb.3.bias",
            "encoder.5.0.fn.eb_lrnorm1.weight", "encoder.9.1.fn.eb_lrnorm2.bias", "classhead.ch_lrnorm.weight",
            "classhead.ch_linear1.weight", "classhead.ch_linear2.weight", "classhead.ch_linear2.bias")
    sdg = dict(model.named_parameters())
    for k in keys:
        gk = sdg[k].grad
        out["gradnorm:" + k] = gk.norm().numpy()
        out["grad:" + k] = gk.reshape(-1)[:4096].numpy()
    out["grad_total_norm"] = torch.sqrt(sum((p.grad ** 2).sum() for p in model.parameters())).numpy()
    # ---- (ii) the reference's optimiser loop body, 3 steps ----------------------------------------------------
    model.load_state_dict(seeded_state_dict(model))
    LR, WD, WARMUP = 3e-3, 3e-4, 10
    optimizer = torch.optim.AdamW(model.parameters(), lr=LR, weight_decay=0, eps=1e-8)
    weight_decayer = coptim.WeightDecay([p for n, p in model.named_parameters() if (".weight" in n) and ("lrnorm" not in n)],
                                        lr=LR, weight_decay=WD)
    lam = [1.0, 0.8, 0.6]                                # step 0 without mixup, then RandomMixup_DCT's convex combination
    int_labels = torch.tensor([(37 * b + 3) % 1000 for b in range(8)])
    onehot = torch.nn.functional.one_hot(int_labels, 1000).float()
    out["train_labels"] = int_labels.numpy()
    losses, current_itr = [], 0
    for s_ in range(3):
        l0 = lam[s_]
        ym, cm = l0 * yf + (1 - l0) * yf.roll(1, 0), l0 * cf + (1 - l0) * cf.roll(1, 0)     # cls_transforms.py:176-179
        soft = l0 * onehot + (1 - l0) * onehot.roll(1, 0)
        optimizer.zero_grad()
        weight_decayer.zero_grad()
        current_itr += 1
        if current_itr < WARMUP:
            for grp in optimizer.param_groups:
                grp["lr"] = LR * (current_itr + 1) / WARMUP
            for gs, gd in zip(optimizer.param_groups, weight_decayer.param_groups):
                gd["lr"] = gs["lr"]
        ls = torch.nn.CrossEntropyLoss()(model(ym, cm), soft)
        ls.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1)
        optimizer.step()
        weight_decayer.step()
        losses.append(float(ls))
    out["train_lams"] = np.array(lam)
    out["train_losses"] = np.array(losses)
    init = seeded_state_dict(model)
    for k in keys:
        pk = sdg[k].detach()
        out["param3:" + k] = pk.reshape(-1)[:4096].numpy()
        out["delta3norm:" + k] = (pk - init[k]).norm().numpy()
    np.savez_compressed(os.path.join(OUT, "vit_s.npz"), **out)
    print("vit_s done; loss", float(loss), "train losses", losses)


EMBED_VARIANTS = ((2, True), (2, False), (1, False))          # (embed_type, use_subblock); (1, True) is embed_vit.npz / vit_s.npz


def gen_embed_variants():
    """The other patch embeddings of `--domain dct`: embed_type 2 with / without sub-block conversion
    (PatchEmbedding_DCT_Separate_subblock / _Separate, plainvit.py:220-350) and embed_type 1 with --no_subblock, from the
    reference's own pvit.ViT (E = 192, depth 2, batch 3): tokens, logits, loss and the gradients of every embedding parameter."""
    out = {"input_seed": np.int64(6)}
    yf, cf = golden_vits_inputs(6, batch=3)
    labels = torch.zeros((3, 1000))
    labels[0, 3], labels[0, 7], labels[1, 999], labels[2, 500] = 0.7, 0.3, 1.0, 1.0
    for ver, sub in EMBED_VARIANTS:
        tag = f"v{ver}{'s' if sub else 'n'}"
        model = pvit.ViT(patch_size=16, emb_size=192, depth=2, n_classes=1000, drop_p=0.0, num_heads=3, head_size=64,
                         pixel_space="DCT", ver=ver, use_subblock=sub)
        model.load_state_dict(seeded_state_dict(model))
        model.eval()
        with torch.no_grad():
            out[f"{tag}:tokens"] = model.patchembed(yf, cf)[0].numpy()
            out[f"{tag}:logits"] = model(yf, cf).numpy()
        model.train()
        loss = torch.nn.CrossEntropyLoss()(model(yf, cf), labels)
        loss.backward()
        out[f"{tag}:loss"] = loss.detach().numpy()
        for k, p_ in model.named_parameters():
            if k.startswith("patchembed") or k in ("encoder.0.0.fn.eb_mha.qkv.weight", "classhead.ch_linear2.bias"):
                out[f"{tag}:grad:{k}"] = p_.grad.reshape(-1)[:4096].numpy()
                out[f"{tag}:gradnorm:{k}"] = p_.grad.norm().numpy()
        out[f"{tag}:state_keys"] = np.array(sorted(model.state_dict().keys()))
    np.savez_compressed(os.path.join(OUT, "embed_variants.npz"), **out)
    print("embed_variants done")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    todo = sys.argv[1:] or ["ops_small", "ops_extra", "resize", "pipeline", "embed_vit", "vit_s", "embed_variants"]
    for name in todo:
        globals()["gen_" + name]()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")
