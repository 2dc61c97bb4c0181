"""Opt-in B200 backend behind the reference's own seams (boundary B4: `RGBNM_BACKEND=b200`).

`install(utils_module)` rebinds three functions of the reference's `utils.pipeline_utils` -- the module object its
`train.py` / `eval.py` look them up on at call time:

    get_model      (pipeline_utils.py:325-373)   -> rgb_no_more_b200.vit.ViT / swin.SwinTransformerV2, same kwargs
    get_dataset    (pipeline_utils.py:260-284)   -> compat.loader.dataset_selector (B200Loader: decode threads + fused K0)
    unpack_data    (pipeline_utils.py:52-76)     -> labels to the device + RandomMixup_DCT on the device batch

Everything else of the reference (CLI, config, optimiser construction, loop, evaluation, checkpoints, tensorboard) runs as is."""
from __future__ import annotations

import os

import torch

from .. import ops as K
from . import loader as L


def requested() -> bool:
    return os.environ.get("RGBNM_BACKEND", "").lower() == "b200"


def get_model(cfg, report=True):
    dom = str(cfg.MODEL.DOMAIN).lower()
    if dom != "dct":
        raise NotImplementedError("rgbnm: RGBNM_BACKEND=b200 covers --domain=dct only; run the RGB path with the reference backend")
    if cfg.MODEL.ARCH != "swinv2":
        from ..vit import ViT
        model = ViT(in_channels=3, patch_size=cfg.MODEL.PATCHSIZE, emb_size=cfg.MODEL.EMBEDSIZE, depth=cfg.MODEL.DEPTH, n_classes=1000,
                    drop_p=cfg.TRAIN.DROP, device=cfg.RANK, dtype=torch.float32, num_heads=cfg.MODEL.HEADS, head_size=cfg.MODEL.HEADSIZE,
                    pixel_space=cfg.MODEL.DOMAIN, ver=cfg.MODEL.VERSION, use_subblock=cfg.MODEL.SUBBLOCK)
    else:
        from ..swin import SwinTransformerV2
        model = SwinTransformerV2(img_size=256, patch_size=cfg.MODEL.PATCHSIZE, embed_dim=cfg.MODEL.EMBEDSIZE, depths=cfg.MODEL.DEPTH,
                                  num_heads=cfg.MODEL.HEADS, window_size=cfg.MODEL.WINDOWSIZE, mlp_ratio=cfg.MODEL.MLPRATIO,
                                  drop_rate=cfg.MODEL.DROP, attn_drop_rate=cfg.MODEL.DROPATTN, drop_path_rate=cfg.MODEL.DROPPATH,
                                  qkv_bias=cfg.MODEL.QKVBIAS, ape=cfg.MODEL.APE, patch_norm=cfg.MODEL.PNORM,
                                  pretrained_window_sizes=cfg.MODEL.PRETRAINED, device=cfg.RANK, pixel_space="dct")
    model.prepare(torch.device("cuda", cfg.RANK) if isinstance(cfg.RANK, int) else cfg.RANK)   # flat buffers BEFORE DDP wraps it
    if cfg.RANK == 0 and report:
        n = sum(p.numel() for p in model.parameters())
        # print, not logging: the reference configures logging in the parent only (train.py:236), spawned ranks log nothing below WARNING
        print(f"rgbnm B200 backend: {type(model).__name__} ({cfg.MODEL.ARCH}), {n:,} parameters, dataset {cfg.TRAIN.DATASET}", flush=True)
    return model


def get_dataset(cfg, temp_datapath, indexpaths):
    kw = dict(dataset=cfg.TRAIN.DATASET, basepath=temp_datapath, batch_size=cfg.TRAIN.BATCHPERGPU, num_workers=cfg.THREADS,
              distributed=True, rank=cfg.RANK, world_size=cfg.WORLDSIZE, seed=cfg.SEED, device=cfg.RANK,
              subblock=bool(cfg.MODEL.SUBBLOCK))
    trainloader = valloader = trainvalloader = None
    if cfg.TRAIN.SPLIT > 0:
        trainloader, valloader, trainvalloader = L.dataset_selector(
            type="train", indexpath=indexpaths[0], shuffle=True, trainval_split=cfg.TRAIN.SPLIT, return_indices=False,
            ops_list=cfg.TRAIN.AUGLIST, num_ops=cfg.TRAIN.NUMOPS, ops_magnitude=cfg.TRAIN.AUGSTR, **kw)
    testloader = L.dataset_selector(type="test", indexpath=indexpaths[1], shuffle=False, **kw)
    return trainloader, valloader, trainvalloader, testloader


def unpack_data(data, dataset, avail_device, mixup=None, nomixup=False):
    inputs, labels = data
    if not isinstance(inputs, L.DCTBatch):
        raise TypeError("rgbnm: the B200 backend expects batches from compat.loader.B200Loader")
    x = inputs.x
    labels = labels.to(x.device, non_blocking=True)
    if mixup and not nomixup:
        # RandomMixup_DCT (cls_transforms.py:135-182) on the embed input: mixup is a convex combination and the embed tail is
        # affine, so it commutes with ToRange / sub-block conversion; same Dirichlet draw from the global CPU generator
        target = torch.nn.functional.one_hot(labels, num_classes=mixup.num_classes).to(torch.float32)
        lam, _ = torch._sample_dirichlet(torch.tensor([mixup.alpha, mixup.alpha])).sort(descending=True)
        lam_dev = lam.to(device=x.device, dtype=torch.float32)
        out = torch.empty_like(x)
        K.mixup(x.contiguous(), out, lam_dev)
        labels = target * lam_dev[0] + target.roll(1, 0) * lam_dev[1]
        x = out
    return (x, L.NO_CHROMA), labels


class B200Mixup:
    """`utils.get_mixup(cfg)` (pipeline_utils.py:169-181) for the B200 backend: the reference's RandomMixup_DCT, which additionally
    accepts the `(embed input, NoChroma)` pair of a DCTBatch -- the reference's benchmark loop calls `mixup((y, cbcr), labels)` itself
    (benchmark.py:333-334) instead of going through `unpack_data`."""

    def __init__(self, reference_mixup):
        self.ref = reference_mixup
        self.num_classes, self.alpha = reference_mixup.num_classes, reference_mixup.alpha

    def __call__(self, batch, target):
        y, cbcr = batch
        if not getattr(cbcr, "_rgbnm_absent", False):
            return self.ref(batch, target)                        # reference-format planes: the reference's own transform
        (x, _), labels = unpack_data((L.DCTBatch(y), target), None, None, mixup=self)
        return (x, cbcr), labels


def install(utils_module) -> None:
    ref_get_mixup = utils_module.get_mixup
    utils_module.get_model = get_model
    utils_module.get_dataset = get_dataset
    utils_module.unpack_data = unpack_data
    utils_module.get_mixup = lambda cfg: B200Mixup(ref_get_mixup(cfg))
    utils_module._rgbnm_backend = "b200"
