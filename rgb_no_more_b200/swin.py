"""SwinV2 DCT drop-in (boundary B3 for `--model_arch swinv2 --domain dct`).

`SwinTransformerV2` has the constructor surface and the 250 `state_dict` keys of the reference
(/root/reference/models/swinv2.py:578-705, built at utils/pipeline_utils.py:351-368), so reference checkpoints load
unchanged.  `forward(y, cbcr)` takes the reference's ToRange'd planes, `forward(x)` the (B, 4096, 24) tensor FusedDCT
writes in the Swin layout.  The forward pass runs on hand-written sm_100a kernels through the C-ABI:

    Linear(24, 96) / qkv / proj / fc1+GELU / fc2 / patch-merging reduction / head   rgbnm_gemm_bf16 (tcgen05, TMA)
    LayerNorm (+ post-norm residual)                                                 rgbnm_layernorm_res_fwd
    window attention incl. window partition / cyclic shift / reverse                 rgbnm_window_attention_fwd
    patch-merging gather, token mean                                                 rgbnm_patch_merge_gather, rgbnm_token_mean_bf16

Inference (eval / --benchmark) runs on the fused forward engine below; with gradients enabled the call goes through the
training engine of swin_train.py (explicit forward / backward incl. stochastic depth, first correct version).  There is
no CPU fallback.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import gemm as G
from . import lib as _lib

import os

_NOFUSE = os.environ.get("RGBNM_SWIN_NOFUSE") is not None      # A/B: separate LayerNorm kernels instead of the fused GEMM epilogue
# Measured (profiles/r01_swin_eval_v4_launches.txt): folding the norm into the GEMM pays where the layer is memory-bound
# (dims 96 / 192: -39 .. -93 us per block at batch 256) and costs where the 384-wide single-accumulator tile exposes the
# longer epilogue (dim 384: +5 .. +30 us) -> fused up to 192 channels.
_FUSE_MAX_DIM = int(os.environ.get("RGBNM_SWIN_FUSE_MAX_DIM", "192"))
RES = 64          # tokens per side for img_size 256 / patch 4
IN_FEAT = 24      # 4x4 luma + 2x2 Cb + 2x2 Cr


class _Box(nn.Module):
    """Parameter container mirroring one level of the reference module tree."""


def swin_embed_input_from_planes(y: torch.Tensor, cbcr: torch.Tensor) -> torch.Tensor:
    """Compatibility path for reference-format inputs (B,1,32,32,8,8) + (B,2,16,16,8,8), already ToRange'd: the tail of
    the fused kernel in the Swin layout (block decomposition + interleaved rearrange, swinv2.py:553-565) with torch ops
    on the GPU.  The fast path never comes here: FusedDCT(out_size=32) writes this tensor directly."""
    dev = y.device

    def basis(L):
        k = torch.arange(L, dtype=torch.float32, device=dev).unsqueeze(1)
        m = torch.arange(L, dtype=torch.float32, device=dev).unsqueeze(0) + 0.5
        d = (k.mm(m) * torch.pi / L).cos()
        d[0] *= 1 / (2 ** 0.5)
        return d * (2 / L) ** 0.5

    def decompose(x, small, pd):
        a = basis(8).mm(torch.block_diag(*[basis(small)] * pd).T)
        x = torch.matmul(torch.matmul(a.T, x.float()), a)
        b, c, h, w, _, _ = x.shape
        x = x.reshape(b, c, h, w, small, pd, small, pd).permute(0, 1, 2, 5, 3, 7, 4, 6)
        return x.reshape(b, c, h * pd, w * pd, small * small)

    yy = decompose(y, 4, 2)
    cc = decompose(cbcr, 2, 4)
    b, _, H, W, _ = yy.shape
    out = torch.cat([yy.permute(0, 2, 3, 1, 4).reshape(b, H, W, 16), cc.permute(0, 2, 3, 1, 4).reshape(b, H, W, 8)], dim=3)
    return out.reshape(b, H * W, IN_FEAT)


def _relative_tables(ws: int):
    # WindowAttention.__init__, swinv2.py:98-125 (pretrained_window_size = 0)
    rh = torch.arange(-(ws - 1), ws, dtype=torch.float32)
    table = torch.stack(torch.meshgrid([rh, rh], indexing="ij")).permute(1, 2, 0).contiguous().unsqueeze(0)
    table /= (ws - 1)
    table *= 8
    table = torch.sign(table) * torch.log2(torch.abs(table) + 1.0) / math.log2(8)
    coords = torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing="ij"))
    cf = torch.flatten(coords, 1)
    rel = (cf[:, :, None] - cf[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return table, rel.sum(-1)


def _shift_mask(H: int, W: int, ws: int, shift: int) -> torch.Tensor:
    # the `attn_mask` buffer of SwinTransformerBlock (swinv2.py:224-242); kept for state_dict compatibility only --
    # the attention kernel derives the same regions from the token coordinates
    img_mask = torch.zeros((1, H, W, 1))
    cnt = 0
    for h in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for w in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img_mask[:, h, w, :] = cnt
            cnt += 1
    mw = img_mask.view(1, H // ws, ws, W // ws, ws, 1).permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws * ws)
    am = mw.unsqueeze(1) - mw.unsqueeze(2)
    return am.masked_fill(am != 0, float(-100.0)).masked_fill(am == 0, float(0.0))


class SwinEngine:
    """Forward executor: bf16 working weights, per-block attention tables, activation buffers per batch size."""

    def __init__(self, model: "SwinTransformerV2", device: torch.device):
        if device.type != "cuda":
            raise _lib.RgbnmError("rgbnm: the SwinV2 engine needs a CUDA device; there is no CPU fallback")
        self.model = model
        self.device = device
        self.L = _lib.load()
        self._versions = None
        self._bufs: Dict[int, dict] = {}
        self.launches = 0
        self.refresh_weights()

    # -- weights ------------------------------------------------------------------------------------
    def _param_versions(self):
        return tuple(p._version for p in self.model.parameters())

    def weights_stale(self) -> bool:
        return self._versions != self._param_versions()

    def _wb(self, w: torch.Tensor) -> torch.Tensor:
        from . import ops
        out = torch.empty(w.shape, dtype=torch.bfloat16, device=self.device)
        ops.weight_prep(w.detach().contiguous(), out, None)
        return out

    @torch.no_grad()
    def refresh_weights(self) -> None:
        m = self.model
        f32 = lambda t: t.detach().to(self.device, torch.float32).contiguous()
        self.embed = (self._wb(m.patch_embed.projection[0].weight), f32(m.patch_embed.projection[0].bias),
                      f32(m.patch_embed.norm.weight), f32(m.patch_embed.norm.bias))
        self.stages: List[dict] = []
        for li, layer in enumerate(m.layers):
            blocks = []
            for blk in layer.blocks:
                a = blk.attn
                heads = a.num_heads
                # continuous relative position bias + clamped logit scale (swinv2.py:158-168): a function of the
                # parameters only, evaluated once per weight refresh (fp32, 225 x heads values)
                h = F.relu(F.linear(a.relative_coords_table.to(self.device), a.cpb_mlp[0].weight, a.cpb_mlp[0].bias))
                t = F.linear(h, a.cpb_mlp[2].weight).view(-1, heads)
                n = a.window_size[0] * a.window_size[1]
                rpb = t[a.relative_position_index.to(self.device).view(-1)].view(n, n, heads).permute(2, 0, 1)
                bias = (16 * torch.sigmoid(rpb)).contiguous().float()
                scale = torch.clamp(a.logit_scale, max=math.log(1.0 / 0.01)).exp().reshape(heads).contiguous().float()
                qkv_bias = None
                if a.q_bias is not None:
                    qkv_bias = torch.cat((a.q_bias, torch.zeros_like(a.v_bias), a.v_bias)).detach().float().contiguous()
                blocks.append(dict(
                    heads=heads, window=a.window_size[0], shift=blk.shift_size, bias=bias, scale=scale,
                    qkv=self._wb(a.qkv.weight), qkv_bias=qkv_bias, proj=self._wb(a.proj.weight), proj_b=f32(a.proj.bias),
                    fc1=self._wb(blk.mlp.fc1.weight), fc1_b=f32(blk.mlp.fc1.bias), fc2=self._wb(blk.mlp.fc2.weight),
                    fc2_b=f32(blk.mlp.fc2.bias), n1=(f32(blk.norm1.weight), f32(blk.norm1.bias)),
                    n2=(f32(blk.norm2.weight), f32(blk.norm2.bias))))
            ds = None
            if layer.downsample is not None:
                ds = dict(red=self._wb(layer.downsample.reduction.weight),
                          norm=(f32(layer.downsample.norm.weight), f32(layer.downsample.norm.bias)))
            self.stages.append(dict(dim=layer.dim, res=layer.input_resolution[0], blocks=blocks, down=ds))
        self.norm = (f32(m.norm.weight), f32(m.norm.bias))
        self.head = (self._wb(m.head.weight), f32(m.head.bias))
        self._versions = self._param_versions()

    # -- buffers ------------------------------------------------------------------------------------
    def _alloc(self, B: int) -> dict:
        if B in self._bufs:
            return self._bufs[B]
        bf = lambda *s: torch.empty(s, dtype=torch.bfloat16, device=self.device)
        bufs = {"stages": []}
        for st in self.stages:
            T, Cd = B * st["res"] ** 2, st["dim"]
            d = dict(xa=bf(T, Cd), xb=bf(T, Cd), qkv=bf(T, 3 * Cd), att=bf(T, Cd), tmp=bf(T, Cd), f=bf(T, 4 * Cd))
            if st["down"] is not None:
                d["gath"] = bf(T // 4, 4 * Cd)
                d["red"] = bf(T // 4, 2 * Cd)
            bufs["stages"].append(d)
        last = self.stages[-1]
        bufs["final"] = bf(B * last["res"] ** 2, last["dim"])
        bufs["pooled"] = bf(B, last["dim"])
        bufs["logits"] = torch.empty((B, self.head[0].shape[0]), dtype=torch.float32, device=self.device)
        self._bufs[B] = bufs
        return bufs

    # -- kernels ------------------------------------------------------------------------------------
    def _ln(self, x, norm, res, y, eps=1e-5):
        rows, emb = x.shape
        _lib.check(self.L.rgbnm_layernorm_res_fwd(x.data_ptr(), norm[0].data_ptr(), norm[1].data_ptr(),
                                                  None if res is None else res.data_ptr(), y.data_ptr(), rows, emb, eps,
                                                  _lib.stream_ptr()), "rgbnm_layernorm_res_fwd")
        self.launches += 1
        return y

    def _gemm(self, *a, **kw):
        self.launches += 1
        return G.gemm(*a, **kw)

    def _attn(self, qkv, out, blk, B, H, Cd):
        _lib.check(self.L.rgbnm_window_attention_fwd(qkv.data_ptr(), out.data_ptr(), blk["bias"].data_ptr(), blk["scale"].data_ptr(),
                                                     B, H, H, Cd, blk["heads"], blk["window"], blk["shift"], _lib.stream_ptr()),
                   "rgbnm_window_attention_fwd")
        self.launches += 1

    @torch.no_grad()
    @_lib.no_autocast
    def forward(self, x_in: torch.Tensor, collect: Optional[list] = None) -> torch.Tensor:
        """x_in: (B, 4096, 24) bf16 operand of the patch projection -> (B, n_classes) fp32 logits."""
        if x_in.dtype != torch.bfloat16:
            x_in = x_in.to(torch.bfloat16)
        B = x_in.shape[0]
        x_in = x_in.reshape(B * x_in.shape[1], IN_FEAT).contiguous()
        bufs = self._alloc(B)
        s0 = bufs["stages"][0]
        wE, bE, gE, beE = self.embed
        if not _NOFUSE:
            x = self._gemm(x_in, wE, G.EPI_LN, bias=bE, ln=(gE, beE), out=s0["xa"])         # Linear(24, 96) + patch_embed.norm
        else:
            self._gemm(x_in, wE, G.EPI_STORE, bias=bE, out=s0["tmp"])
            x = self._ln(s0["tmp"], (gE, beE), None, s0["xa"])
        for li, st in enumerate(self.stages):
            b = bufs["stages"][li]
            H, Cd = st["res"], st["dim"]
            for bi, blk in enumerate(st["blocks"]):
                other = b["xb"] if x is b["xa"] else b["xa"]
                self._gemm(x, blk["qkv"], G.EPI_STORE, bias=blk["qkv_bias"], out=b["qkv"])
                self._attn(b["qkv"], b["att"], blk, B, H, Cd)
                if Cd <= _FUSE_MAX_DIM and not _NOFUSE:
                    # x + norm1(proj(.)) and x1 + norm2(fc2(.)) inside the GEMM epilogue (RGBNM_EPI_LNRES): the row fits one tile
                    x1 = self._gemm(b["att"], blk["proj"], G.EPI_LNRES, bias=blk["proj_b"], aux=x, ln=blk["n1"], out=other)
                    self._gemm(x1, blk["fc1"], G.EPI_GELU_ACT, bias=blk["fc1_b"], out=b["f"])   # only gelu(fc1) is kept (inference)
                    x = self._gemm(b["f"], blk["fc2"], G.EPI_LNRES, bias=blk["fc2_b"], aux=x1, ln=blk["n2"], out=x)
                else:
                    self._gemm(b["att"], blk["proj"], G.EPI_STORE, bias=blk["proj_b"], out=b["tmp"])
                    x1 = self._ln(b["tmp"], blk["n1"], x, other)                   # x + norm1(attn(x))
                    self._gemm(x1, blk["fc1"], G.EPI_GELU_ACT, bias=blk["fc1_b"], out=b["f"])
                    self._gemm(b["f"], blk["fc2"], G.EPI_STORE, bias=blk["fc2_b"], out=b["tmp"])
                    x = self._ln(b["tmp"], blk["n2"], x1, x)                       # x1 + norm2(mlp(x1)); rows are independent
                if collect is not None:
                    collect.append((f"l{li}b{bi}", x.view(B, H * H, Cd).float().clone()))
            if st["down"] is not None:
                _lib.check(self.L.rgbnm_patch_merge_gather(x.data_ptr(), b["gath"].data_ptr(), B, H, H, Cd, _lib.stream_ptr()),
                           "rgbnm_patch_merge_gather")
                self.launches += 1
                if 2 * Cd <= _FUSE_MAX_DIM and not _NOFUSE:
                    x = self._gemm(b["gath"], st["down"]["red"], G.EPI_LN, ln=st["down"]["norm"], out=bufs["stages"][li + 1]["xa"])
                else:
                    self._gemm(b["gath"], st["down"]["red"], G.EPI_STORE, out=b["red"])
                    x = self._ln(b["red"], st["down"]["norm"], None, bufs["stages"][li + 1]["xa"])
                if collect is not None:
                    collect.append((f"stage{li}", x.view(B, (H // 2) ** 2, 2 * Cd).float().clone()))
            elif collect is not None:
                collect.append((f"stage{li}", x.view(B, H * H, Cd).float().clone()))
        last = self.stages[-1]
        xf = self._ln(x, self.norm, None, bufs["final"])
        _lib.check(self.L.rgbnm_token_mean_bf16(xf.data_ptr(), bufs["pooled"].data_ptr(), B, last["res"] ** 2, last["dim"],
                                                _lib.stream_ptr()), "rgbnm_token_mean_bf16")
        self.launches += 1
        return self._gemm(bufs["pooled"], self.head[0], G.EPI_F32, bias=self.head[1], out_f32=bufs["logits"])


class SwinTransformerV2(nn.Module):
    """Same constructor surface as the reference SwinTransformerV2 (swinv2.py:606-613); DCT pixel space only."""

    def __init__(self, img_size=224, patch_size=4, in_chans=3, num_classes=1000, embed_dim=96, depths=[2, 2, 6, 2],
                 num_heads=[3, 6, 12, 24], window_size=7, mlp_ratio=4., qkv_bias=True, drop_rate=0., attn_drop_rate=0.,
                 drop_path_rate=0.1, norm_layer=nn.LayerNorm, ape=False, patch_norm=True, use_checkpoint=False,
                 pretrained_window_sizes=[0, 0, 0, 0], device="cpu", pixel_space="rgb", **kwargs):
        super().__init__()
        if str(pixel_space).lower() != "dct" or patch_size != 4:
            raise NotImplementedError("rgbnm: only pixel_space='dct', patch_size=4 is on the B200 hot path")
        if ape or not patch_norm or norm_layer is not nn.LayerNorm or any(int(p) != 0 for p in pretrained_window_sizes):
            raise NotImplementedError("rgbnm: ape / no patch_norm / pretrained window sizes are not on the hot path "
                                      "(reference defaults: utils/configs.py:123-137)")
        if drop_rate != 0 or attn_drop_rate != 0:
            raise NotImplementedError("rgbnm: dropout > 0 is not on the hot path (reference default MODEL.DROP = 0)")
        if img_size % (patch_size * window_size) != 0:
            raise NotImplementedError("rgbnm: the token grid must be a multiple of the window")
        dev = torch.device(device) if not isinstance(device, int) else torch.device("cuda", device)
        self.num_classes, self.num_layers, self.embed_dim = num_classes, len(depths), embed_dim
        self.num_features = int(embed_dim * 2 ** (self.num_layers - 1))
        self.mlp_ratio, self.pixel_space, self.drop_path_rate = mlp_ratio, "dct", drop_path_rate
        self.patches_resolution = [img_size // patch_size, img_size // patch_size]

        def lin(i, o, bias=True):
            return nn.Linear(i, o, bias=bias, device=dev, dtype=torch.float32)

        def ln(e):
            return nn.LayerNorm(e, device=dev)
        self.patch_embed = _Box()
        self.patch_embed.projection = nn.Sequential(lin(IN_FEAT, embed_dim))
        self.patch_embed.norm = ln(embed_dim)
        self.layers = nn.ModuleList()
        for li in range(self.num_layers):
            dim, res = int(embed_dim * 2 ** li), self.patches_resolution[0] // 2 ** li
            layer = _Box()
            layer.dim, layer.input_resolution, layer.depth = dim, (res, res), depths[li]
            layer.blocks = nn.ModuleList()
            for bi in range(depths[li]):
                ws, shift = window_size, (0 if bi % 2 == 0 else window_size // 2)
                if res <= ws:                                   # swinv2.py:208-211
                    shift, ws = 0, res
                blk = _Box()
                blk.dim, blk.num_heads, blk.window_size, blk.shift_size = dim, num_heads[li], ws, shift
                blk.norm1, blk.norm2 = ln(dim), ln(dim)
                a = _Box()
                a.dim, a.num_heads, a.window_size = dim, num_heads[li], (ws, ws)
                a.logit_scale = nn.Parameter(torch.log(10 * torch.ones((num_heads[li], 1, 1), device=dev)))
                a.cpb_mlp = nn.Sequential(lin(2, 512), nn.ReLU(inplace=True), lin(512, num_heads[li], bias=False))
                table, index = _relative_tables(ws)
                a.register_buffer("relative_coords_table", table.to(dev))
                a.register_buffer("relative_position_index", index.to(dev))
                a.qkv = lin(dim, 3 * dim, bias=False)
                a.q_bias = nn.Parameter(torch.zeros(dim, device=dev)) if qkv_bias else None
                a.v_bias = nn.Parameter(torch.zeros(dim, device=dev)) if qkv_bias else None
                a.proj = lin(dim, dim)
                blk.attn = a
                blk.mlp = _Box()
                blk.mlp.fc1, blk.mlp.fc2 = lin(dim, int(dim * mlp_ratio)), lin(int(dim * mlp_ratio), dim)
                blk.register_buffer("attn_mask", _shift_mask(res, res, ws, shift).to(dev) if shift > 0 else None)
                layer.blocks.append(blk)
            if li < self.num_layers - 1:
                layer.downsample = _Box()
                layer.downsample.reduction = lin(4 * dim, 2 * dim, bias=False)
                layer.downsample.norm = ln(2 * dim)
            else:
                layer.downsample = None
            self.layers.append(layer)
        self.norm = ln(self.num_features)
        self.head = lin(self.num_features, num_classes)
        # swinv2.py:660-674: trunc_normal_(.02) Linear weights, zero biases, unit LayerNorm, then zero-initialised
        # post-norms inside the blocks (_init_respostnorm)
        for mod in self.modules():
            if isinstance(mod, nn.Linear):
                nn.init.trunc_normal_(mod.weight, std=.02)
                if mod.bias is not None:
                    nn.init.constant_(mod.bias, 0)
            elif isinstance(mod, nn.LayerNorm):
                nn.init.constant_(mod.bias, 0)
                nn.init.constant_(mod.weight, 1.0)
        for layer in self.layers:
            for blk in layer.blocks:
                for n in (blk.norm1, blk.norm2):
                    nn.init.constant_(n.bias, 0)
                    nn.init.constant_(n.weight, 0)
        self._engine: Optional[SwinEngine] = None
        self._train_engine = None

    def no_weight_decay(self):
        return {"absolute_pos_embed"}

    def no_weight_decay_keywords(self):
        return {"cpb_mlp", "logit_scale", "relative_position_bias_table"}

    def prepare(self, device=None) -> SwinEngine:
        dev = torch.device(device) if device is not None else next(self.parameters()).device
        if self._engine is None or self._engine.device != dev:
            if next(self.parameters()).device != dev:
                self.to(dev)
            self._engine = SwinEngine(self, dev)
        return self._engine

    def forward(self, y, cbcr=None):
        """forward(y, cbcr) with reference-format tensors (swinv2.py:703-705), or forward(x) with the (B,4096,24) tensor
        FusedDCT(out_size=32) writes."""
        if getattr(cbcr, "_rgbnm_absent", False):           # compat.loader.DCTBatch: chroma already merged by the fused kernel
            cbcr = None
        if cbcr is not None:
            y = swin_embed_input_from_planes(y, cbcr)
        elif y.dim() != 3 or y.shape[2] != IN_FEAT:
            raise ValueError("rgbnm SwinV2: expected (y, cbcr) in the reference layout or a (B,4096,24) embed input")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # training path (swin_train.py): explicit forward / backward behind one autograd.Function
            from . import swin_train as ST
            if y.device.type != "cuda":
                raise _lib.RgbnmError("rgbnm: the SwinV2 training engine needs a CUDA device; there is no CPU fallback")
            if self._train_engine is None or self._train_engine.device != y.device:
                if next(self.parameters()).device != y.device:
                    self.to(y.device)
                self._train_engine = ST.SwinTrainEngine(self, y.device)
            names = [n for n, _ in self.named_parameters()]
            return ST.SwinFunction.apply(y, self._train_engine, names, *[p for _, p in self.named_parameters()])
        eng = self.prepare(y.device)
        if eng.weights_stale():
            eng.refresh_weights()
        return eng.forward(y)
