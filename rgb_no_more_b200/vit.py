"""DCT-domain ViT on B200: drop-in for the reference `models.plainvit.ViT` (boundary B3).

`ViT(...)` takes the constructor arguments of /root/reference/models/plainvit.py:559-600 as passed by
`utils.get_model` (/root/reference/utils/pipeline_utils.py:335-349), exposes `forward(y, cbcr)`
(plainvit.py:601-612) and a `state_dict()` with the reference's 152 keys and shapes, so checkpoints
move both ways.  The parameter-holding sub-modules exist only to give the parameters their reference
names; the arithmetic is done by `ViTEngine`, an explicit forward/backward schedule over hand-written
sm_100a kernels called through the C-ABI (include/rgbnm_b200.h):

  patch projection + sincos position embedding   tcgen05 GEMM, epilogue bias+posemb      plainvit.py:194-198, 97-121
  LayerNorm                                       rgbnm_layernorm_fwd / _bwd               :513, 522, 551
  fused qkv projection                            tcgen05 GEMM (rows regrouped q|k|v)      :441, 447
  attention softmax(QK^T / sqrt(emb)) V           attention kernel                         :450-461
  output projection + residual                    tcgen05 GEMM, epilogue bias+residual     :443, 463, 475-479
  MLP fc1 + GELU(erf), fc2 + residual             tcgen05 GEMMs, epilogues bias+GELU / bias+residual   :485-491
  backward                                        dgrad GEMMs on transposed bf16 weight copies, wgrad GEMMs with
                                                  MN-major operands and fp32 red.add into one flat gradient buffer

Numerics follow the reference under `--amp 1 --ampdtype bf16`: bf16 GEMM inputs, fp32 accumulation,
fp32 LayerNorm / softmax statistics, bf16 residual stream (`x += res` is in place on the bf16 branch
output, plainvit.py:478), fp32 master weights.  There is no CPU fallback."""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import gemm as G
from . import lib as _lib
from . import ops as K

TOKENS = 196          # 14 x 14 patches of 16 x 16 pixels (224 px input)
IN_FEAT = 384         # 256 luma + 64 Cb + 64 Cr coefficients per patch


class _Box(nn.Module):
    """Parameter container mirroring one level of the reference module tree."""


def sincos_posemb(h: int, w: int, e: int, device) -> torch.Tensor:
    """SinCosEmbedding (plainvit.py:97-121) evaluated once: (h*w, e) fp32, added to the projected patches."""
    hg, wg = torch.meshgrid(torch.arange(h, device=device), torch.arange(w, device=device), indexing="ij")
    freq = torch.log(torch.tensor(10000, dtype=torch.int32, device=device)) / (e // 4 - 1)
    freq = torch.exp(-torch.arange(e // 4, dtype=torch.float32, device=device) * freq)
    ph = torch.einsum("p,f->pf", hg.flatten().float(), freq)
    pw = torch.einsum("p,f->pf", wg.flatten().float(), freq)
    return torch.cat((pw.sin(), pw.cos(), ph.sin(), ph.cos()), dim=-1).contiguous()


def embed_input_from_planes(y: torch.Tensor, cbcr: torch.Tensor, subblock: bool = True) -> torch.Tensor:
    """Compatibility path for reference-format inputs (B,1,28,28,8,8) + (B,2,14,14,8,8), already ToRange'd:
    the tail of the fused kernel (2x2 luma blocks -> 16x16 sub-block, collapse, concat; plainvit.py:200-216)
    expressed with torch ops on the GPU.  The fast path never comes here: FusedDCT writes this tensor directly.
    subblock = False: the un-converted 16 x 16 tile (`--no_subblock`, FusedDCT(subblock=False))."""
    b = y.shape[0]
    if not subblock:
        yy = y.float().reshape(b, 14, 2, 14, 2, 8, 8).permute(0, 1, 3, 2, 5, 4, 6).reshape(b, 14, 14, 256)
        cc = cbcr.float().permute(0, 2, 3, 1, 4, 5).reshape(b, 14, 14, 128)
        return torch.cat([yy, cc], dim=3).reshape(b, TOKENS, IN_FEAT)
    n = torch.arange(8, dtype=torch.float32, device=y.device)

    def basis(L):
        k = torch.arange(L, dtype=torch.float32, device=y.device).unsqueeze(1)
        m = (torch.arange(L, dtype=torch.float32, device=y.device).unsqueeze(0) + 0.5)
        d = (k.mm(m) * torch.pi / L).cos()
        d[0] *= 1 / (2 ** 0.5)
        return d * (2 / L) ** 0.5
    a16 = basis(16).mm(torch.block_diag(basis(8), basis(8)).T)
    del n
    yy = y.float().reshape(b, 14, 2, 14, 2, 8, 8).permute(0, 1, 3, 2, 5, 4, 6).reshape(b, 14, 14, 16, 16)
    yy = torch.matmul(torch.matmul(a16, yy), a16.T).reshape(b, 14, 14, 256)
    cc = cbcr.float().permute(0, 2, 3, 1, 4, 5).reshape(b, 14, 14, 128)
    return torch.cat([yy, cc], dim=3).reshape(b, TOKENS, IN_FEAT)


# ------------------------------------------------------------------------------------------------
# Engine
# ------------------------------------------------------------------------------------------------
class _Lin:
    """One nn.Linear of the model: fp32 master views + bf16 working copies.  `weight` / `bias` may also be plain fp32 tensors
    assembled from several parameters (the block-diagonal first stage of the separate embeddings)."""

    def __init__(self, name: str, weight, bias, qkv_heads: int = 0):
        self.name, self.weight, self.bias, self.qkv_heads = name, weight, bias, qkv_heads
        self.n, self.k = weight.shape
        self.wb = self.wt = self.bias_k = None      # bf16 [n,k], bf16 [k,n], fp32 bias in kernel order


def _gelu_grad(u: torch.Tensor) -> torch.Tensor:
    """d/du gelu_erf(u) (nn.GELU() default), fp32."""
    uf = u.float()
    return 0.5 * (1.0 + torch.erf(uf * 0.7071067811865476)) + uf * torch.exp(-0.5 * uf * uf) * 0.3989422804014327


class ViTEngine:
    def __init__(self, model: "ViT", device: torch.device, attention: str = "auto"):
        self.dev = device
        self.E, self.depth, self.H, self.D = model.emb_size, model.depth, model.num_heads, model.head_size
        self.HD = self.H * self.D
        self.n_classes = model.n_classes
        self.attention_backend = attention
        self._lib = _lib.load()
        if device.type != "cuda":
            raise _lib.RgbnmError("rgbnm: the B200 ViT engine needs a CUDA device; there is no CPU fallback")
        named = OrderedDict(model.named_parameters())
        # flat fp32 parameter / gradient buffers: Linear weights first (they are the decoupled-weight-decay group,
        # pipeline_utils.py:537: names containing '.weight' and not 'lrnorm'), everything else after
        decay = [k for k, p in named.items() if p.ndim == 2]
        rest = [k for k in named if k not in set(decay)]
        self.order = decay + rest
        self.n_decay = sum(named[k].numel() for k in decay)
        total = sum(p.numel() for p in named.values())
        self.flat = torch.empty(total, dtype=torch.float32, device=device)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=device)
        self.grad_views: Dict[str, torch.Tensor] = {}
        off = 0
        with torch.no_grad():
            for k in self.order:
                p = named[k]
                n = p.numel()
                view = self.flat[off:off + n].view_as(p)
                view.copy_(p.detach().to(device=device, dtype=torch.float32))
                p.data = view
                self.grad_views[k] = self.flat_grad[off:off + n].view_as(p)
                off += n
        self.params = named
        P = named
        # patch embedding (plainvit.py:140-283): "group" = one Linear on the K0 operand (embed_type 1); "sep_sub" / "sep" = embed_type 2
        # with / without sub-block conversion: per-group Linears (a block-diagonal first stage, assembled into ONE [E, 384] matrix on
        # the K0 column order) -> GELU -> channel-mixing Linear (+ residual for sep_sub)
        self.embed_kind = model.embed_kind
        self.lin_mix = None
        if self.embed_kind == "group":
            self.lin_embed = _Lin("embed", P["patchembed.projection.0.weight"], P["patchembed.projection.0.bias"])
        else:
            self.emb_groups = []          # (weight param, bias param, first output row, K0 column index of every input feature)
            if self.embed_kind == "sep_sub":
                wy, wc = P["patchembed.projection_Y.1.weight"], P["patchembed.projection_C.1.weight"]
                self.emb_groups.append((wy, P["patchembed.projection_Y.1.bias"], 0, torch.arange(0, 256, device=device)))
                self.emb_groups.append((wc, P["patchembed.projection_C.1.bias"], wy.shape[0], torch.arange(256, 384, device=device)))
                mix = ("patchembed.linearMix.weight", "patchembed.linearMix.bias")
            else:
                eo = self.E // 6
                p12 = torch.arange(64, device=device)
                for g in range(4):        # LinearY[g]: block (pdh, pdw) = (g // 2, g % 2) of the un-converted 16 x 16 tile, row-major
                    cols = ((g // 2) * 8 + p12 // 8) * 16 + (g % 2) * 8 + p12 % 8
                    self.emb_groups.append((P[f"patchembed.LinearY.{g}.weight"], P[f"patchembed.LinearY.{g}.bias"], g * eo, cols))
                for ci in range(2):
                    self.emb_groups.append((P[f"patchembed.LinearC.{ci}.weight"], P[f"patchembed.LinearC.{ci}.bias"], (4 + ci) * eo,
                                            256 + ci * 64 + p12))
                mix = ("patchembed.LinearMix.weight", "patchembed.LinearMix.bias")
            n1 = sum(g[0].shape[0] for g in self.emb_groups)
            if n1 != self.E:
                raise NotImplementedError("rgbnm: separate embeddings need emb_size divisible by 6")
            self.W1 = torch.zeros((n1, IN_FEAT), dtype=torch.float32, device=device)
            self.b1 = torch.zeros(n1, dtype=torch.float32, device=device)
            self.gW1 = torch.zeros_like(self.W1)
            self.lin_embed = _Lin("embed1", self.W1, self.b1)
            self.lin_mix = _Lin("mix", P[mix[0]], P[mix[1]])
        self.layers = []
        for l in range(self.depth):
            pre = f"encoder.{l}."
            self.layers.append(dict(
                ln1=(P[pre + "0.fn.eb_lrnorm1.weight"], P[pre + "0.fn.eb_lrnorm1.bias"], pre + "0.fn.eb_lrnorm1"),
                qkv=_Lin("qkv", P[pre + "0.fn.eb_mha.qkv.weight"], P[pre + "0.fn.eb_mha.qkv.bias"], qkv_heads=self.H),
                proj=_Lin("proj", P[pre + "0.fn.eb_mha.projection.weight"], P[pre + "0.fn.eb_mha.projection.bias"]),
                ln2=(P[pre + "1.fn.eb_lrnorm2.weight"], P[pre + "1.fn.eb_lrnorm2.bias"], pre + "1.fn.eb_lrnorm2"),
                fc1=_Lin("fc1", P[pre + "1.fn.eb_ffb.0.weight"], P[pre + "1.fn.eb_ffb.0.bias"]),
                fc2=_Lin("fc2", P[pre + "1.fn.eb_ffb.3.weight"], P[pre + "1.fn.eb_ffb.3.bias"]),
            ))
        self.lnh = (P["classhead.ch_lrnorm.weight"], P["classhead.ch_lrnorm.bias"], "classhead.ch_lrnorm")
        self.head1 = (P["classhead.ch_linear1.weight"], P["classhead.ch_linear1.bias"])
        self.head2 = (P["classhead.ch_linear2.weight"], P["classhead.ch_linear2.bias"])
        self._name_of = {id(p): k for k, p in named.items()}
        bf = dict(dtype=torch.bfloat16, device=device)
        for lin in self._all_lins():
            lin.wb = torch.empty((lin.n, lin.k), **bf)
            lin.wt = torch.empty((lin.k, lin.n), **bf)
            if lin.qkv_heads:
                lin.bias_k = torch.empty(lin.n, dtype=torch.float32, device=device)
        self.posemb = sincos_posemb(14, 14, self.E, device)
        self.posemb_bf16 = self.posemb.to(torch.bfloat16)
        self._versions = None
        self._wprep = None
        self.bufs = None
        self.batch = 0
        self.launches = 0
        self.refresh_weights()

    def _all_lins(self) -> List[_Lin]:
        out = [self.lin_embed] + ([self.lin_mix] if self.lin_mix is not None else [])
        for ly in self.layers:
            out += [ly["qkv"], ly["proj"], ly["fc1"], ly["fc2"]]
        return out

    def grad_of(self, p: nn.Parameter) -> torch.Tensor:
        return self.grad_views[self._name_of[id(p)]]

    # -- weights ------------------------------------------------------------------------------------
    def refresh_weights(self) -> None:
        """fp32 master -> bf16 working copies (+ transposes for dgrad, qkv rows and bias regrouped q|k|v): one launch
        over a device-resident descriptor table (the pointers into the flat buffer never change)."""
        if self.embed_kind != "group":
            with torch.no_grad():         # assemble the block-diagonal first stage on the K0 column order
                for w, b, r0, cols in self.emb_groups:
                    self.W1[r0:r0 + w.shape[0]].index_copy_(1, cols, w.data)
                    self.b1[r0:r0 + w.shape[0]].copy_(b.data)
        if self._wprep is None:
            import numpy as np
            lins = self._all_lins()
            arr = (_lib.WPrepDesc * len(lins))()
            tiles = 0
            for i, lin in enumerate(lins):
                d = arr[i]
                d.w, d.w_bf16, d.wt_bf16 = lin.weight.data_ptr(), lin.wb.data_ptr(), lin.wt.data_ptr()
                d.bias = lin.bias.data.data_ptr() if lin.qkv_heads else None
                d.bias_k = lin.bias_k.data_ptr() if lin.qkv_heads else None
                d.n, d.k, d.qkv_heads, d.head_dim, d.first_tile = lin.n, lin.k, lin.qkv_heads, self.D if lin.qkv_heads else 0, tiles
                tiles += -(-lin.n // 32) * -(-lin.k // 32)
            host = torch.from_numpy(np.frombuffer(bytes(arr), dtype=np.uint8).copy())
            self._wprep = (host.to(self.dev), len(lins), tiles)
        table, n, tiles = self._wprep
        _lib.check(self._lib.rgbnm_weight_prep_batch(table.data_ptr(), n, tiles, _lib.stream_ptr()), "rgbnm_weight_prep_batch")
        self.launches += 1
        self._versions = self._param_versions()

    def _param_versions(self):
        return tuple(p._version for p in self.params.values()) + (self.flat._version,)

    def weights_stale(self) -> bool:
        return self._versions != self._param_versions()

    # -- buffers ------------------------------------------------------------------------------------
    def _alloc(self, B: int) -> None:
        if self.bufs is not None and self.batch == B:
            return
        M, E, HD = B * TOKENS, self.E, self.HD
        bf = dict(dtype=torch.bfloat16, device=self.dev)
        f32 = dict(dtype=torch.float32, device=self.dev)
        L = []
        for _ in range(self.depth):
            L.append(dict(x_in=None, h1=torch.empty((M, E), **bf), qkv=torch.empty((M, 3 * HD), **bf),
                          o=torch.empty((M, HD), **bf), x_mid=torch.empty((M, E), **bf), h2=torch.empty((M, E), **bf),
                          u=torch.empty((M, 4 * E), **bf), f=torch.empty((M, 4 * E), **bf), x_out=torch.empty((M, E), **bf),
                          mean1=torch.empty(M, **f32), rstd1=torch.empty(M, **f32), mean2=torch.empty(M, **f32),
                          rstd2=torch.empty(M, **f32), lse=torch.empty((B, self.H, TOKENS), **f32)))
        extra = {} if self.embed_kind == "group" else dict(emb_u=torch.empty((M, E), **bf), emb_f=torch.empty((M, E), **bf),
                                                           emb_df=torch.empty((M, E), **bf))
        self.bufs = dict(layers=L, **extra, x0=torch.empty((M, E), **bf), hN=torch.empty((M, E), **bf), meanH=torch.empty(M, **f32),
                         rstdH=torch.empty(M, **f32),
                         # backward scratch (shared by all layers)
                         dA=torch.empty((M, E), **bf), dB=torch.empty((M, E), **bf), dC=torch.empty((M, E), **bf),
                         dU=torch.empty((M, 4 * E), **bf), dO=torch.empty((M, HD), **bf), dQKV=torch.empty((M, 3 * HD), **bf),
                         dvec=torch.empty((B, self.H, TOKENS), **f32))
        self.batch = B

    # -- attention ------------------------------------------------------------------------------------
    def _attn_fwd(self, qkv: torch.Tensor, o: torch.Tensor, lse: torch.Tensor, B: int) -> None:
        from . import attention as A
        A.forward(qkv, o, lse, B, self.H, self.D, 1.0 / math.sqrt(self.E), backend=self.attention_backend)
        self.launches += 1

    def _attn_bwd(self, do: torch.Tensor, qkv: torch.Tensor, o: torch.Tensor, lse: torch.Tensor, dqkv: torch.Tensor, B: int) -> None:
        from . import attention as A
        A.backward(do, qkv, o, lse, dqkv, B, self.H, self.D, 1.0 / math.sqrt(self.E), backend=self.attention_backend,
                   dvec=self.bufs["dvec"])
        self.launches += 3

    # -- forward ------------------------------------------------------------------------------------
    def _ln(self, x, ln, y, mean, rstd):
        K.layernorm_fwd(x, ln[0].data, ln[1].data, y, mean, rstd, 1e-5)
        self.launches += 1

    def _gemm(self, *a, **kw):
        self.launches += 1
        return G.gemm(*a, **kw)

    @_lib.no_autocast
    def forward(self, x_in: torch.Tensor, save: bool = True) -> torch.Tensor:
        """x_in: (B, 196, 384) bf16 operand of the patch projection -> (B, n_classes) fp32 logits."""
        if x_in.dtype != torch.bfloat16:
            x_in = x_in.to(torch.bfloat16)
        B = x_in.shape[0]
        M = B * TOKENS
        x_in = x_in.reshape(M, IN_FEAT)
        if not x_in.is_contiguous():
            x_in = x_in.contiguous()
        self._alloc(B)
        bufs = self.bufs
        bufs["x_embed_in"] = x_in
        if self.embed_kind == "group":
            x = self._gemm(x_in, self.lin_embed.wb, G.EPI_POSEMB, bias=self.lin_embed.bias.data, posemb=self.posemb, out=bufs["x0"])
        else:
            # per-group Linears -> GELU (plainvit.py:263-271, 342-347) -> channel mix (+ residual, :348-350) -> sincos
            self._gemm(x_in, self.lin_embed.wb, G.EPI_GELU, bias=self.b1, out=bufs["emb_u"], out2=bufs["emb_f"])
            if self.embed_kind == "sep_sub":
                x = self._gemm(bufs["emb_f"], self.lin_mix.wb, G.EPI_RESIDUAL, bias=self.lin_mix.bias.data, aux=bufs["emb_f"], out=bufs["x0"])
                x.view(B, TOKENS, self.E).add_(self.posemb_bf16)
            else:
                x = self._gemm(bufs["emb_f"], self.lin_mix.wb, G.EPI_POSEMB, bias=self.lin_mix.bias.data, posemb=self.posemb, out=bufs["x0"])
        for l, ly in enumerate(self.layers):
            b = bufs["layers"][l]
            b["x_in"] = x
            self._ln(x, ly["ln1"], b["h1"], b["mean1"], b["rstd1"])
            self._gemm(b["h1"], ly["qkv"].wb, G.EPI_STORE, bias=ly["qkv"].bias_k, out=b["qkv"])
            self._attn_fwd(b["qkv"], b["o"], b["lse"], B)
            self._gemm(b["o"], ly["proj"].wb, G.EPI_RESIDUAL, bias=ly["proj"].bias.data, aux=x, out=b["x_mid"])
            self._ln(b["x_mid"], ly["ln2"], b["h2"], b["mean2"], b["rstd2"])
            self._gemm(b["h2"], ly["fc1"].wb, G.EPI_GELU, bias=ly["fc1"].bias.data, out=b["u"], out2=b["f"])
            x = self._gemm(b["f"], ly["fc2"].wb, G.EPI_RESIDUAL, bias=ly["fc2"].bias.data, aux=b["x_mid"], out=b["x_out"])
        bufs["x_last"] = x
        self._ln(x, self.lnh, bufs["hN"], bufs["meanH"], bufs["rstdH"])
        # classification head (plainvit.py:547-557): token mean -> Linear -> tanh -> Linear; B x E work, torch ops
        pooled = bufs["hN"].view(B, TOKENS, self.E).mean(dim=1, dtype=torch.float32)     # fp32 accumulation straight from bf16
        bufs["pooled"] = pooled
        z = torch.tanh(F.linear(pooled, self.head1[0].data, self.head1[1].data))
        bufs["z"] = z
        return F.linear(z, self.head2[0].data, self.head2[1].data)

    # -- backward -----------------------------------------------------------------------------------
    def _wgrad_gemm(self, dy: torch.Tensor, x: torch.Tensor, gw: torch.Tensor, splits: int) -> None:
        # gw [out, in] += dy^T x.  The CTA-pair kernel tiles 256 (M) x 128 (N): put the longer side on M.
        if x.shape[1] > dy.shape[1]:
            self._gemm(x, dy, G.EPI_WGRAD_ATOMIC, out_f32=gw, splits=splits, trans_out=True)
        else:
            self._gemm(dy, x, G.EPI_WGRAD_ATOMIC, out_f32=gw, splits=splits)

    def _wgrad(self, dy: torch.Tensor, x: torch.Tensor, lin: _Lin, splits: int, bias_done: bool = False) -> None:
        """dW += dy^T x, db += colsum(dy) into the flat gradient buffer (reference parameter order).  bias_done: the
        LayerNorm backward that produced dy already accumulated its column sums into the bias gradient."""
        if lin.qkv_heads:
            # the kernel-side qkv layout is q|k|v head-major: the epilogue / column-sum kernels add each row at its
            # reference position ("(h d qkv)", plainvit.py:447), so the flat buffer keeps the reference layout
            self._gemm(dy, x, G.EPI_WGRAD_ATOMIC, out_f32=self.grad_of(lin.weight), splits=splits,
                       perm_heads=self.H, perm_head_dim=self.D)
            K.colsum(dy, self.grad_of(lin.bias), self.H, self.D)
            self.launches += 1
        else:
            self._wgrad_gemm(dy, x, self.grad_of(lin.weight), splits)
            if not bias_done:
                K.colsum(dy, self.grad_of(lin.bias))
                self.launches += 1

    @_lib.no_autocast
    def backward(self, dlogits: torch.Tensor, zero_grad: bool = True) -> None:
        """Fill the flat gradient buffer (and nothing else) from d(loss)/d(logits), (B, n_classes) fp32."""
        bufs = self.bufs
        B = self.batch
        M, E = B * TOKENS, self.E
        if zero_grad:
            self.flat_grad.zero_()
        splits = 0                                                   # chosen by the library (whole waves of CTA pairs)
        # ---- head (torch ops on B x E) ----
        z, pooled = bufs["z"], bufs["pooled"]
        w2, w1 = self.head2[0].data, self.head1[0].data
        self.grad_of(self.head2[0]).addmm_(dlogits.t(), z)
        self.grad_of(self.head2[1]).add_(dlogits.sum(0))
        dz = dlogits.mm(w2) * (1.0 - z * z)
        self.grad_of(self.head1[0]).addmm_(dz.t(), pooled)
        self.grad_of(self.head1[1]).add_(dz.sum(0))
        dpooled = dz.mm(w1) * (1.0 / TOKENS)
        dA, dB, dC = bufs["dA"], bufs["dB"], bufs["dC"]
        # every LayerNorm backward below also accumulates the column sums of the dx it writes: that is the bias gradient of
        # the Linear whose output (+ residual) this dx is the gradient of (fc2 of the layer, proj of the layer, the embed).
        # The head's LayerNorm takes d(pooled) / tokens for every token of an image: one dy row per image, broadcast by the kernel.
        K.layernorm_bwd(dpooled.to(torch.bfloat16), bufs["x_last"], bufs["meanH"], bufs["rstdH"], self.lnh[0].data, None, dB,
                        self.grad_of(self.lnh[0]), self.grad_of(self.lnh[1]),
                        dxsum=self.grad_of(self.layers[-1]["fc2"].bias), rows_per_dy_row=TOKENS)
        self.launches += 1
        dx, spare1, spare2 = dB, dA, dC
        for l in reversed(range(self.depth)):
            ly, b = self.layers[l], bufs["layers"][l]
            # ---- MLP branch: x_out = x_mid + fc2(gelu(fc1(LN2(x_mid)))) ----
            self._wgrad(dx, b["f"], ly["fc2"], splits, bias_done=True)
            du = self._gemm(dx, ly["fc2"].wt, G.EPI_DGELU, aux=b["u"], out=bufs["dU"])
            self._wgrad(du, b["h2"], ly["fc1"], splits)
            dh2 = self._gemm(du, ly["fc1"].wt, G.EPI_STORE, out=spare1)
            K.layernorm_bwd(dh2, b["x_mid"], b["mean2"], b["rstd2"], ly["ln2"][0].data, dx, spare2,
                            self.grad_of(ly["ln2"][0]), self.grad_of(ly["ln2"][1]), dxsum=self.grad_of(ly["proj"].bias))
            self.launches += 1
            dx_mid = spare2
            # ---- attention branch: x_mid = x_in + proj(attn(qkv(LN1(x_in)))) ----
            self._wgrad(dx_mid, b["o"], ly["proj"], splits, bias_done=True)
            do = self._gemm(dx_mid, ly["proj"].wt, G.EPI_STORE, out=bufs["dO"])
            self._attn_bwd(do, b["qkv"], b["o"], b["lse"], bufs["dQKV"], B)
            self._wgrad(bufs["dQKV"], b["h1"], ly["qkv"], splits)
            dh1 = self._gemm(bufs["dQKV"], ly["qkv"].wt, G.EPI_STORE, out=spare1)
            below = self.layers[l - 1]["fc2"].bias if l > 0 else (self.lin_embed.bias if self.lin_mix is None else self.lin_mix.bias)
            K.layernorm_bwd(dh1, b["x_in"], b["mean1"], b["rstd1"], ly["ln1"][0].data, dx_mid, dx,
                            self.grad_of(ly["ln1"][0]), self.grad_of(ly["ln1"][1]), dxsum=self.grad_of(below))
            self.launches += 1
        if self.embed_kind == "group":
            self._wgrad(dx, bufs["x_embed_in"], self.lin_embed, splits, bias_done=True)
            return
        # separate embeddings: x0 = mix(f) (+ f) + posemb, f = gelu(u), u = x_in W1^T + b1
        self._wgrad(dx, bufs["emb_f"], self.lin_mix, splits, bias_done=True)
        if self.embed_kind == "sep_sub":
            df = self._gemm(dx, self.lin_mix.wt, G.EPI_RESIDUAL, aux=dx, out=bufs["emb_df"])
        else:
            df = self._gemm(dx, self.lin_mix.wt, G.EPI_STORE, out=bufs["emb_df"])
        du = (df.float() * _gelu_grad(bufs["emb_u"])).to(torch.bfloat16)
        self.gW1.zero_()
        self._wgrad_gemm(du, bufs["x_embed_in"], self.gW1, splits)
        db1 = du.float().sum(0)
        for w, b, r0, cols in self.emb_groups:            # the assembled matrix back to the parameters it was built from
            self.grad_of(w).add_(self.gW1[r0:r0 + w.shape[0]].index_select(1, cols))
            self.grad_of(b).add_(db1[r0:r0 + w.shape[0]])


# ------------------------------------------------------------------------------------------------
# autograd bridge + nn.Module
# ------------------------------------------------------------------------------------------------
class _ViTFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_in, engine, *params):
        ctx.engine = engine
        ctx.n = len(params)
        return engine.forward(x_in.detach())

    @staticmethod
    def backward(ctx, dlogits):
        eng = ctx.engine
        eng.backward(dlogits.contiguous().float())
        # clones: autograd may keep the returned tensors as .grad, and the flat buffer is rewritten every step
        grads = [eng.grad_views[k].clone() for k in eng.param_arg_order]
        return (None, None, *grads)


class ViT(nn.Module):
    """Same constructor surface as the reference ViT (plainvit.py:563-577); DCT pixel space, embed_type (`ver`) 1 or 2, with or
    without sub-block conversion; `ver=3` raises."""

    def __init__(self, in_channels: int = 3, patch_size: int = 16, emb_size: int = 768, input_embed: int = -1,
                 depth: int = 12, n_classes: int = 1000, drop_p=0.1, pixel_space="RGB", ver=1, use_subblock=True,
                 device="cpu", dtype=torch.float32, num_heads: int = 8, head_size: int = 64, **kwargs):
        super().__init__()
        if str(pixel_space).lower() not in ("dct", "rgb2dct") or patch_size != 16:
            raise NotImplementedError("rgbnm: only the DCT pixel space with patch_size=16 is on the B200 hot path (SURVEY.md 8f)")
        if ver not in (1, 2):
            # embed_type 3 (PatchEmbedding_DCT_Concat, plainvit.py:352-410) makes 294 tokens (196 luma + 2 x 49 chroma): a different
            # sequence length than the attention kernels are built for
            raise NotImplementedError("rgbnm: embed_type 3 (concatenated Y / CbCr token sequences) is outside the B200 hot path")
        if drop_p not in (0, 0.0):
            raise NotImplementedError("rgbnm: dropout > 0 is not on the hot path (reference default TRAIN.DROP = 0.0)")
        if input_embed not in (-1, emb_size) or num_heads * head_size != emb_size or head_size != 64:
            raise NotImplementedError("rgbnm: expects num_heads * 64 == emb_size (ViT-Ti/S/B)")
        if emb_size not in (192, 384, 768):
            raise NotImplementedError("rgbnm: emb_size must be 192, 384 or 768")
        if dtype != torch.float32:
            raise NotImplementedError("rgbnm: master weights are fp32 (reference CFG.MODEL.DTYPE = 'fp32')")
        self.emb_size, self.depth, self.n_classes = emb_size, depth, n_classes
        self.num_heads, self.head_size = num_heads, head_size
        self.pixel_space = pixel_space
        self.ver, self.use_subblock = ver, bool(use_subblock)
        self.embed_kind = "group" if ver == 1 else ("sep_sub" if use_subblock else "sep")
        E, HD = emb_size, num_heads * head_size
        dev = torch.device(device) if not isinstance(device, int) else torch.device("cuda", device)

        def lin(i, o):
            return nn.Linear(i, o, device=dev, dtype=torch.float32)
        self.patchembed = _Box()
        if self.embed_kind == "group":            # PatchEmbedding_DCT_Group (plainvit.py:173-216), with or without sub-block conversion
            self.patchembed.projection = nn.Sequential(lin(IN_FEAT, E))
        elif self.embed_kind == "sep_sub":        # PatchEmbedding_DCT_Separate_subblock (:285-350)
            self.patchembed.projection_Y = nn.Sequential(nn.Identity(), lin(256, E // 6 * 4))
            self.patchembed.projection_C = nn.Sequential(nn.Identity(), lin(128, E // 6 * 2))
            self.patchembed.linearMix = lin(E, E)
        else:                                     # PatchEmbedding_DCT_Separate (:220-283); `projection.1` is the same module as LinearMix
            self.patchembed.LinearY = nn.ModuleList([lin(64, E // 6) for _ in range(4)])
            self.patchembed.LinearC = nn.ModuleList([lin(64, E // 6) for _ in range(2)])
            self.patchembed.LinearMix = lin(E // 6 * 6, E)
            self.patchembed.projection = nn.Sequential(nn.Identity(), self.patchembed.LinearMix)
        blocks = []
        for _ in range(depth):
            mha = _Box()
            mha.qkv = lin(E, 3 * HD)
            mha.projection = lin(HD, E)
            r1 = _Box()
            r1.fn = nn.Sequential(OrderedDict([("eb_lrnorm1", nn.LayerNorm(E, device=dev)), ("eb_mha", mha)]))
            r2 = _Box()
            r2.fn = nn.Sequential(OrderedDict([
                ("eb_lrnorm2", nn.LayerNorm(E, device=dev)),
                ("eb_ffb", nn.Sequential(lin(E, 4 * E), nn.Identity(), nn.Identity(), lin(4 * E, E)))]))
            blocks.append(nn.Sequential(r1, r2))
        self.encoder = nn.Sequential(*blocks)
        self.classhead = nn.Sequential(OrderedDict([("ch_lrnorm", nn.LayerNorm(E, device=dev)),
                                                    ("ch_linear1", lin(E, E)), ("ch_linear2", lin(E, n_classes))]))
        self._engine: Optional[ViTEngine] = None
        self.attention_backend = kwargs.get("attention", "auto")

    # -- engine management ----------------------------------------------------------------------------
    def prepare(self, device=None) -> ViTEngine:
        """Pack the parameters into the flat fp32 buffer on `device` and build the engine.  Call after `.to(device)`
        / `load_state_dict` placement and before wrapping in DistributedDataParallel."""
        dev = torch.device(device) if device is not None else next(self.parameters()).device
        if self._engine is None or self._engine.dev != dev or any(
                p.data_ptr() != self._engine.flat.data_ptr() + 4 * o for p, o in self._engine_offsets()):
            self._engine = ViTEngine(self, dev, attention=self.attention_backend)
            self._engine.param_arg_order = [k for k, _ in self.named_parameters()]
        return self._engine

    def _engine_offsets(self):
        eng, off = self._engine, 0
        for k in eng.order:
            p = eng.params[k]
            yield p, off
            off += p.numel()

    def forward(self, x, cbcr=None):
        """forward(y, cbcr) with reference-format tensors (plainvit.py:601-612), or forward(x) with the
        (B,196,384) tensor FusedDCT writes."""
        if getattr(cbcr, "_rgbnm_absent", False):           # compat.loader.DCTBatch: chroma already merged by the fused kernel
            cbcr = None
        if cbcr is not None:
            x = embed_input_from_planes(x, cbcr, subblock=self.use_subblock)
        elif x.dim() != 3 or x.shape[1:] != (TOKENS, IN_FEAT):
            raise ValueError("rgbnm ViT: expected (y, cbcr) in the reference layout or a (B,196,384) embed input")
        eng = self.prepare(x.device)
        if eng.weights_stale():
            eng.refresh_weights()
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return _ViTFunction.apply(x, eng, *self.parameters())
        return eng.forward(x, save=False)
