"""SwinV2 DCT training path (training): explicit forward-with-saved-activations and backward on
the hand-written kernels, exposed to autograd through one `torch.autograd.Function` so that the reference's train loop
(`loss.backward()`, `clip_grad_norm_`, its optimiser; train.py:146-176) works on `swin.SwinTransformerV2` unchanged.

First correct version (round 1): the dense contractions run on the tcgen05 GEMM (forward, dgrad with the GELU' / residual
epilogues, split-K wgrad), LayerNorm / window-attention backward on the CUDA-core kernels of csrc/swin_bwd_kernels.cu;
the LayerNorm-in-epilogue fusions of the inference engine are not used here because the backward needs the pre-norm tensors.
Stochastic depth (timm DropPath, scale_by_keep; `drop_path_rate`, swinv2.py:648) is one Bernoulli draw per image and
branch, applied as the per-image scale of the post-norm residual kernel.  The tiny parameter-only graph behind the
attention tables (cpb_mlp -> 16 * sigmoid -> gather; exp(clamp(logit_scale)); swinv2.py:158-168) is differentiated by
torch from the kernels' d(bias tile) / d(scale).
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

from . import gemm as G
from . import lib as _lib
from . import ops as K

IN_FEAT = 24


class _Lin:
    """bf16 working copies of one Linear: w [out, in] (forward), wt [in, out] (dgrad)."""

    def __init__(self, weight: torch.Tensor, device):
        n, k = weight.shape
        self.w = torch.empty((n, k), dtype=torch.bfloat16, device=device)
        self.wt = torch.empty((k, n), dtype=torch.bfloat16, device=device)
        K.weight_prep(weight.detach().contiguous(), self.w, self.wt)


class SwinTrainEngine:
    def __init__(self, model, device: torch.device):
        if device.type != "cuda":
            raise _lib.RgbnmError("rgbnm: the SwinV2 training engine needs a CUDA device; there is no CPU fallback")
        self.model, self.device, self.L = model, device, _lib.load()
        depths = [len(layer.blocks) for layer in model.layers]
        self.dpr = [float(v) for v in torch.linspace(0, model.drop_path_rate, sum(depths))]      # swinv2.py:648
        self._lins: Dict[int, tuple] = {}
        self.static_weights = False      # SwinFlatEngine: working copies are refreshed by one table launch, never rebuilt here
        self.grads_static = None         # SwinFlatEngine: gradient views into its flat buffer (zeroed by the caller)

    # ---------------------------------------------------------------------------------------------
    def _lin(self, weight) -> "_Lin":
        """bf16 working copies of a Linear weight, rebuilt only when the parameter has been written (optimiser step, load)."""
        if not hasattr(self, "_lins"):
            self._lins = {}
        hit = self._lins.get(id(weight))
        if hit is not None and getattr(self, "static_weights", False):
            return hit[1]
        if hit is None or hit[0] != weight._version or hit[1].w.device != self.device:
            hit = (weight._version, _Lin(weight, self.device))
            self._lins[id(weight)] = hit
        return hit[1]

    def _draw(self, B: int, keep: float) -> torch.Tensor:
        """Per-image stochastic-depth scale mask / keep (timm DropPath, scale_by_keep), fp32 [B] on the device."""
        return torch.empty(B, device=self.device).bernoulli_(keep) / keep

    def _f32(self, t):
        return t.detach().to(self.device, torch.float32).contiguous()

    def _ln_fwd(self, x, norm, res, scale, rows_per_scale, y):
        rows, emb = x.shape
        _lib.check(self.L.rgbnm_layernorm_res_scaled_fwd(x.data_ptr(), norm[0].data_ptr(), norm[1].data_ptr(),
                                                         None if res is None else res.data_ptr(),
                                                         None if scale is None else scale.data_ptr(), rows_per_scale,
                                                         y.data_ptr(), rows, emb, 1e-5, _lib.stream_ptr()), "rgbnm_layernorm_res_scaled_fwd")
        return y

    def _ln_bwd(self, dy, x, gamma, scale, rows_per_scale, gname, bname, dxsum_name=None):
        """dxsum_name: bias of the Linear whose output `x` is -- its gradient (column sums of dx) is accumulated by the same kernel."""
        rows, emb = x.shape
        dx = torch.empty_like(x)
        _lib.check(self.L.rgbnm_layernorm_res_bwd_ex(dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), None if scale is None else scale.data_ptr(),
                                                     rows_per_scale, dx.data_ptr(), self.grads[gname].data_ptr(), self.grads[bname].data_ptr(),
                                                     None if dxsum_name is None else self.grads[dxsum_name].data_ptr(),
                                                     rows, emb, 1e-5, _lib.stream_ptr()), "rgbnm_layernorm_res_bwd_ex")
        return dx

    def _attn_fwd(self, qkv, bias, scale, B, H, Cd, heads, window, shift):
        att = torch.empty((qkv.shape[0], Cd), dtype=torch.bfloat16, device=qkv.device)
        _lib.check(self.L.rgbnm_window_attention_fwd(qkv.data_ptr(), att.data_ptr(), bias.data_ptr(), scale.data_ptr(), B, H, H, Cd, heads,
                                                     window, shift, _lib.stream_ptr()), "rgbnm_window_attention_fwd")
        return att

    def _attn_bwd(self, qkv, datt, bias, scale, dbias, dscale, B, H, Cd, heads, window, shift):
        dqkv = torch.empty_like(qkv)
        _lib.check(self.L.rgbnm_window_attention_bwd(qkv.data_ptr(), datt.data_ptr(), bias.data_ptr(), scale.data_ptr(), dqkv.data_ptr(),
                                                     dbias.data_ptr(), dscale.data_ptr(), B, H, H, Cd, heads, window, shift, _lib.stream_ptr()),
                   "rgbnm_window_attention_bwd")
        return dqkv, dbias, dscale

    def _gather(self, x, B, H, Cd):
        out = torch.empty((x.shape[0] // 4, 4 * Cd), dtype=torch.bfloat16, device=x.device)
        _lib.check(self.L.rgbnm_patch_merge_gather(x.data_ptr(), out.data_ptr(), B, H, H, Cd, _lib.stream_ptr()), "rgbnm_patch_merge_gather")
        return out

    def _scatter(self, dy, B, H, Cd):
        out = torch.empty((dy.shape[0] * 4, Cd), dtype=torch.bfloat16, device=dy.device)
        _lib.check(self.L.rgbnm_patch_merge_scatter(dy.data_ptr(), out.data_ptr(), B, H, H, Cd, _lib.stream_ptr()), "rgbnm_patch_merge_scatter")
        return out

    def _wgrad(self, dy, x, wname, bname=None):
        gw = self.grads[wname]
        if x.shape[1] > dy.shape[1]:
            G.gemm(x, dy, G.EPI_WGRAD_ATOMIC, out_f32=gw, trans_out=True)
        else:
            G.gemm(dy, x, G.EPI_WGRAD_ATOMIC, out_f32=gw)
        if bname is not None:
            K.colsum(dy, self.grads[bname])

    # ---------------------------------------------------------------------------------------------
    # The parameter-only graph behind the attention tables (swinv2.py:158-168): bias tile = 16 * sigmoid(cpb_mlp(coords table))
    # gathered by relative_position_index, logit scale = exp(min(logit_scale, log 100)).  Evaluated and differentiated for all blocks
    # of a (heads, window) group at once with batched matmuls (12 blocks x ~45 tiny launches each otherwise).
    def _table_groups(self):
        groups = {}
        for li, layer in enumerate(self.model.layers):
            for bi, blk in enumerate(layer.blocks):
                groups.setdefault((blk.attn.num_heads, blk.attn.window_size[0]), []).append((li, bi, blk.attn))
        return groups

    def _tables_forward(self, sv: dict) -> dict:
        out, sv["tables"] = {}, []
        cap = math.log(1.0 / 0.01)
        for (heads, ws), members in self._table_groups().items():
            a0 = members[0][2]
            nb, n = len(members), ws * ws
            coords = a0.relative_coords_table.reshape(-1, 2).float()                       # (R, 2), R = (2 ws - 1)^2
            idx = a0.relative_position_index.reshape(-1)                                   # (n * n,)
            W0 = torch.stack([a.cpb_mlp[0].weight.detach() for _, _, a in members])         # (nb, 512, 2)
            b0 = torch.stack([a.cpb_mlp[0].bias.detach() for _, _, a in members])           # (nb, 512)
            W2 = torch.stack([a.cpb_mlp[2].weight.detach() for _, _, a in members])         # (nb, heads, 512)
            ls = torch.stack([a.logit_scale.detach().reshape(heads) for _, _, a in members])   # (nb, heads)
            cx = coords.unsqueeze(0).expand(nb, -1, -1)
            pre = torch.baddbmm(b0.unsqueeze(1), cx, W0.transpose(1, 2))                    # (nb, R, 512)
            hh = torch.relu(pre)
            tt = torch.bmm(hh, W2.transpose(1, 2))                                          # (nb, R, heads)
            sg = torch.sigmoid(tt.index_select(1, idx))                                     # (nb, n * n, heads)
            bias = (16.0 * sg).permute(0, 2, 1).reshape(nb, heads, n, n).contiguous()
            scale = torch.clamp(ls, max=cap).exp().contiguous()
            dbias, dscale = torch.zeros_like(bias), torch.zeros_like(scale)
            for k, (li, bi, _) in enumerate(members):
                out[(li, bi)] = (bias[k], scale[k], dbias[k], dscale[k])
            sv["tables"].append(dict(members=members, heads=heads, n=n, coords=cx, idx=idx, W2=W2, ls=ls, pre=pre, hh=hh, sg=sg, scale=scale,
                                     dbias=dbias, dscale=dscale, cap=cap))
        return out

    def _tables_backward(self, sv: dict) -> None:
        gr = self.grads
        for t in sv["tables"]:
            members, heads, n = t["members"], t["heads"], t["n"]
            nb = len(members)
            sg = t["sg"]
            dg = t["dbias"].reshape(nb, heads, n * n).permute(0, 2, 1) * (16.0 * sg * (1.0 - sg))      # (nb, n * n, heads)
            dtt = torch.zeros((nb, t["coords"].shape[1], heads), dtype=torch.float32, device=dg.device).index_add_(1, t["idx"], dg)
            dW2 = torch.bmm(dtt.transpose(1, 2), t["hh"])                                    # (nb, heads, 512)
            dpre = torch.bmm(dtt, t["W2"]) * (t["pre"] > 0)                                  # (nb, R, 512)
            dW0 = torch.bmm(dpre.transpose(1, 2), t["coords"])                               # (nb, 512, 2)
            db0 = dpre.sum(1)
            dls = t["dscale"] * t["scale"] * (t["ls"] <= t["cap"])
            names = [f"layers.{li}.blocks.{bi}.attn" for li, bi, _ in members]
            torch._foreach_add_([gr[p + ".cpb_mlp.0.weight"] for p in names], list(dW0.unbind(0)))
            torch._foreach_add_([gr[p + ".cpb_mlp.0.bias"] for p in names], list(db0.unbind(0)))
            torch._foreach_add_([gr[p + ".cpb_mlp.2.weight"] for p in names], list(dW2.unbind(0)))
            torch._foreach_add_([gr[p + ".logit_scale"] for p in names], [g_.reshape(heads, 1, 1) for g_ in dls.unbind(0)])

    # ---------------------------------------------------------------------------------------------
    @_lib.no_autocast
    def forward(self, x_in: torch.Tensor) -> torch.Tensor:
        """x_in (B, 4096, 24) -> fp32 logits; keeps what the backward needs in self.saved."""
        m, dev = self.model, self.device
        x_in = x_in.detach().to(torch.bfloat16)
        B = x_in.shape[0]
        x_in = x_in.reshape(B * x_in.shape[1], IN_FEAT).contiguous()
        bf = lambda *s: torch.empty(s, dtype=torch.bfloat16, device=dev)
        sv: dict = {"B": B, "x_in": x_in, "stages": []}
        pe = m.patch_embed
        sv["embed"] = self._lin(pe.projection[0].weight)
        T0 = x_in.shape[0]
        e = G.gemm(x_in, sv["embed"].w, G.EPI_STORE, bias=self._f32(pe.projection[0].bias), out=bf(T0, m.embed_dim))
        sv["e"] = e
        x = self._ln_fwd(e, (self._f32(pe.norm.weight), self._f32(pe.norm.bias)), None, None, 1, bf(T0, m.embed_dim))
        blk_index = 0
        tables = self._tables_forward(sv)
        for li, layer in enumerate(m.layers):
            H, Cd = layer.input_resolution[0], layer.dim
            T = B * H * H
            st = {"H": H, "C": Cd, "blocks": []}
            for bi, blk in enumerate(layer.blocks):
                a = blk.attn
                heads = a.num_heads
                pfx = f"layers.{li}.blocks.{bi}"
                bias_d, scale_d, dbias_d, dscale_d = tables[(li, bi)]
                keep = 1.0 - self.dpr[blk_index]
                blk_index += 1
                s1 = s2 = None
                if m.training and keep < 1.0:
                    s1, s2 = self._draw(B, keep), self._draw(B, keep)
                lq, lp, l1, l2 = self._lin(a.qkv.weight), self._lin(a.proj.weight), self._lin(blk.mlp.fc1.weight), self._lin(blk.mlp.fc2.weight)
                qkv_bias = torch.cat((a.q_bias, torch.zeros_like(a.v_bias), a.v_bias)).detach().float().contiguous()
                qkv = G.gemm(x, lq.w, G.EPI_STORE, bias=qkv_bias, out=bf(T, 3 * Cd))
                att = self._attn_fwd(qkv, bias_d, scale_d, B, H, Cd, heads, a.window_size[0], blk.shift_size)
                p = G.gemm(att, lp.w, G.EPI_STORE, bias=self._f32(a.proj.bias), out=bf(T, Cd))
                g1 = self._f32(blk.norm1.weight)
                x1 = self._ln_fwd(p, (g1, self._f32(blk.norm1.bias)), x, s1, H * H, bf(T, Cd))
                u, f = G.gemm(x1, l1.w, G.EPI_GELU, bias=self._f32(blk.mlp.fc1.bias), out=bf(T, 4 * Cd), out2=bf(T, 4 * Cd))
                mm = G.gemm(f, l2.w, G.EPI_STORE, bias=self._f32(blk.mlp.fc2.bias), out=bf(T, Cd))
                g2 = self._f32(blk.norm2.weight)
                x2 = self._ln_fwd(mm, (g2, self._f32(blk.norm2.bias)), x1, s2, H * H, bf(T, Cd))
                st["blocks"].append(dict(pfx=pfx, attn=a, heads=heads, window=a.window_size[0], shift=blk.shift_size, x=x, qkv=qkv, att=att, p=p, x1=x1,
                                         u=u, f=f, m=mm, g1=g1, g2=g2, s1=s1, s2=s2, lq=lq, lp=lp, l1=l1, l2=l2,
                                         bias_d=bias_d, scale_d=scale_d, dbias_d=dbias_d, dscale_d=dscale_d))
                x = x2
            if layer.downsample is not None:
                ds = layer.downsample
                gath = self._gather(x, B, H, Cd)
                lr = self._lin(ds.reduction.weight)
                red = G.gemm(gath, lr.w, G.EPI_STORE, out=bf(T // 4, 2 * Cd))
                gd = self._f32(ds.norm.weight)
                x = self._ln_fwd(red, (gd, self._f32(ds.norm.bias)), None, None, 1, bf(T // 4, 2 * Cd))
                st["down"] = dict(gath=gath, red=red, lr=lr, g=gd)
            sv["stages"].append(st)
        last = m.layers[-1]
        Ltok, Cf = last.input_resolution[0] ** 2, m.num_features
        sv["x_last"] = x
        sv["gn"] = self._f32(m.norm.weight)
        xf = self._ln_fwd(x, (sv["gn"], self._f32(m.norm.bias)), None, None, 1, bf(B * Ltok, Cf))
        pooled = xf.view(B, Ltok, Cf).mean(dim=1, dtype=torch.float32)        # B x C work: torch (like the ViT head)
        sv["pooled"] = pooled
        self.saved = sv
        return F.linear(pooled, m.head.weight.detach().float(), m.head.bias.detach().float())

    # ---------------------------------------------------------------------------------------------
    @_lib.no_autocast
    def backward(self, dlogits: torch.Tensor) -> Dict[str, torch.Tensor]:
        m, dev, sv = self.model, self.device, self.saved
        B = sv["B"]
        self.grads = self.grads_static if getattr(self, "grads_static", None) is not None else \
            {k: torch.zeros_like(p, dtype=torch.float32, device=dev) for k, p in m.named_parameters()}
        gr = self.grads
        dlogits = dlogits.float()
        gr["head.weight"] += dlogits.t() @ sv["pooled"]
        gr["head.bias"] += dlogits.sum(0)
        dpooled = dlogits @ m.head.weight.detach().float()                   # (B, C)
        last = m.layers[-1]
        Ltok, Cf = last.input_resolution[0] ** 2, m.num_features
        dxf = (dpooled / Ltok).to(torch.bfloat16).unsqueeze(1).expand(B, Ltok, Cf).reshape(B * Ltok, Cf).contiguous()
        dx = self._ln_bwd(dxf, sv["x_last"], sv["gn"], None, 1, "norm.weight", "norm.bias")
        for li in range(len(m.layers) - 1, -1, -1):
            st = sv["stages"][li]
            H, Cd = st["H"], st["C"]
            T = B * H * H
            if "down" in st:
                d = st["down"]
                pfx = f"layers.{li}.downsample"
                dred = self._ln_bwd(dx, d["red"], d["g"], None, 1, pfx + ".norm.weight", pfx + ".norm.bias")
                self._wgrad(dred, d["gath"], pfx + ".reduction.weight")
                dgath = G.gemm(dred, d["lr"].wt, G.EPI_STORE)
                dx = self._scatter(dgath, B, H, Cd)
            for b in reversed(st["blocks"]):
                pfx = b["pfx"]
                # x2 = x1 + s2 * norm2(m)
                dm = self._ln_bwd(dx, b["m"], b["g2"], b["s2"], H * H, pfx + ".norm2.weight", pfx + ".norm2.bias", pfx + ".mlp.fc2.bias")
                self._wgrad(dm, b["f"], pfx + ".mlp.fc2.weight")
                du = G.gemm(dm, b["l2"].wt, G.EPI_DGELU, aux=b["u"])                                   # (dm W2) * gelu'(u)
                self._wgrad(du, b["x1"], pfx + ".mlp.fc1.weight", pfx + ".mlp.fc1.bias")
                dx1 = G.gemm(du, b["l1"].wt, G.EPI_RESIDUAL, aux=dx)                                   # + residual path
                # x1 = x + s1 * norm1(p)
                dp = self._ln_bwd(dx1, b["p"], b["g1"], b["s1"], H * H, pfx + ".norm1.weight", pfx + ".norm1.bias", pfx + ".attn.proj.bias")
                self._wgrad(dp, b["att"], pfx + ".attn.proj.weight")
                datt = G.gemm(dp, b["lp"].wt, G.EPI_STORE)
                dqkv, _, _ = self._attn_bwd(b["qkv"], datt, b["bias_d"], b["scale_d"], b["dbias_d"], b["dscale_d"], B, H, Cd, b["heads"],
                                            b["window"], b["shift"])
                self._wgrad(dqkv, b["x"], pfx + ".attn.qkv.weight")
                K.colsum(dqkv[:, :Cd], gr[pfx + ".attn.q_bias"])
                K.colsum(dqkv[:, 2 * Cd:], gr[pfx + ".attn.v_bias"])
                dx = G.gemm(dqkv, b["lq"].wt, G.EPI_RESIDUAL, aux=dx1)
        self._tables_backward(sv)
        # patch embedding: x0 = norm(e), e = x_in W^T + b
        de = self._ln_bwd(dx, sv["e"], self._f32(m.patch_embed.norm.weight), None, 1, "patch_embed.norm.weight", "patch_embed.norm.bias",
                          "patch_embed.projection.0.bias")
        self._wgrad(de, sv["x_in"], "patch_embed.projection.0.weight")
        self.saved = None
        return gr


class SwinFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_in, engine, names, *params):
        ctx.engine, ctx.names = engine, names
        return engine.forward(x_in)

    @staticmethod
    def backward(ctx, dlogits):
        grads = ctx.engine.backward(dlogits.contiguous())
        return (None, None, None, *[grads[n].to(torch.float32) for n in ctx.names])


class SwinFlatEngine:
    """SwinTrainEngine behind the interface train_step.TrainStage drives (flat fp32 parameter / gradient buffers, forward,
    backward, refresh_weights): the parameters become views into ONE flat buffer -- first the group the reference's WeightDecay
    optimiser acts on ('.weight' in the name, pipeline_utils.py:537), then the rest -- so that the step is one flat all-reduce, one
    clip + AdamW + decay kernel and one weight-refresh launch, and forward + backward can be captured in a CUDA graph."""
    ALIGN = 8          # elements: every parameter starts on a 32-byte boundary (vector reds of the wgrad epilogue)

    def __init__(self, model, device: torch.device):
        import numpy as np
        self.model, self.dev = model, torch.device(device)
        self.inner = SwinTrainEngine(model, self.dev)
        named = list(model.named_parameters())
        decay = [(k, p) for k, p in named if ".weight" in k]
        rest = [(k, p) for k, p in named if ".weight" not in k]
        pad = lambda n: -(-n // self.ALIGN) * self.ALIGN
        self.n_decay = sum(pad(p.numel()) for _, p in decay)
        total = self.n_decay + sum(pad(p.numel()) for _, p in rest)
        self.flat = torch.zeros(total, dtype=torch.float32, device=self.dev)
        self.flat_grad = torch.zeros_like(self.flat)
        grads, off = {}, 0
        with torch.no_grad():
            for k, p in decay + rest:
                n = p.numel()
                self.flat[off:off + n].copy_(p.data.reshape(-1))
                p.data = self.flat[off:off + n].view(p.shape)
                grads[k] = self.flat_grad[off:off + n].view(p.shape)
                off += pad(n)
        self.inner.grads_static = grads
        # bf16 working copies of every Linear the engine multiplies with, refreshed by one launch over a descriptor table
        lins = [model.patch_embed.projection[0].weight]
        for layer in model.layers:
            for blk in layer.blocks:
                lins += [blk.attn.qkv.weight, blk.attn.proj.weight, blk.mlp.fc1.weight, blk.mlp.fc2.weight]
            if layer.downsample is not None:
                lins.append(layer.downsample.reduction.weight)
        arr = (_lib.WPrepDesc * len(lins))()
        tiles = 0
        for i, w in enumerate(lins):
            lin = _Lin(w, self.dev)
            self.inner._lins[id(w)] = (None, lin)
            d = arr[i]
            d.w, d.w_bf16, d.wt_bf16, d.bias, d.bias_k = w.data_ptr(), lin.w.data_ptr(), lin.wt.data_ptr(), None, None
            d.n, d.k, d.qkv_heads, d.head_dim, d.first_tile = w.shape[0], w.shape[1], 0, 0, tiles
            tiles += -(-w.shape[0] // 32) * -(-w.shape[1] // 32)
        self.inner.static_weights = True
        self._wprep = (torch.from_numpy(np.frombuffer(bytes(arr), dtype=np.uint8).copy()).to(self.dev), len(lins), tiles)
        self.launches = 0

    def forward(self, x_in: torch.Tensor) -> torch.Tensor:
        return self.inner.forward(x_in)

    def backward(self, dlogits: torch.Tensor) -> None:
        self.flat_grad.zero_()
        self.inner.backward(dlogits)

    def refresh_weights(self) -> None:
        table, n, tiles = self._wprep
        _lib.check(self.inner.L.rgbnm_weight_prep_batch(table.data_ptr(), n, tiles, _lib.stream_ptr()), "rgbnm_weight_prep_batch")
