"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not product code.

Functional restatement (plain torch fp32 on the CPU) of the reference model path
/root/reference/models/plainvit.py for pixel_space='DCT', ver=1, use_subblock=True, patch 16, driven by a
reference-layout `state_dict` (the 152 keys of `pvit.ViT.state_dict()`).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import it.

Pinned against outputs of the reference itself: tests/golden/embed_vit.npz (tools/make_golden.py runs the real
`pvit.ViT` with a seeded state_dict): embed input, tokens, block-0 output, logits, loss and gradients --
tests/test_oracle_golden.py::test_vit_oracle_matches_reference.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

from . import dct_oracle as O


def sincos(h: int, w: int, e: int) -> torch.Tensor:
    """SinCosEmbedding.forward (plainvit.py:97-121) for a (1, h, w, e) input."""
    ph = torch.arange(h)
    pw = torch.arange(w)
    hg, wg = torch.meshgrid(ph, pw, indexing="ij")
    freq = torch.log(torch.tensor(10000, dtype=torch.int32)) / (e // 4 - 1)
    freq = torch.exp(-torch.arange(e // 4, dtype=torch.float32) * freq)
    fh = torch.einsum("p,f->pf", hg.flatten().float(), freq)
    fw = torch.einsum("p,f->pf", wg.flatten().float(), freq)
    return torch.cat((fw.sin(), fw.cos(), fh.sin(), fh.cos()), dim=-1).view(1, h, w, e)


def tokens(sd: Dict[str, torch.Tensor], emb_in: torch.Tensor) -> torch.Tensor:
    """The patch embedding from the K0 operand on (everything after rearrange / sub-block conversion / collapse), chosen by the
    state_dict's key set.  emb_in: (B, 14, 14, 384) as produced by dct_oracle.embed_input(..., subblock=...).
      embed_type 1  PatchEmbedding_DCT_Group.projection (plainvit.py:194-198): Linear + sincos + 'b h w e -> b (h w) e'
      embed_type 2, sub-block  PatchEmbedding_DCT_Separate_subblock (:312-350): Linear(256) on Y | Linear(128) on CbCr -> GELU ->
                    linearMix + residual -> sincos
      embed_type 2, no sub-block  PatchEmbedding_DCT_Separate (:245-283): one Linear(64) per 8 x 8 block (4 luma, Cb, Cr) -> GELU ->
                    LinearMix -> sincos.  Its rearrange puts the blocks side by side ('(c pdh pdw) (p1 p2)'); emb_in holds the
                    un-converted 16 x 16 luma tile row-major, so block (pdh, pdw) is gathered from it here."""
    b, h, w, _ = emb_in.shape
    if "patchembed.projection.0.weight" in sd:
        x = F.linear(emb_in, sd["patchembed.projection.0.weight"], sd["patchembed.projection.0.bias"])
    elif "patchembed.projection_Y.1.weight" in sd:
        y = F.linear(emb_in[..., :256], sd["patchembed.projection_Y.1.weight"], sd["patchembed.projection_Y.1.bias"])
        c = F.linear(emb_in[..., 256:], sd["patchembed.projection_C.1.weight"], sd["patchembed.projection_C.1.bias"])
        t = F.gelu(torch.cat([y, c], dim=3))
        x = F.linear(t, sd["patchembed.linearMix.weight"], sd["patchembed.linearMix.bias"]) + t
    else:
        tile = emb_in[..., :256].reshape(b, h, w, 2, 8, 2, 8)                                    # (pdh p1) (pdw p2)
        blocks = tile.permute(0, 1, 2, 3, 5, 4, 6).reshape(b, h, w, 4, 64)                       # (pdh pdw) (p1 p2)
        outs = [F.linear(blocks[..., g, :], sd[f"patchembed.LinearY.{g}.weight"], sd[f"patchembed.LinearY.{g}.bias"]) for g in range(4)]
        outs += [F.linear(emb_in[..., 256 + 64 * ci:320 + 64 * ci], sd[f"patchembed.LinearC.{ci}.weight"],
                          sd[f"patchembed.LinearC.{ci}.bias"]) for ci in range(2)]
        x = F.linear(F.gelu(torch.cat(outs, dim=3)), sd["patchembed.LinearMix.weight"], sd["patchembed.LinearMix.bias"])
    e = x.shape[-1]
    return (x + sincos(h, w, e)).reshape(b, h * w, e)


def attention(sd, pre: str, x: torch.Tensor, heads: int, emb: int) -> torch.Tensor:
    """MultiHeadAttention.forward (plainvit.py:445-464), dropout 0."""
    b, n, _ = x.shape
    qkv = F.linear(x, sd[pre + "qkv.weight"], sd[pre + "qkv.bias"])
    d = qkv.shape[-1] // (3 * heads)
    qkv = qkv.view(b, n, heads, d, 3).permute(4, 0, 2, 1, 3)          # "b n (h d qkv) -> (qkv) b h n d"
    q, k, v = qkv[0], qkv[1], qkv[2]
    energy = torch.einsum("bhqd,bhkd->bhqk", q, k)
    att = F.softmax(energy / (emb ** 0.5), dim=-1)                    # scaled by sqrt(emb_size), :455-457
    out = torch.einsum("bhal,bhlv->bhav", att, v).permute(0, 2, 1, 3).reshape(b, n, heads * d)
    return F.linear(out, sd[pre + "projection.weight"], sd[pre + "projection.bias"])


def block(sd, l: int, x: torch.Tensor, heads: int, emb: int) -> torch.Tensor:
    """TransformerEncoderBlock (plainvit.py:493-529): two pre-LN residual branches."""
    p0, p1 = f"encoder.{l}.0.fn.", f"encoder.{l}.1.fn."
    h = F.layer_norm(x, (emb,), sd[p0 + "eb_lrnorm1.weight"], sd[p0 + "eb_lrnorm1.bias"], 1e-5)
    x = x + attention(sd, p0 + "eb_mha.", h, heads, emb)
    h = F.layer_norm(x, (emb,), sd[p1 + "eb_lrnorm2.weight"], sd[p1 + "eb_lrnorm2.bias"], 1e-5)
    h = F.gelu(F.linear(h, sd[p1 + "eb_ffb.0.weight"], sd[p1 + "eb_ffb.0.bias"]))      # nn.GELU(): exact erf
    return x + F.linear(h, sd[p1 + "eb_ffb.3.weight"], sd[p1 + "eb_ffb.3.bias"])


def head(sd, x: torch.Tensor, emb: int) -> torch.Tensor:
    """ClassificationHead (plainvit.py:547-557)."""
    x = F.layer_norm(x, (emb,), sd["classhead.ch_lrnorm.weight"], sd["classhead.ch_lrnorm.bias"], 1e-5).mean(dim=1)
    x = torch.tanh(F.linear(x, sd["classhead.ch_linear1.weight"], sd["classhead.ch_linear1.bias"]))
    return F.linear(x, sd["classhead.ch_linear2.weight"], sd["classhead.ch_linear2.bias"])


def forward(sd: Dict[str, torch.Tensor], yf: torch.Tensor, cf: torch.Tensor, depth: int = 12, upto_block=None,
            subblock: bool = True) -> torch.Tensor:
    """ViT.forward(y, cbcr) (plainvit.py:601-612) on ToRange'd planes (B,1,28,28,8,8) + (B,2,14,14,8,8)."""
    return forward_embedded(sd, O.embed_input(yf, cf, subblock), depth, upto_block)


def forward_embedded(sd, emb_in: torch.Tensor, depth: int = 12, upto_block=None) -> torch.Tensor:
    emb = sd["classhead.ch_lrnorm.weight"].shape[0]
    heads = emb // 64
    x = tokens(sd, emb_in.reshape(emb_in.shape[0], 14, 14, 384))
    for l in range(depth):
        if upto_block is not None and l == upto_block:
            return x
        x = block(sd, l, x, heads, emb)
    return head(sd, x, emb)


def train_step(sd: Dict[str, torch.Tensor], emb_in: torch.Tensor, soft_labels: torch.Tensor, state: dict, lr: float,
               wd: float, base_lr: float = 3e-3) -> float:
    """One optimisation step of train.py:153-176 on CPU fp32: CE(soft labels) -> backward -> clip_grad_norm_(1.0) ->
    AdamW(wd=0) -> decoupled WeightDecay on '.weight' params without 'lrnorm' (custom_optims.py:37-43)."""
    params = {k: v.requires_grad_(True) for k, v in sd.items()}
    logits = forward_embedded(params, emb_in)
    loss = torch.nn.CrossEntropyLoss()(logits, soft_labels)
    grads = torch.autograd.grad(loss, list(params.values()))
    total = torch.sqrt(sum((g * g).sum() for g in grads))
    clip = min(1.0, 1.0 / (float(total) + 1e-6))
    state["t"] = state.get("t", 0) + 1
    t = state["t"]
    with torch.no_grad():
        for (k, p), g in zip(params.items(), grads):
            g = g * clip
            m = state.setdefault("m:" + k, torch.zeros_like(p))
            v = state.setdefault("v:" + k, torch.zeros_like(p))
            m.mul_(0.9).add_(g, alpha=0.1)
            v.mul_(0.999).addcmul_(g, g, value=0.001)
            denom = (v.sqrt() / math.sqrt(1 - 0.999 ** t)).add_(1e-8)
            p.addcdiv_(m, denom, value=-lr / (1 - 0.9 ** t))
            if k.endswith(".weight") and "lrnorm" not in k:
                p.sub_(p * (lr / base_lr * wd))
    for v in sd.values():
        v.requires_grad_(False)
    return float(loss)
