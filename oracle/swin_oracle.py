"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not product code.

Functional fp32 restatement of the reference SwinV2 DCT forward (models/swinv2.py), driven by a
reference-format `state_dict`.  Pinned against outputs of the reference itself (tests/golden/swin_model.npz, written by
tools/make_golden_swin.py from /root/reference): tests/test_swin_oracle_golden.py.  Only tests/, smoke() and bench.py's
CPU legs may import this module.

Each function cites the reference file:line it follows.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F

from . import dct_oracle as O


def _ln(x, sd, prefix, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


def window_partition(x, ws):
    # swinv2.py:39-51
    B, H, W, C = x.shape
    x = x.view(B, H // ws, ws, W // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, C)


def window_reverse(windows, ws, H, W):
    # swinv2.py:54-66
    B = int(windows.shape[0] / (H * W / ws / ws))
    x = windows.view(B, H // ws, W // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(B, H, W, -1)


def relative_tables(ws: int):
    """relative_coords_table and relative_position_index of WindowAttention.__init__ (swinv2.py:98-125),
    pretrained_window_size = 0."""
    rh = torch.arange(-(ws - 1), ws, dtype=torch.float32)
    table = torch.stack(torch.meshgrid([rh, rh], indexing="ij")).permute(1, 2, 0).contiguous().unsqueeze(0)
    table = table / (ws - 1)
    table = table * 8
    table = torch.sign(table) * torch.log2(torch.abs(table) + 1.0) / math.log2(8)
    coords = torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing="ij"))
    cf = torch.flatten(coords, 1)
    rel = (cf[:, :, None] - cf[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return table, rel.sum(-1)


def shift_mask(H: int, W: int, ws: int, shift: int):
    # swinv2.py:224-242
    img_mask = torch.zeros((1, H, W, 1))
    cnt = 0
    for h in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for w in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img_mask[:, h, w, :] = cnt
            cnt += 1
    mw = window_partition(img_mask, ws).view(-1, ws * ws)
    am = mw.unsqueeze(1) - mw.unsqueeze(2)
    return am.masked_fill(am != 0, float(-100.0)).masked_fill(am == 0, float(0.0))


def attention_bias(sd, prefix: str, heads: int, ws: int):
    """16 * sigmoid(cpb_mlp(table))[index] -> (heads, ws^2, ws^2), and the clamped logit scale (heads,)
    (swinv2.py:158-168)."""
    table = sd.get(prefix + ".relative_coords_table")
    index = sd.get(prefix + ".relative_position_index")
    if table is None or index is None:
        table, index = relative_tables(ws)
    h = F.relu(F.linear(table, sd[prefix + ".cpb_mlp.0.weight"], sd[prefix + ".cpb_mlp.0.bias"]))
    t = F.linear(h, sd[prefix + ".cpb_mlp.2.weight"]).view(-1, heads)
    rpb = t[index.view(-1)].view(ws * ws, ws * ws, -1).permute(2, 0, 1).contiguous()
    scale = torch.clamp(sd[prefix + ".logit_scale"], max=math.log(1.0 / 0.01)).exp().view(heads)
    return 16 * torch.sigmoid(rpb), scale


def window_attention(xw, sd, prefix, heads, ws, mask):
    # WindowAttention.forward, swinv2.py:143-182
    B_, N, C = xw.shape
    qkv_bias = torch.cat((sd[prefix + ".q_bias"], torch.zeros_like(sd[prefix + ".v_bias"]), sd[prefix + ".v_bias"]))
    qkv = F.linear(xw, sd[prefix + ".qkv.weight"], qkv_bias).reshape(B_, N, 3, heads, -1).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = F.normalize(q, dim=-1) @ F.normalize(k, dim=-1).transpose(-2, -1)
    bias, scale = attention_bias(sd, prefix, heads, ws)
    attn = attn * scale.view(1, heads, 1, 1) + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = attn.view(B_ // nW, nW, heads, N, N) + mask.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, heads, N, N)
    attn = attn.softmax(dim=-1)
    x = (attn @ v).transpose(1, 2).reshape(B_, N, C)
    return F.linear(x, sd[prefix + ".proj.weight"], sd[prefix + ".proj.bias"])


def block(x, sd, prefix, H, W, heads, ws, shift, s1=None, s2=None):
    # SwinTransformerBlock.forward, swinv2.py:270-309.  s1 / s2: per-image stochastic-depth scales (mask / keep_prob, timm
    # DropPath with scale_by_keep) of the two branches, shape (B,); None = identity (eval, or drop_path 0)
    B, L, C = x.shape
    shortcut = x
    x = x.view(B, H, W, C)
    if shift > 0:
        x = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2))
    xw = window_partition(x, ws).view(-1, ws * ws, C)
    mask = shift_mask(H, W, ws, shift) if shift > 0 else None
    aw = window_attention(xw, sd, prefix + ".attn", heads, ws, mask).view(-1, ws, ws, C)
    x = window_reverse(aw, ws, H, W)
    if shift > 0:
        x = torch.roll(x, shifts=(shift, shift), dims=(1, 2))
    x = x.view(B, H * W, C)
    a = _ln(x, sd, prefix + ".norm1")
    x = shortcut + (a if s1 is None else a * s1.view(-1, 1, 1))
    m = F.linear(F.gelu(F.linear(x, sd[prefix + ".mlp.fc1.weight"], sd[prefix + ".mlp.fc1.bias"])),
                 sd[prefix + ".mlp.fc2.weight"], sd[prefix + ".mlp.fc2.bias"])
    m = _ln(m, sd, prefix + ".norm2")
    return x + (m if s2 is None else m * s2.view(-1, 1, 1))


def patch_merging(x, sd, prefix, H, W):
    # PatchMerging.forward, swinv2.py:340-363
    B, L, C = x.shape
    x = x.view(B, H, W, C)
    x = torch.cat([x[:, 0::2, 0::2, :], x[:, 1::2, 0::2, :], x[:, 0::2, 1::2, :], x[:, 1::2, 1::2, :]], -1).view(B, -1, 4 * C)
    return _ln(F.linear(x, sd[prefix + ".reduction.weight"]), sd, prefix + ".norm")


def tokens(sd, emb_in):
    """Linear(24, E) + LayerNorm on the (B, 64, 64, 24) embed input (swinv2.py:566-570)."""
    x = F.linear(emb_in, sd["patch_embed.projection.0.weight"], sd["patch_embed.projection.0.bias"])
    x = x.reshape(x.shape[0], -1, x.shape[-1])
    return _ln(x, sd, "patch_embed.norm")


def forward_from_embed(sd: Dict[str, torch.Tensor], emb_in: torch.Tensor, depths: Sequence[int] = (2, 2, 6, 2),
                       heads: Sequence[int] = (3, 6, 12, 24), window: int = 8, collect: List = None,
                       drop_scales: Sequence = None) -> torch.Tensor:
    """SwinTransformerV2.forward (swinv2.py:681-705) from the embed input (B, R, R, 24) to the logits."""
    res = emb_in.shape[1]
    x = tokens(sd, emb_in)
    for li, depth in enumerate(depths):
        H = W = res // 2 ** li
        ws = min(window, H)                                            # swinv2.py:208-211
        for bi in range(depth):
            shift = 0 if (bi % 2 == 0 or H <= window) else window // 2
            s1 = s2 = None
            if drop_scales is not None:                                # one (s1, s2) pair per block, in forward order
                s1, s2 = drop_scales[sum(depths[:li]) + bi]
            x = block(x, sd, f"layers.{li}.blocks.{bi}", H, W, heads[li], ws, shift, s1, s2)
            if collect is not None:
                collect.append((f"l{li}b{bi}", x))
        if li < len(depths) - 1:
            x = patch_merging(x, sd, f"layers.{li}.downsample", H, W)
        if collect is not None:
            collect.append((f"stage{li}", x))
    x = _ln(x, sd, "norm").mean(dim=1)                                 # AdaptiveAvgPool1d(1) over tokens
    return F.linear(x, sd["head.weight"], sd["head.bias"])


def forward(sd, yf, cf, **kw) -> torch.Tensor:
    """forward(y, cbcr) with ToRange'd planes (B,1,32,32,8,8) / (B,2,16,16,8,8)."""
    return forward_from_embed(sd, O.embed_input_swin(yf, cf), **kw)
