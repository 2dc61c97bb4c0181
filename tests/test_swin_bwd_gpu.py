"""GPU parity tests of the SwinV2 backward kernels (training; first CUDA versions) against torch
autograd through the reference's formulas (oracle/swin_oracle.py), via the C-ABI.  Tolerance: bf16 outputs, fp32 arithmetic:
<= 2e-2 of each gradient's range; fp32 parameter gradients (dgamma, dbeta, dbias, dscale) <= 1e-2 relative to their norm."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import swin_oracle as SO
from rgb_no_more_b200 import lib as L

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-12)


@pytest.mark.parametrize("emb", [96, 192, 384, 768])
@pytest.mark.parametrize("scaled", [False, True])
def test_layernorm_res_scaled_fwd_bwd(emb, scaled):
    lib = L.load()
    B, rpi = 5, 64
    rows = B * rpi
    g = torch.Generator().manual_seed(emb + int(scaled))
    x = (torch.randn(rows, emb, generator=g) * 1.5 + 0.3).bfloat16()
    res = torch.randn(rows, emb, generator=g).bfloat16()
    dy = torch.randn(rows, emb, generator=g).bfloat16()
    gamma = 1 + 0.2 * torch.randn(emb, generator=g)
    beta = 0.2 * torch.randn(emb, generator=g)
    sc = torch.tensor([0.0, 1.25, 1.25, 0.0, 1.25]) if scaled else None
    # torch reference (fp32, autograd)
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    branch = F.layer_norm(xr, (emb,), gr, br, 1e-5)
    if scaled:
        branch = branch * sc.repeat_interleave(rpi).unsqueeze(1)
    yref = res.float() + branch
    yref.backward(dy.float())
    xd, rd, dyd = x.to(DEV), res.to(DEV), dy.to(DEV)
    gd, bd = gamma.to(DEV), beta.to(DEV)
    scd = sc.to(DEV) if scaled else None
    y = torch.empty_like(xd)
    L.check(lib.rgbnm_layernorm_res_scaled_fwd(xd.data_ptr(), gd.data_ptr(), bd.data_ptr(), rd.data_ptr(), None if scd is None else scd.data_ptr(),
                                               rpi, y.data_ptr(), rows, emb, 1e-5, L.stream_ptr()))
    assert _rel(y.float().cpu(), yref.detach()) < 1e-2
    dx = torch.empty_like(xd)
    dg, db = torch.zeros(emb, device=DEV), torch.zeros(emb, device=DEV)
    L.check(lib.rgbnm_layernorm_res_bwd(dyd.data_ptr(), xd.data_ptr(), gd.data_ptr(), None if scd is None else scd.data_ptr(), rpi,
                                        dx.data_ptr(), dg.data_ptr(), db.data_ptr(), rows, emb, 1e-5, L.stream_ptr()))
    assert _rel(dx.float().cpu(), xr.grad) < 2e-2
    assert _rel(dg.cpu(), gr.grad) < 1e-2 and _rel(db.cpu(), br.grad) < 1e-2


@pytest.mark.parametrize("H,heads,shift", [(16, 3, 0), (16, 3, 4), (8, 6, 0), (32, 3, 4)])
def test_window_attention_bwd_vs_autograd(H, heads, shift):
    lib = L.load()
    B, Cd = 2, heads * 32
    T = B * H * H
    g = torch.Generator().manual_seed(H + shift + heads)
    qkv = torch.randn(T, 3 * Cd, generator=g).bfloat16()
    dout = torch.randn(T, Cd, generator=g).bfloat16()
    bias = (16 * torch.sigmoid(torch.randn(heads, 64, 64, generator=g))).contiguous()
    scale = torch.exp(torch.randn(heads, generator=g) * 0.3 + 1.5)
    # reference: the forward sequence of swinv2.py on fp32 copies, differentiated by autograd
    qr = qkv.float().requires_grad_(True)
    br, sr = bias.clone().requires_grad_(True), scale.clone().requires_grad_(True)
    x = qr.view(B, H, H, 3 * Cd)
    if shift:
        x = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2))
    xw = SO.window_partition(x, 8).view(-1, 64, 3 * Cd)
    q, k, v = xw.reshape(-1, 64, 3, heads, 32).permute(2, 0, 3, 1, 4)
    attn = F.normalize(q, dim=-1) @ F.normalize(k, dim=-1).transpose(-2, -1)
    attn = attn * sr.view(1, heads, 1, 1) + br.unsqueeze(0)
    if shift:
        mask = SO.shift_mask(H, H, 8, shift)
        attn = (attn.view(-1, mask.shape[0], heads, 64, 64) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, 64, 64)
    o = (attn.softmax(-1) @ v).transpose(1, 2).reshape(-1, 8, 8, Cd)
    o = SO.window_reverse(o, 8, H, H)
    if shift:
        o = torch.roll(o, shifts=(shift, shift), dims=(1, 2))
    o.reshape(T, Cd).backward(dout.float())
    qd, dd, bd, sd = qkv.to(DEV), dout.to(DEV), bias.to(DEV), scale.to(DEV)
    dqkv = torch.zeros((T, 3 * Cd), dtype=torch.bfloat16, device=DEV)
    dbias = torch.zeros((heads, 64, 64), device=DEV)
    dscale = torch.zeros(heads, device=DEV)
    L.check(lib.rgbnm_window_attention_bwd(qd.data_ptr(), dd.data_ptr(), bd.data_ptr(), sd.data_ptr(), dqkv.data_ptr(), dbias.data_ptr(),
                                           dscale.data_ptr(), B, H, H, Cd, heads, 8, shift, L.stream_ptr()))
    torch.cuda.synchronize()
    got = dqkv.float().cpu()
    for name, sl in (("dq", slice(0, Cd)), ("dk", slice(Cd, 2 * Cd)), ("dv", slice(2 * Cd, 3 * Cd))):
        assert _rel(got[:, sl], qr.grad[:, sl]) < 2e-2, name
    assert _rel(dbias.cpu(), br.grad) < 1e-2
    assert _rel(dscale.cpu(), sr.grad) < 1e-2


def test_patch_merge_scatter_is_the_inverse_gather():
    lib = L.load()
    B, H, Cd = 3, 16, 96
    x = torch.randn(B, H, H, Cd).bfloat16().to(DEV)
    merged = torch.empty((B, H // 2, H // 2, 4 * Cd), dtype=torch.bfloat16, device=DEV)
    back = torch.empty_like(x)
    L.check(lib.rgbnm_patch_merge_gather(x.data_ptr(), merged.data_ptr(), B, H, H, Cd, L.stream_ptr()))
    L.check(lib.rgbnm_patch_merge_scatter(merged.data_ptr(), back.data_ptr(), B, H, H, Cd, L.stream_ptr()))
    assert torch.equal(back, x)
