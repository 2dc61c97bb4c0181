"""GPU parity tests of the SwinV2 backward kernels (training; first CUDA versions) against torch
autograd through the reference's formulas (oracle/swin_oracle.py), via the C-ABI.  Tolerance: bf16 outputs, fp32 arithmetic:
<= 2e-2 of each gradient's range; fp32 parameter gradients (dgamma, dbeta, dbias, dscale) <= 1e-2 relative to their norm."""
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
    # the _ex entry also accumulates the column sums of the dx it writes (bias gradient of the Linear in front of the LayerNorm)
    dx2 = torch.empty_like(xd)
    dg2, db2, dxs = torch.zeros(emb, device=DEV), torch.zeros(emb, device=DEV), torch.ones(emb, device=DEV)
    L.check(lib.rgbnm_layernorm_res_bwd_ex(dyd.data_ptr(), xd.data_ptr(), gd.data_ptr(), None if scd is None else scd.data_ptr(), rpi,
                                           dx2.data_ptr(), dg2.data_ptr(), db2.data_ptr(), dxs.data_ptr(), rows, emb, 1e-5, L.stream_ptr()))
    assert torch.equal(dx2, dx)
    want = 1.0 + dx.float().sum(0)                                     # accumulated on top of what the buffer held
    assert float((dxs - want).abs().max()) < 1e-3 * max(1.0, float(want.abs().max()))


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


def test_swin_training_gradients_match_reference():
    """SwinV2-T DCT in train() mode on the GPU (stochastic depth off): loss and gradients against the reference's own
    (tests/golden/swin_train.npz, written by tools/make_golden_swin.py from /root/reference).  bf16 operands, fp32
    accumulation: loss within 2 %, gradient direction cos > 0.99 and norm within 5 % for ten parameters of every kind."""
    from rgb_no_more_b200 import swin as S
    from tests.helpers import load, seeded_swin_state_dict, golden_swin_inputs
    g = load("swin_train.npz")
    m = S.SwinTransformerV2(img_size=256, patch_size=4, embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=8,
                            mlp_ratio=4, drop_rate=0, attn_drop_rate=0, drop_path_rate=0.0, qkv_bias=True, ape=False, patch_norm=True,
                            pretrained_window_sizes=[0, 0, 0, 0], device="cpu", pixel_space="dct")
    m.load_state_dict(seeded_swin_state_dict(m))
    m.train().to(DEV)
    yf, cf = golden_swin_inputs(g["input_seed"])
    labels = torch.zeros((2, 1000))
    labels[0, 3], labels[0, 7], labels[1, 999] = 0.7, 0.3, 1.0
    logits = m(yf.to(DEV), cf.to(DEV))
    loss = torch.nn.CrossEntropyLoss()(logits, labels.to(DEV))
    loss.backward()
    torch.cuda.synchronize()
    assert abs(float(loss) - float(g["loss"])) < 2e-2 * float(g["loss"])
    named = dict(m.named_parameters())
    for k in [k[5:] for k in g.files if k.startswith("grad:")]:
        got = named[k].grad.float().cpu()
        ref = torch.from_numpy(g["grad:" + k])
        cos = float(F.cosine_similarity(got.reshape(-1)[:4096], ref, dim=0))
        assert cos > 0.99, (k, cos)
        assert abs(float(got.norm()) - float(g["gradnorm:" + k])) < 5e-2 * float(g["gradnorm:" + k]), (k, float(got.norm()), float(g["gradnorm:" + k]))


def test_swin_training_step_with_stochastic_depth_learns():
    """Two-stage SwinV2 with drop_path 0.2 in train() mode: a few AdamW steps on one batch reduce the loss (the whole
    forward / backward incl. the per-image branch scales runs; gradients are finite)."""
    from rgb_no_more_b200 import swin as S
    torch.manual_seed(3)
    m = S.SwinTransformerV2(img_size=256, patch_size=4, embed_dim=96, depths=[2, 2], num_heads=[3, 6], window_size=8,
                            drop_path_rate=0.2, pretrained_window_sizes=[0, 0], device="cpu", pixel_space="dct")
    with torch.no_grad():
        for p in m.parameters():
            if p.ndim == 1:
                p.add_(0.3 * torch.randn_like(p))          # the reference zero-initialises the block post-norms
    m.train().to(DEV)
    x = (torch.randn(4, 4096, 24) * 0.5).to(DEV)
    y = torch.tensor([1, 5, 9, 2], device=DEV)
    opt = torch.optim.AdamW(m.parameters(), lr=2e-3)
    losses = []
    for _ in range(6):
        opt.zero_grad()
        loss = F.cross_entropy(m(x), y)
        loss.backward()
        assert all(torch.isfinite(p.grad).all() for p in m.parameters())
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0], losses


def test_swin_flat_engine_matches_the_autograd_path():
    """swin_train.SwinFlatEngine (parameters as views of one flat buffer, gradients written into its flat twin, working copies
    refreshed by one table launch) against the same model through the autograd.Function path: identical logits, gradients equal
    up to the atomics' summation order; after an in-place update of the flat buffer + refresh_weights() the logits follow."""
    from rgb_no_more_b200 import swin as S, swin_train as ST
    from tests.helpers import seeded_swin_state_dict

    def build():
        m = S.SwinTransformerV2(img_size=256, patch_size=4, embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=8,
                                drop_path_rate=0.0, pretrained_window_sizes=[0, 0, 0, 0], device="cpu", pixel_space="dct")
        m.load_state_dict(seeded_swin_state_dict(m))
        return m.train().to(DEV)
    a, b = build(), build()
    x = (torch.randn(2, 4096, 24, generator=torch.Generator().manual_seed(4)) * 0.5).to(DEV).to(torch.bfloat16)
    y = torch.tensor([7, 123], device=DEV)
    loss = F.cross_entropy(a(x), y)
    loss.backward()
    eng = ST.SwinFlatEngine(b, torch.device(DEV))
    assert eng.n_decay % 8 == 0 and all(p.data_ptr() % 32 == 0 for p in b.parameters())
    logits = eng.forward(x)
    assert torch.equal(logits, a(x).detach())
    dlogits = (torch.softmax(logits, 1) - F.one_hot(y, 1000).float()) / 2
    eng.backward(dlogits)
    torch.cuda.synchronize()
    for (k, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        gq = eng.inner.grads_static[k]
        assert _rel(gq, p.grad) < 2e-3, k
    with torch.no_grad():
        eng.flat.mul_(1.01)
        for p, q in zip(a.parameters(), b.parameters()):
            p.mul_(1.01)
    eng.refresh_weights()
    assert float((eng.forward(x) - a(x).detach()).abs().max()) < 1e-3 * float(logits.abs().max())


@pytest.mark.parametrize("use_graph", [True, False])
def test_swin_train_stage_learns(use_graph):
    """train_step.TrainStage(arch='swinv2t'): mixup -> SwinV2-T forward / backward (stochastic depth 0.2) -> clip + AdamW + decoupled
    decay kernel -> working-copy refresh, as two CUDA graphs or eagerly: finite, decreasing loss on one repeated batch."""
    from rgb_no_more_b200 import train_step as TS
    st = TS.TrainStage(DEV, arch="swinv2t", batch=4, lr=1e-3, warmup_steps=2, total_steps=100, mixup_alpha=0.0, use_graph=use_graph)
    with torch.no_grad():                                   # the reference zero-initialises the block post-norms
        st.eng.flat.add_(0.05 * torch.randn_like(st.eng.flat) * (st.eng.flat == 0))
    st.eng.refresh_weights()
    x = (torch.randn(4, 4096, 24, generator=torch.Generator().manual_seed(1)) * 0.5).to(DEV).to(torch.bfloat16)
    y = torch.tensor([1, 5, 9, 2], device=DEV)
    losses = [float(st.step(x, y)) for _ in range(12)]
    assert all(l == l for l in losses), losses
    assert losses[-1] < 0.7 * losses[0], losses
