"""GPU parity tests of the SwinV2 DCT forward path, through the C-ABI.

Tolerances (bf16 activations and weights, fp32 accumulation / statistics -- the reference's `--amp` regime -- against
the fp32 reference): single kernels <= 1e-2 of the output range (bf16 output rounding); whole model: logits within
3e-2 of the logit range of the reference's own outputs (tests/golden/swin_model.npz) with the same argmax."""

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import dct_oracle as O
from oracle import swin_oracle as SO
from rgb_no_more_b200 import lib as L
from rgb_no_more_b200 import swin as S
from rgb_no_more_b200 import transforms as TF
from rgb_no_more_b200 import plan as P
from rgb_no_more_b200 import synth
from tests.helpers import load, seeded_swin_state_dict, golden_swin_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _model():
    m = S.SwinTransformerV2(img_size=256, patch_size=4, embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24],
                            window_size=8, mlp_ratio=4, drop_rate=0, attn_drop_rate=0, drop_path_rate=0.2, qkv_bias=True,
                            ape=False, patch_norm=True, pretrained_window_sizes=[0, 0, 0, 0], device="cpu", pixel_space="dct")
    sd = seeded_swin_state_dict(m)
    m.load_state_dict(sd)
    return m.eval(), sd


@pytest.mark.parametrize("emb", [96, 192, 384, 768, 1536])
@pytest.mark.parametrize("with_res", [False, True])
def test_layernorm_res_fwd(emb, with_res):
    lib = L.load()
    rows = 1000
    g = torch.Generator().manual_seed(emb)
    x = (torch.randn(rows, emb, generator=g) * 2 + 0.5).bfloat16().to(DEV)
    res = torch.randn(rows, emb, generator=g).bfloat16().to(DEV) if with_res else None
    gamma = (1 + 0.1 * torch.randn(emb, generator=g)).to(DEV)
    beta = (0.1 * torch.randn(emb, generator=g)).to(DEV)
    y = torch.empty_like(x)
    L.check(lib.rgbnm_layernorm_res_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), None if res is None else res.data_ptr(),
                                        y.data_ptr(), rows, emb, 1e-5, L.stream_ptr()))
    ref = F.layer_norm(x.float(), (emb,), gamma, beta, 1e-5)
    if with_res:
        ref = ref + res.float()
    assert float((y.float() - ref).abs().max()) < 1e-2 * float(ref.abs().max())


@pytest.mark.parametrize("H,heads,shift", [(64, 3, 0), (64, 3, 4), (32, 6, 4), (16, 12, 4), (8, 24, 0)])
def test_window_attention_vs_oracle(H, heads, shift):
    """The kernel (gather by window + shift, cosine attention, bias, mask, scatter back) against the reference's
    roll / window_partition / attention / window_reverse / roll sequence restated in oracle/swin_oracle.py."""
    lib = L.load()
    B, Cd = 2, heads * 32
    g = torch.Generator().manual_seed(H + shift)
    qkv = torch.randn(B * H * H, 3 * Cd, generator=g).bfloat16()
    bias = (16 * torch.sigmoid(torch.randn(heads, 64, 64, generator=g))).contiguous()
    scale = torch.exp(torch.randn(heads, generator=g) * 0.5 + 2.0)
    out = torch.empty((B * H * H, Cd), dtype=torch.bfloat16, device=DEV)
    qd, bd, sd_ = qkv.to(DEV), bias.to(DEV), scale.to(DEV)
    L.check(lib.rgbnm_window_attention_fwd(qd.data_ptr(), out.data_ptr(), bd.data_ptr(), sd_.data_ptr(), B, H, H, Cd, heads, 8,
                                           shift, L.stream_ptr()))
    torch.cuda.synchronize()
    # reference sequence on the CPU, fp32, from the same bf16 qkv values
    x = qkv.float().view(B, H, H, 3 * Cd)
    if shift:
        x = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2))
    xw = SO.window_partition(x, 8).view(-1, 64, 3 * Cd)
    q, k, v = xw.reshape(-1, 64, 3, heads, 32).permute(2, 0, 3, 1, 4)
    attn = F.normalize(q, dim=-1) @ F.normalize(k, dim=-1).transpose(-2, -1)
    attn = attn * scale.view(1, heads, 1, 1) + bias.unsqueeze(0)
    if shift:
        mask = SO.shift_mask(H, H, 8, shift)
        nW = mask.shape[0]
        attn = (attn.view(-1, nW, heads, 64, 64) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, 64, 64)
    o = (attn.softmax(-1) @ v).transpose(1, 2).reshape(-1, 8, 8, Cd)
    o = SO.window_reverse(o, 8, H, H)
    if shift:
        o = torch.roll(o, shifts=(shift, shift), dims=(1, 2))
    ref = o.reshape(B * H * H, Cd)
    assert float((out.float().cpu() - ref).abs().max()) < 1e-2 * float(ref.abs().max())


def test_window_attention_rejects_other_geometries():
    lib = L.load()
    t = torch.zeros(64 * 64 * 2, 288, dtype=torch.bfloat16, device=DEV)
    f = torch.zeros(3 * 49 * 49, device=DEV)
    assert lib.rgbnm_window_attention_fwd(t.data_ptr(), t.data_ptr(), f.data_ptr(), f.data_ptr(), 1, 56, 56, 96, 3, 7, 0, L.stream_ptr()) != 0
    assert lib.rgbnm_window_attention_fwd(t.data_ptr(), t.data_ptr(), f.data_ptr(), f.data_ptr(), 1, 64, 64, 96, 2, 8, 0, L.stream_ptr()) != 0


def test_patch_merge_and_token_mean():
    lib = L.load()
    B, H, Cd = 3, 16, 192
    x = torch.randn(B, H, H, Cd).bfloat16().to(DEV)
    out = torch.empty((B, H // 2, H // 2, 4 * Cd), dtype=torch.bfloat16, device=DEV)
    L.check(lib.rgbnm_patch_merge_gather(x.data_ptr(), out.data_ptr(), B, H, H, Cd, L.stream_ptr()))
    ref = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1)
    assert torch.equal(out, ref)                                      # pure index permutation: bit exact
    m = torch.empty((B, Cd), dtype=torch.bfloat16, device=DEV)
    L.check(lib.rgbnm_token_mean_bf16(x.data_ptr(), m.data_ptr(), B, H * H, Cd, L.stream_ptr()))
    refm = x.float().view(B, H * H, Cd).mean(1)
    assert float((m.float() - refm).abs().max()) < 2e-3


def test_swin_forward_matches_reference_golden():
    """Whole SwinV2-T DCT forward on the GPU against the outputs of the reference's own swinv2.py (golden) --
    via forward(y, cbcr) in the reference's input format -- stage by stage and at the logits."""
    g = load("swin_model.npz")
    m, sd = _model()
    m.to(DEV)
    yf, cf = golden_swin_inputs(g["input_seed"])
    eng = m.prepare(DEV)
    acts = []
    with torch.no_grad():
        x_in = S.swin_embed_input_from_planes(yf.to(DEV), cf.to(DEV))
        logits = eng.forward(x_in, collect=acts).cpu()
        logits2 = m(yf.to(DEV), cf.to(DEV)).cpu()
    assert torch.equal(logits, logits2)
    acts = dict(acts)
    for name, rows in (("l0b0", 512), ("l0b1", 512), ("stage0", 256), ("stage1", 256), ("stage2", 256), ("stage3", 256)):
        ref = g["act:" + name]
        got = acts[name][0, :rows].cpu().numpy()
        err = np.abs(got - ref).max() / max(1.0, np.abs(ref).max())
        assert err < 4e-2, (name, err)
    ref = g["logits"]
    rng = float(ref.max() - ref.min())
    assert float(np.abs(logits.numpy() - ref).max()) < 3e-2 * rng, float(np.abs(logits.numpy() - ref).max())
    assert (logits.numpy().argmax(1) == ref.argmax(1)).all()


def test_swin_end_to_end_from_jpeg_coefficients():
    """JPEG coefficients -> K0 (Swin layout, eval geometry) -> SwinV2-T logits on the GPU, against the oracle fed with
    the same quantised coefficients (dct_oracle data path + swin_oracle model, fp32 CPU)."""
    from rgb_no_more_b200 import dct_manip as dm
    B = 2
    y, c, q, flags = dm.decode_batch(synth.synth_jpeg_set(B), 64, 64, nthreads=2)
    tf = TF.get_transform("imagenet_dct_swin", "test", dtype=torch.bfloat16, device=DEV)
    m, sd = _model()
    m.to(DEV)
    with torch.no_grad():
        x = tf(y.to(DEV), c.to(DEV), q.to(DEV))
        assert x.shape == (B, 4096, 24) and x.dtype == torch.bfloat16
        logits = m(x).cpu()
    pl = P.eval_plan_swin(64, 64)
    ref = []
    for b in range(B):
        e = O.transform_embed_swin(y[b].reshape(1, 64, 64, 8, 8), c[b].reshape(2, 32, 32, 8, 8), q[b].reshape(3, 8, 8), pl, None)
        ref.append(SO.forward_from_embed(sd, e.reshape(1, 64, 64, 24)))
    ref = torch.cat(ref)
    rng = float(ref.max() - ref.min())
    assert float((logits - ref).abs().max()) < 3e-2 * rng
    assert torch.equal(logits.argmax(1), ref.argmax(1))


@pytest.mark.parametrize("M,N,K", [(4096, 384, 96), (2048, 1536, 384), (300, 768, 192)])
def test_gemm_gelu_act_epilogue(M, N, K):
    """RGBNM_EPI_GELU_ACT (inference: only gelu(acc + bias) is stored) against torch's exact-erf GELU."""
    from rgb_no_more_b200 import gemm as G
    g = torch.Generator().manual_seed(M + N)
    a = (torch.randn(M, K, generator=g) * 0.7).bfloat16().to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16().to(DEV)
    bias = (torch.randn(N, generator=g) * 0.2).to(DEV)
    got = G.gemm(a, w, G.EPI_GELU_ACT, bias=bias).float()
    ref = F.gelu(a.float() @ w.float().T + bias)
    assert float((got - ref).abs().max()) < 1e-2 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("M,N,K", [(4096, 96, 96), (2048, 192, 768), (1000, 384, 1536), (300, 96, 384), (8192, 384, 384)])
def test_gemm_lnres_epilogue(M, N, K):
    """RGBNM_EPI_LNRES: out = aux + LayerNorm(a w^T + bias) * gamma + beta inside the GEMM epilogue (swinv2.py:302-306),
    incl. N = 96 (half-filled 192-wide tile), ragged M and the 384-wide single-accumulator tile."""
    from rgb_no_more_b200 import gemm as G
    g = torch.Generator().manual_seed(M + N + K)
    a = (torch.randn(M, K, generator=g) * 0.7).bfloat16().to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16().to(DEV)
    bias = (torch.randn(N, generator=g) * 0.5 + 0.3).to(DEV)
    aux = torch.randn(M, N, generator=g).bfloat16().to(DEV)
    gamma = (1 + 0.2 * torch.randn(N, generator=g)).to(DEV)
    beta = (0.2 * torch.randn(N, generator=g)).to(DEV)
    got = G.gemm(a, w, G.EPI_LNRES, bias=bias, aux=aux, ln=(gamma, beta)).float()
    ref = aux.float() + F.layer_norm(a.float() @ w.float().T + bias, (N,), gamma, beta, 1e-5)
    assert float((got - ref).abs().max()) < 1e-2 * max(1.0, float(ref.abs().max()))
    with pytest.raises(Exception):
        G.gemm(a, w.repeat(8, 1)[: 768], G.EPI_LNRES, bias=None, aux=aux.repeat(1, 8)[:, :768].contiguous(),
               ln=(gamma.repeat(8)[:768].contiguous(), beta.repeat(8)[:768].contiguous()))      # N = 768 > one tile


@pytest.mark.parametrize("M,N,K", [(8192, 96, 24), (2048, 192, 384), (1024, 384, 768)])
def test_gemm_ln_epilogue(M, N, K):
    """RGBNM_EPI_LN: LayerNorm(a w^T (+ bias)) inside the GEMM epilogue (patch_embed.norm, PatchMerging.norm)."""
    from rgb_no_more_b200 import gemm as G
    g = torch.Generator().manual_seed(M + N + K)
    a = (torch.randn(M, K, generator=g) * 0.7).bfloat16().to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16().to(DEV)
    bias = (torch.randn(N, generator=g) * 0.5 + 0.3).to(DEV) if K == 24 else None
    gamma = (1 + 0.2 * torch.randn(N, generator=g)).to(DEV)
    beta = (0.2 * torch.randn(N, generator=g)).to(DEV)
    got = G.gemm(a, w, G.EPI_LN, bias=bias, ln=(gamma, beta)).float()
    pre = a.float() @ w.float().T
    ref = F.layer_norm(pre + bias if bias is not None else pre, (N,), gamma, beta, 1e-5)
    assert float((got - ref).abs().max()) < 1e-2 * max(1.0, float(ref.abs().max()))
