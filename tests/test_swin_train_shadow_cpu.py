"""CPU: the host-side orchestration of the SwinV2 training engine (rgb_no_more_b200/swin_train.py: what is saved, which
gradient flows where, residual / post-norm / stochastic-depth bookkeeping, the parameter-only graph behind the attention
tables) checked against the reference's own loss and gradients (tests/golden/swin_train.npz).  The CUDA kernels cannot run
here, so every kernel entry point of the engine is replaced by a few lines of torch with the same contract; the kernels
themselves are tested one by one on the GPU (tests/test_swin_bwd_gpu.py, tests/test_swin_gpu.py)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import swin_oracle as SO
from rgb_no_more_b200 import gemm as G
from rgb_no_more_b200 import swin as S
from rgb_no_more_b200 import swin_train as ST
from tests.helpers import load, seeded_swin_state_dict, golden_swin_inputs

BF = torch.bfloat16


def _gelu_grad(u):
    u = u.float()
    return 0.5 * (1 + torch.erf(u / 2 ** 0.5)) + u * torch.exp(-0.5 * u * u) / (2 * torch.pi) ** 0.5


def fake_gemm(a, b, epilogue=G.EPI_STORE, bias=None, aux=None, out=None, out2=None, posemb=None, out_f32=None, splits=0, alpha=1.0,
              trans_out=False, perm_heads=0, perm_head_dim=0, ln=None, ln_eps=1e-5):
    if epilogue == G.EPI_WGRAD_ATOMIC:
        r = a.float().t() @ b.float()                       # [M, N]
        out_f32 += r.t() if trans_out else r
        return out_f32
    acc = a.float() @ b.float().t()
    if bias is not None:
        acc = acc + bias
    if epilogue == G.EPI_GELU:
        return acc.to(BF), F.gelu(acc).to(BF)
    if epilogue == G.EPI_DGELU:
        acc = acc * _gelu_grad(aux)
    elif epilogue == G.EPI_RESIDUAL:
        acc = acc + aux.float()
    return acc.to(BF)


class ShadowEngine(ST.SwinTrainEngine):
    def __init__(self, model):
        self.model, self.device, self.L = model, torch.device("cpu"), None
        depths = [len(layer.blocks) for layer in model.layers]
        self.dpr = [float(v) for v in torch.linspace(0, model.drop_path_rate, sum(depths))]

    def _ln_fwd(self, x, norm, res, scale, rows_per_scale, y):
        o = F.layer_norm(x.float(), (x.shape[1],), norm[0], norm[1], 1e-5)
        if scale is not None:
            o = o * scale.repeat_interleave(rows_per_scale).unsqueeze(1)
        if res is not None:
            o = o + res.float()
        y.copy_(o.to(BF))
        return y

    def _ln_bwd(self, dy, x, gamma, scale, rows_per_scale, gname, bname, dxsum_name=None):
        with torch.enable_grad():               # autograd runs Function.backward with grad mode off
            xr = x.float().requires_grad_(True)
            g = gamma.clone().requires_grad_(True)
            b = torch.zeros_like(gamma).requires_grad_(True)
            o = F.layer_norm(xr, (x.shape[1],), g, b, 1e-5)
            if scale is not None:
                o = o * scale.repeat_interleave(rows_per_scale).unsqueeze(1)
            o.backward(dy.float())
        self.grads[gname] += g.grad
        self.grads[bname] += b.grad
        dx = xr.grad.to(BF)
        if dxsum_name is not None:                  # the kernel also accumulates the column sums of the dx it writes
            self.grads[dxsum_name] += dx.float().sum(0)
        return dx

    @staticmethod
    def _attn(qkv, bias, scale, B, H, Cd, heads, shift):
        x = qkv.view(B, H, H, 3 * Cd)
        if shift:
            x = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2))
        xw = SO.window_partition(x, 8).view(-1, 64, 3 * Cd)
        q, k, v = xw.reshape(-1, 64, 3, heads, 32).permute(2, 0, 3, 1, 4)
        attn = F.normalize(q, dim=-1) @ F.normalize(k, dim=-1).transpose(-2, -1)
        attn = attn * scale.view(1, heads, 1, 1) + bias.unsqueeze(0)
        if shift:
            mask = SO.shift_mask(H, H, 8, shift)
            attn = (attn.view(-1, mask.shape[0], heads, 64, 64) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, 64, 64)
        o = (attn.softmax(-1) @ v).transpose(1, 2).reshape(-1, 8, 8, Cd)
        o = SO.window_reverse(o, 8, H, H)
        if shift:
            o = torch.roll(o, shifts=(shift, shift), dims=(1, 2))
        return o.reshape(B * H * H, Cd)

    def _attn_fwd(self, qkv, bias, scale, B, H, Cd, heads, window, shift):
        return self._attn(qkv.float(), bias, scale, B, H, Cd, heads, shift).to(BF)

    def _attn_bwd(self, qkv, datt, bias, scale, dbias, dscale, B, H, Cd, heads, window, shift):
        with torch.enable_grad():
            q = qkv.float().requires_grad_(True)
            b, s = bias.clone().requires_grad_(True), scale.clone().requires_grad_(True)
            self._attn(q, b, s, B, H, Cd, heads, shift).backward(datt.float())
        dbias.add_(b.grad)                       # the kernel accumulates into the group's preallocated slices
        dscale.add_(s.grad)
        return q.grad.to(BF), dbias, dscale

    def _gather(self, x, B, H, Cd):
        x = x.view(B, H, H, Cd)
        return torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1).reshape(-1, 4 * Cd).contiguous()

    def _scatter(self, dy, B, H, Cd):
        dy = dy.view(B, H // 2, H // 2, 4, Cd)
        out = torch.empty((B, H, H, Cd), dtype=dy.dtype)
        out[:, 0::2, 0::2], out[:, 1::2, 0::2], out[:, 0::2, 1::2], out[:, 1::2, 1::2] = dy[..., 0, :], dy[..., 1, :], dy[..., 2, :], dy[..., 3, :]
        return out.reshape(-1, Cd)


@pytest.fixture
def patched(monkeypatch):
    monkeypatch.setattr(ST.G, "gemm", fake_gemm)
    monkeypatch.setattr(ST.K, "colsum", lambda a, out, *r: out.add_(a.float().sum(0)))
    monkeypatch.setattr(ST.K, "weight_prep", lambda w, wb, wt, *r: (wb.copy_(w.to(BF)), wt.copy_(w.t().to(BF))))


def test_training_orchestration_matches_reference_gradients(patched):
    g = load("swin_train.npz")
    m = S.SwinTransformerV2(img_size=256, patch_size=4, embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=8,
                            mlp_ratio=4, drop_path_rate=0.0, pretrained_window_sizes=[0, 0, 0, 0], device="cpu", pixel_space="dct")
    m.load_state_dict(seeded_swin_state_dict(m))
    m.train()
    eng = ShadowEngine(m)
    yf, cf = golden_swin_inputs(g["input_seed"])
    x = S.swin_embed_input_from_planes(yf, cf)
    names = [n for n, _ in m.named_parameters()]
    logits = ST.SwinFunction.apply(x, eng, names, *[p for _, p in m.named_parameters()])
    labels = torch.zeros((2, 1000))
    labels[0, 3], labels[0, 7], labels[1, 999] = 0.7, 0.3, 1.0
    loss = torch.nn.CrossEntropyLoss()(logits, labels)
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) < 2e-2 * float(g["loss"])
    named = dict(m.named_parameters())
    for k in [k[5:] for k in g.files if k.startswith("grad:")]:
        got = named[k].grad.float()
        ref = torch.from_numpy(g["grad:" + k])
        cos = float(F.cosine_similarity(got.reshape(-1)[:4096], ref, dim=0))
        assert cos > 0.99, (k, cos)
        assert abs(float(got.norm()) - float(g["gradnorm:" + k])) < 5e-2 * float(g["gradnorm:" + k]), (k, float(got.norm()), float(g["gradnorm:" + k]))


def test_stochastic_depth_bookkeeping_matches_oracle(patched):
    """Per-image, per-branch DropPath scales: the engine's forward / backward with a fixed set of masks against autograd
    through the oracle with the same masks (two-stage model, drop_path 0.5 so that many branches are dropped)."""
    torch.manual_seed(5)
    depths, heads = (2, 2), (3, 6)
    m = S.SwinTransformerV2(img_size=256, patch_size=4, embed_dim=96, depths=list(depths), num_heads=list(heads), window_size=8,
                            drop_path_rate=0.5, pretrained_window_sizes=[0, 0], device="cpu", pixel_space="dct")
    with torch.no_grad():
        for p in m.parameters():
            if p.ndim == 1:
                p.add_(0.2 * torch.randn_like(p))
    m.train()
    B = 3
    eng = ShadowEngine(m)
    drawn = []

    def draw(Bn, keep):
        s = torch.empty(Bn).bernoulli_(keep) / keep
        drawn.append(s)
        return s
    eng._draw = draw
    x = torch.randn(B, 4096, 24) * 0.5
    names = [n for n, _ in m.named_parameters()]
    logits = ST.SwinFunction.apply(x, eng, names, *[p for _, p in m.named_parameters()])
    y = torch.tensor([1, 2, 3])
    F.cross_entropy(logits, y).backward()
    # block 0 has drop probability 0 (linspace(0, rate, n)[0]): no draw; the other three blocks draw two masks each
    assert len(drawn) == 6 and any(float(s.min()) == 0.0 for s in drawn)
    scales = [(None, None)] + [(drawn[2 * i], drawn[2 * i + 1]) for i in range(3)]
    sd = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() and "relative_coords_table" not in k and "attn_mask" not in k
              else v.detach().clone()) for k, v in m.state_dict().items()}
    ref = SO.forward_from_embed(sd, x.to(BF).float().reshape(B, 64, 64, 24), depths=depths, heads=heads, drop_scales=scales)
    F.cross_entropy(ref, y).backward()
    assert float((logits.detach() - ref.detach()).abs().max()) < 3e-2 * float(ref.detach().abs().max())
    named = dict(m.named_parameters())
    for k in ("layers.0.blocks.1.attn.qkv.weight", "layers.1.blocks.1.mlp.fc2.weight", "layers.1.blocks.0.norm1.weight",
              "patch_embed.projection.0.weight", "layers.0.downsample.reduction.weight"):
        cos = float(F.cosine_similarity(named[k].grad.reshape(-1), sd[k].grad.reshape(-1), dim=0))
        assert cos > 0.99, (k, cos)


def test_weight_copies_are_cached_until_the_parameter_changes(patched):
    m = S.SwinTransformerV2(img_size=256, patch_size=4, embed_dim=96, depths=[2], num_heads=[3], window_size=8,
                            drop_path_rate=0.0, pretrained_window_sizes=[0], device="cpu", pixel_space="dct")
    eng = ShadowEngine(m)
    w = m.layers[0].blocks[0].attn.qkv.weight
    a = eng._lin(w)
    assert eng._lin(w) is a
    with torch.no_grad():
        w.add_(1.0)                                   # what an optimiser step does: bumps the parameter's version
    b = eng._lin(w)
    assert b is not a and torch.equal(b.w.float(), w.detach().to(BF).float())
