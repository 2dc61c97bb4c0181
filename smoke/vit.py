"""Second half of the smoke run: the embed input produced by K0 goes through a small DCT ViT (forward, backward and
one optimiser step) on cuda:0 -- every tcgen05 / memory-bound kernel of the path is launched once -- and the logits and
a gradient are checked against the CPU oracle (oracle/vit_oracle.py: checker only)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def run() -> None:
    from oracle import vit_oracle as VO          # checker only
    from rgb_no_more_b200 import train_step as TS
    from rgb_no_more_b200 import vit as V

    dev = torch.device("cuda", 0)
    torch.manual_seed(11997733)
    B, depth = 4, 2
    m = V.ViT(patch_size=16, emb_size=192, depth=depth, n_classes=1000, drop_p=0.0, pixel_space="DCT", ver=1, use_subblock=True,
              num_heads=3, head_size=64).to(dev)
    sd = {k: v.detach().float().cpu().clone() for k, v in m.state_dict().items()}
    x = (torch.randn(B, 196, 384) * 0.5).to(torch.bfloat16)
    labels = torch.zeros(B, 1000)
    labels[torch.arange(B), torch.arange(B) * 7] = 1.0
    logits = m(x.to(dev))
    loss = torch.nn.CrossEntropyLoss()(logits, labels.to(dev))
    loss.backward()
    torch.cuda.synchronize()
    # oracle: same bf16-rounded input, fp32 arithmetic on the CPU
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = VO.forward_embedded(params, x.float().reshape(B, 14, 14, 384), depth=depth)
    ref_loss = torch.nn.CrossEntropyLoss()(ref, labels)
    (gref,) = torch.autograd.grad(ref_loss, params["encoder.0.1.fn.eb_ffb.0.weight"])
    err = float((logits.float().cpu() - ref.detach()).abs().max())
    rng = float(ref.detach().abs().max())
    if not err <= 2e-2 * max(rng, 1.0):          # bf16 operands, fp32 accumulation
        raise AssertionError(f"rgbnm smoke: ViT logits differ from the oracle by {err} (range {rng})")
    got = dict(m.named_parameters())["encoder.0.1.fn.eb_ffb.0.weight"].grad.float().cpu()
    cos = float(F.cosine_similarity(got.reshape(-1), gref.reshape(-1), dim=0))
    if not cos > 0.99:
        raise AssertionError(f"rgbnm smoke: ViT weight gradient direction differs from the oracle (cos {cos})")
    # the CUDA-graph train stage (mixup -> fwd -> CE -> bwd -> clip + AdamW + decay -> bf16 refresh)
    st = TS.TrainStage(dev, arch="vitti", batch=B, warmup_steps=1, total_steps=100, use_graph=True)
    y = torch.arange(B, device=dev) % 1000
    l0 = float(st.step(x.to(dev), y))
    for _ in range(5):
        l1 = float(st.step(x.to(dev), y))
    if not (l1 == l1 and l1 < l0):
        raise AssertionError(f"rgbnm smoke: train stage loss did not decrease ({l0} -> {l1})")
    print(f"rgbnm smoke: ViT ok (|dlogits| {err:.3g} of range {rng:.3g}, grad cos {cos:.4f}, train stage loss {l0:.3f} -> {l1:.3f})")
