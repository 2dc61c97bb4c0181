"""Third part of the smoke run: JPEG coefficients -> K0 in the SwinV2 layout -> a two-stage SwinV2 DCT forward on cuda:0
(window attention with and without the cyclic shift, patch merging, post-norm LayerNorm), logits checked against the CPU
oracle (oracle/dct_oracle.py + oracle/swin_oracle.py: checkers only)."""
from __future__ import annotations

import torch


def run() -> None:
    from oracle import dct_oracle as O           # checker only
    from oracle import swin_oracle as SO         # checker only
    from rgb_no_more_b200 import dct_manip as dm
    from rgb_no_more_b200 import plan as P
    from rgb_no_more_b200 import swin as S
    from rgb_no_more_b200 import synth
    from rgb_no_more_b200 import transforms as TF

    dev = torch.device("cuda", 0)
    torch.manual_seed(11997733)
    B = 2
    y, c, q, flags = dm.decode_batch(synth.synth_jpeg_set(B), 64, 64, nthreads=2)
    tf = TF.get_transform("imagenet_dct_swin", "test", dtype=torch.float32, device=dev)
    emb = tf(y.to(dev), c.to(dev), q.to(dev))
    depths, heads = (2, 2), (3, 6)
    m = S.SwinTransformerV2(img_size=256, patch_size=4, embed_dim=96, depths=list(depths), num_heads=list(heads), window_size=8,
                            drop_path_rate=0.0, pretrained_window_sizes=[0, 0], device="cpu", pixel_space="dct")
    with torch.no_grad():
        for p in m.parameters():
            if p.ndim == 1:
                p.add_(0.2 * torch.randn_like(p))          # the reference zero-initialises the block post-norms
    sd = {k: (v.detach().float() if v.is_floating_point() else v.detach()).cpu().clone() for k, v in m.state_dict().items()}
    m.eval().to(dev)
    with torch.no_grad():
        logits = m(emb.to(torch.bfloat16)).float().cpu()
    torch.cuda.synchronize()
    pl = P.eval_plan_swin(64, 64)
    worst = 0.0
    for b in range(B):
        e = O.transform_embed_swin(y[b].reshape(1, 64, 64, 8, 8), c[b].reshape(2, 32, 32, 8, 8), q[b].reshape(3, 8, 8), pl, None)
        d = float((emb[b].cpu() - e).abs().max())
        # resize ties move single int16 LSBs (1/1020 in ToRange units, spread by the orthonormal decomposition)
        if d > 2e-3:
            raise AssertionError(f"rgbnm smoke: Swin embed input differs from the oracle by {d}")
        ref = SO.forward_from_embed(sd, e.reshape(1, 64, 64, 24), depths=depths, heads=heads)
        err = float((logits[b] - ref[0]).abs().max())
        rng = float(ref.max() - ref.min())
        worst = max(worst, err / max(rng, 1e-6))
        if not err <= 3e-2 * max(rng, 1.0):
            raise AssertionError(f"rgbnm smoke: SwinV2 logits differ from the oracle by {err} (range {rng})")
    print(f"rgbnm smoke: SwinV2 DCT forward ok (|dlogits| <= {worst:.3g} of the logit range)")
    # one training step of the same model (training engine: saved-activation forward, backward kernels): loss and two
    # gradients against autograd through the CPU oracle
    import torch.nn.functional as F
    m.train()
    x = emb.to(torch.bfloat16)
    labels = torch.tensor([3, 7], device=dev)
    loss = F.cross_entropy(m(x), labels)
    loss.backward()
    torch.cuda.synchronize()
    params = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "relative_coords_table" not in k and "attn_mask" not in k
                  else v.clone()) for k, v in sd.items()}
    ref_loss = F.cross_entropy(SO.forward_from_embed(params, x.float().cpu().reshape(B, 64, 64, 24), depths=depths, heads=heads), labels.cpu())
    ref_loss.backward()
    if not abs(float(loss) - float(ref_loss)) <= 2e-2 * max(float(ref_loss), 1.0):
        raise AssertionError(f"rgbnm smoke: SwinV2 training loss {float(loss)} vs oracle {float(ref_loss)}")
    named = dict(m.named_parameters())
    for k in ("layers.0.blocks.1.attn.qkv.weight", "layers.1.blocks.0.mlp.fc1.weight", "layers.0.blocks.1.attn.cpb_mlp.2.weight"):
        cos = float(F.cosine_similarity(named[k].grad.float().cpu().reshape(-1), params[k].grad.reshape(-1), dim=0))
        if not cos > 0.98:
            raise AssertionError(f"rgbnm smoke: SwinV2 gradient of {k} differs from the oracle (cos {cos})")
    print(f"rgbnm smoke: SwinV2 DCT training step ok (loss {float(loss):.4f} vs oracle {float(ref_loss):.4f})")
