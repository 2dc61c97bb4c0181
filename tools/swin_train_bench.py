"""GPU: SwinV2-T DCT training step (first version of the training engine: K0 in the Swin layout with the RandAugment mix ->
forward with saved activations -> CE -> backward -> torch AdamW), batch 64 per GPU (reference: 512 over 8 GPUs,
utils/configs.py:137), drop_path 0.2, CUDA events.  Prints one JSON line."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from rgb_no_more_b200 import plan as P, synth, swin as S, transforms as TF

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--stage", action="store_true", help="train_step.TrainStage(arch='swinv2t'): flat buffers, fused optimiser, CUDA graphs")
ap.add_argument("--no-graph", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda:0")
B = args.batch
if args.stage:
    from rgb_no_more_b200 import train_step as TS
    st = TS.TrainStage(dev, arch="swinv2t", batch=B, use_graph=not args.no_graph)
    with torch.no_grad():
        st.eng.flat.add_(0.1 * torch.randn_like(st.eng.flat) * (st.eng.flat == 0))
    st.eng.refresh_weights()
    tf = TF.FusedDCT(dev, "train", P.AUGLIST_VITS, 2, 9, torch.bfloat16, out_size=32)
    y, c, q = synth.synth_coefficients(B, 64, 64, seed=90, dense=False)
    y, c, q = torch.from_numpy(y).to(dev), torch.from_numpy(c).to(dev), torch.from_numpy(q).to(dev)
    labels = torch.randint(0, 1000, (B,), device=dev)
    gen = torch.Generator().manual_seed(7)

    def sstep():
        plans = tf.sample_plans_packed(B, 64, 64, clamp_in=[False] * B, generator=gen)
        x = tf.run(y, c, q, plans, needs_stats=bool(plans["needs_stats"].any()), out=st.x_static)
        return st.step(x, labels).clone()
    for _ in range(4):
        l0 = sstep()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        l1 = sstep()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({"metric": "images/sec (SwinV2-T DCT window 8, train step, TrainStage)", "value": B / ms * 1e3, "unit": "images/s", "n_gpus": 1,
                      "batch": B, "ms_per_step": ms, "loss_first": float(l0), "loss_last": float(l1), "graph": not args.no_graph,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))
    sys.exit(0)
torch.manual_seed(0)
model = S.SwinTransformerV2(img_size=256, patch_size=4, embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=8,
                            mlp_ratio=4, drop_path_rate=0.2, pretrained_window_sizes=[0, 0, 0, 0], device="cpu", pixel_space="dct")
with torch.no_grad():
    for p in model.parameters():
        if p.ndim == 1:
            p.add_(0.1 * torch.randn_like(p))
model.train().to(dev)
opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
tf = TF.FusedDCT(dev, "train", P.AUGLIST_VITS, 2, 9, torch.bfloat16, out_size=32)
y, c, q = synth.synth_coefficients(B, 64, 64, seed=90, dense=False)
y, c, q = torch.from_numpy(y).to(dev), torch.from_numpy(c).to(dev), torch.from_numpy(q).to(dev)
labels = torch.randint(0, 1000, (B,), device=dev)


def step():
    x = tf(y, c, q)
    opt.zero_grad(set_to_none=True)
    loss = F.cross_entropy(model(x), labels)
    loss.backward()
    opt.step()
    return loss


for _ in range(2):
    l0 = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    l1 = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
print(json.dumps({"metric": "images/sec (SwinV2-T DCT window 8, train step, first version)", "value": B / ms * 1e3, "unit": "images/s",
                  "n_gpus": 1, "batch": B, "ms_per_step": ms, "loss_first": float(l0), "loss_last": float(l1),
                  "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30, "dtype": "bf16", "data": "synthetic",
                  "what": "K0 (Swin layout, RandAugment mix) + forward + CE + backward + torch AdamW, eager launches, drop_path 0.2"}))
