"""GPU: SwinV2-T DCT eval forward (BASELINE config 5 shape on one GPU): K0 in the Swin layout (Resize_DCT(32) geometry,
bf16 out) -> SwinTransformerV2 forward, batch 256, coefficients resident in HBM, 2 distinct input batches cycled
(2 x 201 MB > 126 MB L2), CUDA-graph replay, CUDA events.  Prints one JSON line.
    python tools/swin_bench.py [--steps N] [--batch B] [--no-graph]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rgb_no_more_b200 import plan as P, synth, swin as S, transforms as TF

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--no-graph", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda:0")
B, NB = args.batch, 2
torch.manual_seed(0)
model = S.SwinTransformerV2(img_size=256, patch_size=4, embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=8,
                            mlp_ratio=4, drop_path_rate=0.2, pretrained_window_sizes=[0, 0, 0, 0], device="cpu", pixel_space="dct")
with torch.no_grad():          # the reference zero-initialises the post-norms (blocks = identity at init): randomise them
    for p in model.parameters():
        if p.ndim == 1:
            p.add_(0.1 * torch.randn_like(p))
model.eval().to(dev)
eng = model.prepare(dev)
tf = TF.FusedDCT(dev, "test", None, 0, 0, torch.bfloat16, out_size=32)
plans = tf.sample_plans(B)
pdev = torch.from_numpy(P.pack_plans(plans, [False] * B, out_size=32).view(np.uint8).reshape(B, -1)).to(dev)
pool = []
for i in range(NB):
    y, c, q = synth.synth_coefficients(B, 64, 64, seed=70 + i, dense=False)
    pool.append((torch.from_numpy(y).to(dev), torch.from_numpy(c).to(dev), torch.from_numpy(q).to(dev)))
x = torch.empty((B, 4096, 24), dtype=torch.bfloat16, device=dev)


def step(k):
    tf.run(*pool[k % NB], None, plans_dev=pdev, out=x)
    return eng.forward(x)


with torch.no_grad():
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for k in range(3):
            logits = step(k)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    eng.launches = 0
    step(0)
    launches = eng.launches + 2
    if not args.no_graph:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for k in range(NB):
                step(k)
        run = lambda: g.replay()
    else:
        run = lambda: [step(k) for k in range(NB)]
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        run()
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / (args.steps * NB)
# algorithmic forward FLOPs per image (dense contractions + attention MACs x 2)
fl = 2 * 4096 * 24 * 96 + 2 * 768 * 1000
for s, depth in enumerate((2, 2, 6, 2)):
    T, Cd = 4096 // 4 ** s, 96 * 2 ** s
    fl += depth * (24 * T * Cd * Cd + 256 * T * Cd)
    if s < 3:
        fl += 4 * T * Cd * Cd
print(json.dumps({"metric": "images/sec (SwinV2-T DCT window 8, eval forward)", "value": B / ms * 1e3, "unit": "images/s", "n_gpus": 1,
                  "batch": B, "ms_per_batch": ms, "gflop_per_image_fwd": fl / 1e9, "achieved_tflops": fl * B / ms / 1e9,
                  "gpu_launches_per_batch": launches, "dtype": "bf16", "data": "synthetic",
                  "what": "K0 (Swin layout, Resize_DCT(32)) + SwinTransformerV2 forward, no gradients, CUDA-graph replay" if not args.no_graph
                  else "eager", "logits_finite": bool(torch.isfinite(logits).all())}))
