"""GPU: CUDA-event timing of K0 in the SwinV2 layout (eval geometry = Resize_DCT(32) of the whole 64 x 64-block image,
and the RandomResizedCrop_DCT(32) training mix), batch 256, 4 distinct input batches cycled (> L2).  Prints a JSON line
with the algorithmic-bytes roofline (SURVEY.md 8d convention: int16 coefficients inside the crop window + 384 B tables
+ 112 B plan + output elements)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rgb_no_more_b200 import plan as P, synth, transforms as TF

dev = "cuda:0"
B, NB = 256, 4
peak = 6548.8
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    pass
batches = []
for i in range(NB):
    y, c, q = synth.synth_coefficients(B, 64, 64, seed=50 + i, dense=False)
    batches.append((torch.from_numpy(y).to(dev), torch.from_numpy(c).to(dev), torch.from_numpy(q).to(dev)))
res = {}
for kind in ("eval", "train"):
    tf = TF.FusedDCT(dev, "train" if kind == "train" else "test", P.AUGLIST_VITS, 2, 9, torch.bfloat16, out_size=32)
    torch.manual_seed(5)
    plans = tf.sample_plans(B)
    packed = P.pack_plans(plans, [False] * B, out_size=32)
    pdev = torch.from_numpy(packed.view(np.uint8).reshape(B, -1)).to(dev)
    out = torch.empty((B, 4096, 24), dtype=torch.bfloat16, device=dev)
    byt = sum(p.crop_size ** 2 * 128 + 2 * (p.crop_size // 2) ** 2 * 128 + 496 + 4096 * 24 * 2 for p in plans)
    for i in range(5):
        tf.run(*batches[i % NB], None, plans_dev=pdev, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 40
    e0.record()
    for i in range(K):
        tf.run(*batches[i % NB], None, plans_dev=pdev, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    res[kind] = {"ms_per_batch_incl_dcstats": round(ms, 4), "images_per_s": round(B / ms * 1e3), "alg_bytes_per_image": byt // B,
                 "achieved_gbps": round(byt / ms / 1e6, 1), "peak_gbps": peak, "frac": round(byt / ms / 1e6 / peak, 3),
                 "crop_mix": {str(s): sum(1 for p in plans if p.crop_size == s) for s in (16, 32, 64)}}
print(json.dumps({"kernel": "k0_fused<bf16, SWIN4>", "batch": B, **res}))
