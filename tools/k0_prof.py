"""GPU: K0 (ViT layout) alone, batch 256, 4 distinct resident input batches cycled (> L2): eval geometry (crop 56 -> 28) and the
training mix.  `python tools/k0_prof.py [reps]` prints CUDA-graph-replay timings (one JSON line); under ncu it is the short
command the `--set full` capture wraps (profiles/README.md).  RGBNM_K0_V1=1 selects the first-generation kernel for A/B."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rgb_no_more_b200 import plan as P, synth, transforms as TF

dev = "cuda:0"
B, NB = 256, 4
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6548.8
try:
    peak = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    pass
batches = []
for i in range(NB):
    y, c, q = synth.synth_coefficients(B, 64, 64, seed=50 + i, dense=False)
    batches.append((torch.from_numpy(y).to(dev), torch.from_numpy(c).to(dev), torch.from_numpy(q).to(dev)))
res = {}
for kind in ("eval", "train", "train_fused_only"):       # the last: the same plans without the statistics pre-pass launch (timing only)
    tf = TF.FusedDCT(dev, "train" if kind != "eval" else "test", P.AUGLIST_VITS, 2, 9, torch.bfloat16)
    ns = False if kind != "train" else None
    torch.manual_seed(5)
    plans = tf.sample_plans(B)
    pdev = torch.from_numpy(P.pack_plans(plans, [False] * B).view(np.uint8).reshape(B, -1)).to(dev)
    out = torch.empty((B, 196, 384), dtype=torch.bfloat16, device=dev)
    byt = sum(p.crop_size ** 2 * 128 + 2 * (p.crop_size // 2) ** 2 * 128 + 496 + 196 * 384 * 2 for p in plans)
    for i in range(3):
        tf.run(*batches[i % NB], None, plans_dev=pdev, out=out, needs_stats=ns)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(NB):
            tf.run(*batches[i], None, plans_dev=pdev, out=out, needs_stats=ns)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (reps * NB)
    res[kind] = {"ms_per_batch_incl_dcstats": round(ms, 4), "alg_bytes_per_image": byt // B, "achieved_gbps": round(byt / ms / 1e6, 1),
                 "peak_gbps": peak, "frac": round(byt / ms / 1e6 / peak, 3),
                 "crop_mix": {str(s): sum(1 for p in plans if p.crop_size == s) for s in (14, 28, 56)}}
print(json.dumps({"kernel": "k0 ViT layout, bf16 out" + (" (v1)" if os.environ.get("RGBNM_K0_V1") else " (v2)"), "batch": B, **res}))
