"""One tiny invocation of the hot path on cuda:0, checked against the CPU oracle
(called by __graft_entry__.smoke(); the oracle is the checker, never the product path)."""
from __future__ import annotations

import torch


def run() -> None:
    from oracle import dct_oracle as O          # checker only
    from rgb_no_more_b200 import dct_manip as dm
    from rgb_no_more_b200 import plan as P
    from rgb_no_more_b200 import synth
    from rgb_no_more_b200 import transforms as TF

    if not torch.cuda.is_available():
        raise RuntimeError("rgbnm smoke: no CUDA device; the B200 path has no CPU fallback")
    dev = "cuda:0"
    B = 4
    jpegs = synth.synth_jpeg_set(B)
    y, c, q, flags = dm.decode_batch(jpegs, 64, 64, nthreads=4)
    tf = TF.FusedDCT(dev, "train", P.AUGLIST_VITS, 2, 9)
    torch.manual_seed(11997733)
    plans = tf.sample_plans(B)
    yd, cd, qd = y.to(dev), c.to(dev), q.to(dev)
    cl = flags.tolist()
    planes = tf.run(yd, cd, qd, plans, clamp_in=cl, out_mode=TF.OUT_INT16_PLANES)
    resized = tf.run(yd, cd, qd, [P.Plan(p.crop_i, p.crop_j, p.crop_size) for p in plans], clamp_in=cl,
                     out_mode=TF.OUT_INT16_PLANES)
    emb = tf.run(yd, cd, qd, plans, clamp_in=cl, out_mode=TF.OUT_F32)
    torch.cuda.synchronize()
    gy, gc = TF.split_planes(planes.cpu())
    ry, rc = TF.split_planes(resized.cpu())
    ties = 0
    for b in range(B):
        v = (y[b].reshape(1, 64, 64, 8, 8), c[b].reshape(2, 32, 32, 8, 8), q[b].reshape(3, 8, 8))
        # (1) resize: identical to the oracle except one LSB on exact .5 ties of the float64 result
        ey, ec = O.resized_planes(*v, plans[b])
        xy, xc = O.resized_planes_exact(*v, plans[b])
        for got, ref, ex in ((ry[b], ey, xy), (rc[b], ec, xc)):
            d = (got.int() - ref.int()).abs()
            frac = (ex - ex.floor() - 0.5).abs()
            if int(d.max()) > 1 or bool(((d != 0) & (frac > 2e-3)).any()):
                raise AssertionError(f"rgbnm smoke: K0 resize differs from the oracle off a rounding tie (image {b})")
            ties += int((d != 0).sum())
        # (2) flip + RandAugment ops: bit-exact given the resized planes
        fy, fc = O.transform_from_resized(ry[b].clone(), rc[b].clone(), plans[b], tf.bank.table)
        if not (torch.equal(fy, gy[b]) and torch.equal(fc, gc[b])):
            raise AssertionError(f"rgbnm smoke: K0 augmentation stage is not bit-exact (image {b})")
        # (3) ToRange + sub-block conversion (fp32)
        ref = O.embed_input(O.to_range(gy[b]).unsqueeze(0), O.to_range(gc[b]).unsqueeze(0)).reshape(196, 384)
        err = float((emb[b].cpu() - ref).abs().max())
        if err > 2e-5:
            raise AssertionError(f"rgbnm smoke: K0 embed input differs from the oracle by {err}")
    print(f"rgbnm smoke: K0 ok on {torch.cuda.get_device_name(0)} (ops bit-exact; {ties} tie-rounding LSB flips in resize)")
    from . import swin as swin_smoke
    from . import vit as vit_smoke
    vit_smoke.run()
    swin_smoke.run()
