"""One tiny invocation of the hot path on cuda:0, checked against the CPU oracle
(called by __graft_entry__.smoke(); the oracle is the checker, never the product path)."""
from __future__ import annotations

import torch


def run() -> None:
    from oracle import dct_oracle as O          # checker only
    from . import dct_manip as dm
    from . import plan as P
    from . import synth
    from . import transforms as TF

    if not torch.cuda.is_available():
        raise RuntimeError("rgbnm smoke: no CUDA device; the B200 path has no CPU fallback")
    dev = "cuda:0"
    B = 4
    jpegs = synth.synth_jpeg_set(B)
    y, c, q, flags = dm.decode_batch(jpegs, 64, 64, nthreads=4)
    tf = TF.FusedDCT(dev, "train", P.AUGLIST_VITS, 2, 9)
    torch.manual_seed(11997733)
    plans = tf.sample_plans(B)
    planes = tf.run(y.to(dev), c.to(dev), q.to(dev), plans, clamp_in=flags.tolist(), out_mode=TF.OUT_INT16_PLANES)
    emb = tf.run(y.to(dev), c.to(dev), q.to(dev), plans, clamp_in=flags.tolist(), out_mode=TF.OUT_F32)
    torch.cuda.synchronize()
    gy, gc = TF.split_planes(planes.cpu())
    worst = 0
    for b in range(B):
        ry, rc = O.transform_int16(y[b].reshape(1, 64, 64, 8, 8), c[b].reshape(2, 32, 32, 8, 8), q[b].reshape(3, 8, 8),
                                   plans[b], tf.bank.table)
        worst = max(worst, int((gy[b].int() - ry.int()).abs().max()), int((gc[b].int() - rc.int()).abs().max()))
        ref = O.transform_embed(y[b].reshape(1, 64, 64, 8, 8), c[b].reshape(2, 32, 32, 8, 8), q[b].reshape(3, 8, 8),
                                plans[b], tf.bank.table)
        err = float((emb[b].cpu() - ref).abs().max())
        if err > 1.5 * 2.0 / 2040:
            raise AssertionError(f"rgbnm smoke: K0 embed input differs from the oracle by {err}")
    if worst > 1:
        raise AssertionError(f"rgbnm smoke: K0 int16 planes differ from the oracle by {worst} LSB")
    print(f"rgbnm smoke: K0 ok on {torch.cuda.get_device_name(0)} (max int16 diff {worst} LSB)")
    try:
        from . import vit_smoke
    except ImportError:
        return
    vit_smoke.run()
