#!/usr/bin/env python
"""GPU diagnostic for the tcgen05 attention kernels against torch (prints error statistics and timings)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rgb_no_more_b200 import attention as A

dev = "cuda:0"
torch.manual_seed(0)
fails = 0


def check(B, H, E, do_time=False):
    global fails
    D, N = 64, 196
    scale = 1.0 / math.sqrt(E)
    qkv = (torch.randn(B * N, 3 * H * D, device=dev) * 2.0).to(torch.bfloat16)
    o = torch.zeros(B * N, H * D, dtype=torch.bfloat16, device=dev)
    lse = torch.zeros(B, H, N, device=dev)
    A.forward(qkv, o, lse, B, H, D, scale, backend="b200")
    torch.cuda.synchronize()
    v = qkv.float().view(B, N, 3, H, D).permute(2, 0, 3, 1, 4)
    s = torch.einsum("bhqd,bhkd->bhqk", v[0], v[1]) * scale
    ref = torch.einsum("bhqk,bhkd->bhqd", torch.softmax(s, -1), v[2]).transpose(1, 2).reshape(B * N, H * D)
    lref = torch.logsumexp(s, -1)
    d = (o.float() - ref).abs()
    dl = (lse - lref).abs()
    bad = float(d.max()) > 2e-2 * float(ref.abs().max()) or float(dl.max()) > 1e-2 or not torch.isfinite(o.float()).all()
    print(f"{'FAIL' if bad else 'ok  '} fwd B{B} H{H}: max|do|={float(d.max()):.4g} (ref max {float(ref.abs().max()):.3g}) "
          f"mean|do|={float(d.mean()):.3g} max|dlse|={float(dl.max()):.3g}", flush=True)
    if bad:
        fails += 1
        idx = torch.nonzero(d > 2e-2 * ref.abs().max())
        print("   n_bad", idx.shape[0], "of", d.numel(), "first", idx[:8].tolist())
        rows = torch.unique(idx[:, 0] % N)[:20].tolist(); cols = torch.unique(idx[:, 1])[:20].tolist()
        print("   bad token rows", rows, "bad cols", cols)
        print("   sample got", o[0, :8].float().tolist(), "\n   sample ref", ref[0, :8].tolist())
    # ---- backward ----
    do = torch.randn(B * N, H * D, device=dev).to(torch.bfloat16)
    dqkv = torch.zeros_like(qkv)
    A.backward(do, qkv, o, lse, dqkv, B, H, D, scale, backend="b200")
    torch.cuda.synchronize()
    leaf = qkv.float().requires_grad_(True)
    vv = leaf.view(B, N, 3, H, D).permute(2, 0, 3, 1, 4)
    ss = torch.einsum("bhqd,bhkd->bhqk", vv[0], vv[1]) * scale
    oo = torch.einsum("bhqk,bhkd->bhqd", torch.softmax(ss, -1), vv[2]).transpose(1, 2).reshape(B * N, H * D)
    (gref,) = torch.autograd.grad(oo, leaf, do.float())
    HDs = H * D
    for nm, sl in (("dq", slice(0, HDs)), ("dk", slice(HDs, 2 * HDs)), ("dv", slice(2 * HDs, 3 * HDs))):
        dd = (dqkv[:, sl].float() - gref[:, sl]).abs()
        rmax = float(gref[:, sl].abs().max())
        bad = float(dd.max()) > 3e-2 * rmax or not torch.isfinite(dqkv.float()).all()
        print(f"{'FAIL' if bad else 'ok  '} bwd {nm} B{B} H{H}: max|d|={float(dd.max()):.4g} (ref max {rmax:.3g}) mean|d|={float(dd.mean()):.3g}", flush=True)
        if bad:
            fails += 1
            idx = torch.nonzero(dd > 3e-2 * rmax)
            print("   n_bad", idx.shape[0], "of", dd.numel(), "first", idx[:8].tolist())
            print("   bad token rows", torch.unique(idx[:, 0] % N)[:24].tolist(), "bad cols", torch.unique(idx[:, 1])[:24].tolist())
            print("   got", dqkv[:, sl][0, :6].float().tolist(), "\n   ref", gref[:, sl][0, :6].tolist())
    if do_time:
        for name, be in (("b200", "b200"),):
            for _ in range(3):
                A.backward(do, qkv, o, lse, dqkv, B, H, D, scale, backend=be)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                A.backward(do, qkv, o, lse, dqkv, B, H, D, scale, backend=be)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"time bwd {name} B{B} H{H}: {ms*1e3:.1f} us", flush=True)
        for name, be in (("b200", "b200"),):
            for _ in range(3):
                A.forward(qkv, o, lse, B, H, D, scale, backend=be)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                A.forward(qkv, o, lse, B, H, D, scale, backend=be)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"time fwd {name} B{B} H{H}: {ms*1e3:.1f} us  {4*B*H*N*N*D/ms/1e9:.1f} TFLOP/s (algorithmic)", flush=True)


check(1, 1, 64)
check(2, 3, 192)
check(5, 6, 384)
check(256, 6, 384, do_time=True)
print("FAILS", fails)
sys.exit(1 if fails else 0)
