#!/usr/bin/env python
"""GPU diagnostic for the tcgen05 GEMM: every epilogue against torch, with error statistics printed
(not only asserted) so that one gpurun call tells as much as possible.  Exits non-zero on failure."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rgb_no_more_b200 import gemm as G

dev = "cuda:0"
torch.manual_seed(0)
fails = 0


def report(name, got, ref, tol):
    global fails
    d = (got.float() - ref.float()).abs()
    rel = float(d.max()) / (float(ref.float().abs().max()) + 1e-9)
    bad = rel > tol or not torch.isfinite(got.float()).all()
    print(f"{'FAIL' if bad else 'ok  '} {name}: max|d|={float(d.max()):.4g} rel={rel:.3g} mean|d|={float(d.mean()):.3g} "
          f"ref_max={float(ref.float().abs().max()):.3g}", flush=True)
    if bad:
        fails += 1
        idx = torch.nonzero(d > tol * ref.float().abs().max())
        print("   first bad idx:", idx[:6].tolist(), " n_bad:", idx.shape[0], "of", d.numel())
        rows = torch.unique(idx[:, 0])[:12].tolist(); cols = torch.unique(idx[:, 1])[:12].tolist()
        print("   bad rows:", rows, " bad cols:", cols)


def mk(m, k, scale=1.0):
    return (torch.randn(m, k, device=dev) * scale).to(torch.bfloat16)


def run_case(M, N, K):
    a, w = mk(M, K), mk(N, K, K ** -0.5)
    bias = torch.randn(N, device=dev)
    ref = a.float() @ w.float().t()
    report(f"STORE nobias M{M} N{N} K{K}", G.gemm(a, w, G.EPI_STORE), ref, 1e-2)
    report(f"STORE bias   M{M} N{N} K{K}", G.gemm(a, w, G.EPI_STORE, bias=bias), ref + bias, 1e-2)
    res = mk(M, N)
    report(f"RESIDUAL     M{M} N{N} K{K}", G.gemm(a, w, G.EPI_RESIDUAL, bias=bias, aux=res), ref + bias + res.float(), 1e-2)
    u, f = G.gemm(a, w, G.EPI_GELU, bias=bias)
    report(f"GELU pre     M{M} N{N} K{K}", u, ref + bias, 1e-2)
    report(f"GELU act     M{M} N{N} K{K}", f, torch.nn.functional.gelu(ref + bias), 1e-2)
    pre = mk(M, N)
    x = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(x).sum().backward()
    report(f"DGELU        M{M} N{N} K{K}", G.gemm(a, w, G.EPI_DGELU, aux=pre), ref * x.grad, 1e-2)
    if N % 32 == 0:
        pos = torch.randn(196, N, device=dev)
        rows = torch.arange(M, device=dev) % 196
        report(f"POSEMB       M{M} N{N} K{K}", G.gemm(a, w, G.EPI_POSEMB, bias=bias, posemb=pos), ref + bias + pos[rows], 1e-2)
    report(f"F32          M{M} N{N} K{K}", G.gemm(a, w, G.EPI_F32, bias=bias), ref + bias, 5e-3)


def run_wgrad(T, M, N, splits):
    dy, x = mk(T, M), mk(T, N)
    ref = dy.float().t() @ x.float()
    out = torch.zeros(M, N, device=dev)
    G.gemm(dy, x, G.EPI_WGRAD_ATOMIC, out_f32=out, splits=splits, alpha=1.0)
    report(f"WGRAD T{T} M{M} N{N} splits{splits}", out, ref, 5e-3)
    G.gemm(dy, x, G.EPI_WGRAD_ATOMIC, out_f32=out, splits=splits, alpha=0.5)
    report(f"WGRAD accumulate alpha", out, 1.5 * ref, 5e-3)
    if not os.environ.get("RGBNM_GEMM_V1"):
        outT = torch.zeros(N, M, device=dev)
        G.gemm(dy, x, G.EPI_WGRAD_ATOMIC, out_f32=outT, splits=0, trans_out=True)
        report(f"WGRAD trans_out auto-splits T{T} M{M} N{N}", outT, ref.t(), 5e-3)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "small"):
    run_case(128, 192, 64)
    run_case(256, 384, 128)
    run_case(1000, 1000, 384)      # ragged M and N
    run_case(392, 384, 1536)       # N = 384, long K: the 256 x 384 tile
    run_case(1000, 768, 1152)
if which in ("all", "wgrad"):
    run_wgrad(64, 128, 192, 1)
    run_wgrad(256, 384, 384, 2)
    run_wgrad(1024, 200, 104, 4)   # ragged outputs (leading dimensions must stay multiples of 8)
if which in ("all", "big"):
    run_case(50176, 1152, 384)
    run_wgrad(50176, 1536, 384, 16)
    run_case(50176, 1536, 384)
    # timing at the ViT-S shapes (B=256)
    shapes = [("qkv", 50176, 1152, 384, G.EPI_STORE), ("proj", 50176, 384, 384, G.EPI_RESIDUAL),
              ("fc1", 50176, 1536, 384, G.EPI_GELU), ("fc2", 50176, 384, 1536, G.EPI_RESIDUAL),
              ("dfc1", 50176, 1536, 384, G.EPI_DGELU), ("dfc2", 50176, 384, 1536, G.EPI_STORE),
              ("dqkv", 50176, 384, 1152, G.EPI_STORE), ("dproj", 50176, 384, 384, G.EPI_STORE)]
    for name, M, N, K, epi in shapes:
        a, w = mk(M, K), mk(N, K, K ** -0.5)
        bias = torch.randn(N, device=dev)
        aux = mk(M, N) if epi in (G.EPI_RESIDUAL, G.EPI_DGELU) else None
        out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        out2 = torch.empty(M, N, dtype=torch.bfloat16, device=dev) if epi == G.EPI_GELU else None
        for _ in range(3):
            G.gemm(a, w, epi, bias=bias, aux=aux, out=out, out2=out2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            G.gemm(a, w, epi, bias=bias, aux=aux, out=out, out2=out2)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"time {name}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
        e0.record()
        for _ in range(10):
            torch.nn.functional.linear(a, w)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"     cublas plain {name}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
    for name, Mo, No in (("fc1", 1536, 384), ("qkv", 1152, 384), ("proj", 384, 384)):
        dy, x = mk(50176, Mo), mk(50176, No)
        out = torch.zeros(Mo, No, device=dev)
        for sp in ((0, 2, 4, 8) if not os.environ.get("RGBNM_GEMM_V1") else (8, 16)):
            G.gemm(dy, x, G.EPI_WGRAD_ATOMIC, out_f32=out, splits=sp)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                G.gemm(dy, x, G.EPI_WGRAD_ATOMIC, out_f32=out, splits=sp)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"time wgrad {name} splits {sp}: {ms*1e3:.1f} us  {2*50176*Mo*No/ms/1e9:.1f} TFLOP/s", flush=True)
    if not os.environ.get("RGBNM_GEMM_V1"):
        x, dy = mk(50176, 1536), mk(50176, 384)          # fc2: the longer side is the input -> transposed form
        out = torch.zeros(384, 1536, device=dev)
        G.gemm(x, dy, G.EPI_WGRAD_ATOMIC, out_f32=out, splits=0, trans_out=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            G.gemm(x, dy, G.EPI_WGRAD_ATOMIC, out_f32=out, splits=0, trans_out=True)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"time wgrad fc2 (trans_out) auto: {ms*1e3:.1f} us  {2*50176*1536*384/ms/1e9:.1f} TFLOP/s", flush=True)
torch.cuda.synchronize()
print("FAILS", fails)
sys.exit(1 if fails else 0)
