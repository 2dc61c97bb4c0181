"""GPU probe: does rgbnm_gemm_bf16 (built and tuned on the ViT-S shapes) handle the SwinV2-T shapes -- N = 96 / 288 / 576,
K = 24 / 96, ragged tiles -- for the epilogues the Swin forward uses?  Prints max |err| vs torch (fp32 matmul of the bf16
operands) and CUDA-event timings per shape."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rgb_no_more_b200 import gemm as G

dev = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(0)
shapes = [("embed", B * 4096, 96, 24, G.EPI_STORE, True)]
for s in range(4):
    T, Cd = B * 4096 // 4 ** s, 96 * 2 ** s
    shapes += [(f"s{s}.qkv", T, 3 * Cd, Cd, G.EPI_STORE, True), (f"s{s}.proj", T, Cd, Cd, G.EPI_STORE, True),
               (f"s{s}.fc1", T, 4 * Cd, Cd, G.EPI_GELU, True), (f"s{s}.fc2", T, Cd, 4 * Cd, G.EPI_STORE, True)]
    if s < 3:
        shapes.append((f"s{s}.merge", T // 4, 2 * Cd, 4 * Cd, G.EPI_STORE, False))
shapes.append(("head", B, 1000, 768, G.EPI_F32, True))
bad = 0
for name, M, N, K, epi, has_bias in shapes:
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device=dev) * 0.1 if has_bias else None
    ref = a.float() @ w.float().T
    if bias is not None:
        ref = ref + bias
    try:
        if epi == G.EPI_GELU:
            pre, act = G.gemm(a, w, epi, bias=bias)
            got, ref2 = act.float(), torch.nn.functional.gelu(ref)
            err_pre = float((pre.float() - ref).abs().max())
        elif epi == G.EPI_F32:
            got, ref2, err_pre = G.gemm(a, w, epi, bias=bias), ref, 0.0
        else:
            got, ref2, err_pre = G.gemm(a, w, epi, bias=bias).float(), ref, 0.0
        torch.cuda.synchronize()
        err = float((got - ref2).abs().max())
        tol = 2e-2 * max(1.0, float(ref2.abs().max()))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            G.gemm(a, w, epi, bias=bias)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        ok = err < tol and err_pre < tol and bool(torch.isfinite(got).all())
        bad += 0 if ok else 1
        print(f"{name:10s} M={M:7d} N={N:5d} K={K:5d} epi={epi} err={err:.3e} pre={err_pre:.3e} {'OK ' if ok else 'BAD'} {us:8.1f} us "
              f"{2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s", flush=True)
    except Exception as ex:  # noqa: BLE001
        bad += 1
        print(f"{name:10s} M={M} N={N} K={K} epi={epi} EXC {ex}", flush=True)
print("bad:", bad)
