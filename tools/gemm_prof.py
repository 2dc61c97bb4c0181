#!/usr/bin/env python
"""One launch of each GEMM flavour at the ViT-S bench shapes (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rgb_no_more_b200 import gemm as G

dev = "cuda:0"
torch.manual_seed(0)
M = 50176
def mk(m, k, s=1.0): return (torch.randn(m, k, device=dev) * s).to(torch.bfloat16)
x384, x1536 = mk(M, 384), mk(M, 1536)
w_qkv, w_fc1, w_fc2 = mk(1152, 384, 0.05), mk(1536, 384, 0.05), mk(384, 1536, 0.03)
b1152, b1536, b384 = torch.randn(1152, device=dev), torch.randn(1536, device=dev), torch.randn(384, device=dev)
o1152, o1536a, o1536b, o384 = (torch.empty(M, n, dtype=torch.bfloat16, device=dev) for n in (1152, 1536, 1536, 384))
gw = torch.zeros(1536, 384, device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    G.gemm(x384, w_qkv, G.EPI_STORE, bias=b1152, out=o1152)                       # qkv forward
    G.gemm(x384, w_fc1, G.EPI_GELU, bias=b1536, out=o1536a, out2=o1536b)          # fc1 + GELU
    G.gemm(x1536, w_fc2, G.EPI_RESIDUAL, bias=b384, aux=x384, out=o384)           # fc2 + residual
    G.gemm(x384, w_fc1, G.EPI_DGELU, aux=x1536, out=o1536a)                       # dgrad fc2 x GELU'
    G.gemm(x1536, x384, G.EPI_WGRAD_ATOMIC, out_f32=gw)                           # wgrad fc1
torch.cuda.synchronize()
print("done")
