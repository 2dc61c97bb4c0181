#!/usr/bin/env python
"""One forward + one backward of the attention kernels at the ViT-S bench shape (for ncu captures)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rgb_no_more_b200 import attention as A

B, H, D, N = 256, 6, 64, 196
dev = "cuda:0"
torch.manual_seed(0)
scale = 1.0 / math.sqrt(H * D)
qkv = (torch.randn(B * N, 3 * H * D, device=dev) * 2.0).to(torch.bfloat16)
o = torch.zeros(B * N, H * D, dtype=torch.bfloat16, device=dev)
lse = torch.zeros(B, H, N, device=dev)
do = torch.randn(B * N, H * D, device=dev).to(torch.bfloat16)
dqkv = torch.zeros_like(qkv)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    A.forward(qkv, o, lse, B, H, D, scale, backend="b200")
    A.backward(do, qkv, o, lse, dqkv, B, H, D, scale, backend="b200")
torch.cuda.synchronize()
print("done")
