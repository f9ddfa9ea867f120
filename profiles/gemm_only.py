"""Runs the three fp32 GEMM kernels of cfg 2 a few times (for ncu captures)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pinthememory_b200 import capi

dev = "cuda"
dtype = torch.bfloat16 if "--bf16" in sys.argv else torch.float32
B, h, w = 8, 96, 96
x = torch.randn(B, 288, h, w, device=dev).to(dtype)
g = torch.randn(B, 256, h, w, device=dev).to(dtype)
W = torch.randn(256, 288, device=dev) / 17
hi, lo = capi.conv1x1_prep(W, False, dtype)
hit, lot = capi.conv1x1_prep(W, True, dtype)
y = torch.empty(B, 256, h, w, device=dev, dtype=dtype)
dx = torch.empty(B, 288, h, w, device=dev, dtype=dtype)
st = torch.zeros(512, dtype=torch.float64, device=dev)
dW = torch.empty(256, 288, device=dev)
for _ in range(4):
    capi.conv1x1_fwd(x, hi, lo, 256, y=y, stats=st)
    capi.conv1x1_fwd(g, hit, lot, 288, y=dx)
    capi.conv1x1_wgrad(g, x, dW=dW)
torch.cuda.synchronize()
