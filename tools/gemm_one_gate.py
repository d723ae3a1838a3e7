"""Run the gated + residual GEMM of TriangleMultiplication's output projection once (for ncu captures)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abx_b200 import ops
m, n, k = 980000, 192, 128
x = torch.randn(m, k, device='cuda'); w = torch.randn(n, k, device='cuda'); b = torch.randn(n, device='cuda')
g = torch.randn(m, n, device='cuda'); r = torch.randn(m, n, device='cuda'); y = torch.empty(m, n, device='cuda')
for _ in range(2):
    ops.linear(x, w, b, act='gate', gate=g, residual=r, out=y)
torch.cuda.synchronize()
