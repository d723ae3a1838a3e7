"""Run abx_gemm_tf32x3 once per shape (for ncu captures)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abx_b200 import ops
shapes = [tuple(int(v) for v in a.split('x')) for a in sys.argv[1:]] or [(490000, 768, 192, 128), (8192, 8192, 8192, 128)]
for m, n, k, tn in shapes:
    x = torch.randn(m, k, device='cuda'); w = torch.randn(n, k, device='cuda'); b = torch.randn(n, device='cuda')
    y = torch.empty(m, n, device='cuda')
    for _ in range(2):
        ops.linear(x, w, b, out=y, tile_n=tn)
    torch.cuda.synchronize()
