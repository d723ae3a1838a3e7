"""Does the tensor core ignore the 13 low mantissa bits of a tf32 operand?  Run with ABX_GEMM_TRUST_TRUNC=0/1
and compare the saved outputs bit for bit."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abx_b200 import ops
g = torch.Generator(device='cuda').manual_seed(0)
x = torch.randn(777, 544, device='cuda', generator=g); w = torch.randn(300, 544, device='cuda', generator=g)
y = ops.linear(x, w)
torch.save(y.cpu(), sys.argv[1])
