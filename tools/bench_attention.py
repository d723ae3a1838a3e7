"""Micro-benchmark of abx_pair_attention (SIMT vs tensor-core implementation) at the trunk's shape."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abx_b200 import ops
B, S, L, H, D = 8, 350, 350, 4, 48
g = torch.Generator(device='cuda').manual_seed(0)
qkv = torch.randn(B, S, L, 3 * H * D, device='cuda', generator=g)
bias = torch.randn(B, H, L, L, device='cuda', generator=g)
mask = torch.ones(B, L, device='cuda', dtype=torch.bool)
ref = None
for impl in ('tc5', 'mma'):
    for _ in range(2):
        out = ops.pair_attention(qkv, bias, mask, H, impl=impl)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = ops.pair_attention(qkv, bias, mask, H, impl=impl)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    flops = 4.0 * B * S * H * L * L * D
    if ref is None:
        q, k, v = (t.reshape(B, S, L, H, D).permute(0, 1, 3, 2, 4)[:1, :8].double() for t in qkv.chunk(3, dim=-1))
        w = torch.softmax(torch.einsum('bshqd,bshkd->bshqk', q * D ** -0.5, k) + bias.double()[:1, None], dim=-1)
        ref = torch.einsum('bshqk,bshkd->bsqhd', w, v).reshape(1, 8, L, H * D)
    err = float((out[:1, :8].double() - ref).abs().max())
    print(json.dumps(dict(impl=impl, ms=ms, tflops=flops / ms / 1e9, maxerr=err)))
