"""Run abx_layernorm on the pair tensor shape of the benchmark (for ncu captures and a CUDA-event time)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abx_b200 import ops
B, N, C = 8, 350, 192
x = torch.randn(B, N, N, C, device='cuda'); w = torch.randn(C, device='cuda'); b = torch.randn(C, device='cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for tn in (0, N):
    for _ in range(3):
        y = ops.layer_norm(x, w, b, transpose_n=tn)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); y = ops.layer_norm(x, w, b, transpose_n=tn); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(json.dumps(dict(shape=[B, N, N, C], transpose_n=tn, ms=ms, gbs=2 * x.numel() * 4 / ms / 1e6)))
