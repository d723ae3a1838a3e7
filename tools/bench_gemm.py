"""Micro-benchmark of abx_gemm_tf32x3 against torch's fp32 matmul (cuBLAS SIMT sgemm) on cuda:0."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def time_ms(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    from abx_b200 import ops
    torch.backends.cuda.matmul.allow_tf32 = False
    shapes = [(1400, 1152, 256), (1400, 256, 2112), (1400, 256, 256), (490000, 768, 192), (490000, 192, 768), (490000, 192, 192),
              (490000, 128, 192), (8192, 8192, 8192)]
    for m, n, k in shapes:
        x = torch.randn(m, k, device='cuda')
        w = torch.randn(n, k, device='cuda')
        b = torch.randn(n, device='cuda')
        y = torch.empty(m, n, device='cuda')
        row = dict(M=m, N=n, K=k)
        row['torch_ms'] = time_ms(lambda: torch.nn.functional.linear(x, w, b))
        for tn in (32, 64, 128):
            row[f'abx_bn{tn}_ms'] = time_ms(lambda: ops.linear(x, w, b, out=y, tile_n=tn))
        best = min(row[f'abx_bn{tn}_ms'] for tn in (32, 64, 128))
        row['abx_tflops_fp32_equiv'] = 2 * m * n * k / best / 1e9
        row['torch_tflops'] = 2 * m * n * k / row['torch_ms'] / 1e9
        if m > 100000:
            g = torch.randn(m, n, device='cuda')
            r = torch.randn(m, n, device='cuda')
            row['abx_gate_res_ms'] = time_ms(lambda: ops.linear(x, w, b, act='gate', gate=g, residual=r, out=y))
            row['abx_res_ms'] = time_ms(lambda: ops.linear(x, w, b, residual=r, out=y))
            del g, r
        ref = torch.nn.functional.linear(x[:2048].double(), w.double(), b.double())
        row['abx_maxerr'] = float((ops.linear(x[:2048], w, b).double() - ref).abs().max())
        row['torch_maxerr'] = float((torch.nn.functional.linear(x[:2048], w, b).double() - ref).abs().max())
        print(json.dumps(row), flush=True)


if __name__ == '__main__':
    main()
