"""Micro-benchmark of the trunk's streaming kernels at the benchmark's shape (B=8, N=350): achieved GB/s over their algorithmic bytes."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abx_b200 import lib, ops

B, N = 8, 350
g = torch.Generator(device='cuda').manual_seed(0)
L = lib.load()


def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / it


def row(name, ms, nbytes):
    print(json.dumps({'kernel': name, 'ms': round(ms, 4), 'algorithmic_MB': round(nbytes / 1e6, 1), 'GBps': round(nbytes / ms / 1e6, 1)}))


# pair input: concat(static, te, te) + LN(prev_pair) + emb[prev_pos]
C, Cs, Ct = 192, 128, 32
stat = torch.randn(1, N, N, Cs, device='cuda', generator=g); te = torch.randn(B, Ct, device='cuda', generator=g)
prev = torch.randn(B, N, N, C, device='cuda', generator=g); norm = torch.nn.LayerNorm(C).cuda()
pos = torch.randint(0, 15, (B, N, N), device='cuda', generator=g); emb = torch.randn(15, C, device='cuda', generator=g)
row('pair_input_kernel', t(lambda: ops.pair_input(stat, te, prev, norm, pos, emb)), B * N * N * (2 * C * 4 + 8) + N * N * Cs * 4)
# outer product mean features
l, r = torch.randn(B, N, 32, device='cuda', generator=g), torch.randn(B, N, 32, device='cuda', generator=g)
row('outer_product_kernel', t(lambda: ops.outer_product(l, r)), B * N * N * 64 * 4)
# channel-major LayerNorm after the triangle product
Cc, npad = 128, 352
x = torch.randn(B, Cc, N, npad, device='cuda', generator=g); y = torch.empty(B, N, N, Cc, device='cuda')
gam, bet = torch.randn(Cc, device='cuda', generator=g), torch.randn(Cc, device='cuda', generator=g)
row('layernorm_cm_kernel', t(lambda: lib.check(L.abx_layernorm_cm(lib.stream(), B, Cc, N, npad, lib.ptr(x), lib.ptr(gam), lib.ptr(bet), 1e-5, lib.ptr(y)))),
    B * N * N * Cc * 4 * 2)
# row-major LayerNorm of the pair tensor (the 6.3 TB/s yardstick)
p = torch.randn(B, N, N, C, device='cuda', generator=g)
row('layernorm_kernel<2>', t(lambda: ops.layer_norm(p, norm.weight, norm.bias)), B * N * N * C * 4 * 2)
# IPA pair bias
z = torch.randn(B, N, N, 128, device='cuda', generator=g); wp = torch.randn(12, 128, device='cuda', generator=g); bp = torch.randn(12, device='cuda', generator=g)
out = torch.empty(L.abx_ipa_pair_bias_floats(B, N), device='cuda')
row('ipa_pair_bias_chunked_kernel', t(lambda: lib.check(L.abx_ipa_pair_bias(lib.stream(), B, N, lib.ptr(z), lib.ptr(wp), lib.ptr(bp), lib.ptr(out)))),
    B * N * N * 128 * 4 + out.numel() * 4)
