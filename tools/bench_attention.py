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
if os.environ.get('ABX_ATTN_PROF') == '1':
    import ctypes
    from abx_b200 import lib
    out = ops.pair_attention(qkv, bias, mask, H, impl='tc5')
    buf = (ctypes.c_ulonglong * 32)()
    lib.check(lib.load().abx_attention_profile(buf))
    names = {0: 'soft0_wait_o', 1: 'soft0_wait_s', 2: 'soft0_total', 4: 'load_wait_empty', 5: 'load_total', 8: 'mma_wait_kv', 9: 'mma_wait_p',
             10: 'mma_issue', 11: 'mma_total', 12: 'key_tiles', 13: 'prologue', 14: 'store', 15: 'cta_total',
             16: 'ph_drain_o', 17: 'ph_load_s', 18: 'ph_logits_max', 19: 'ph_exp_stP', 20: 'ph_stwait_arrive', 21: 'ph_bias_issue'}
    nk = max(int(buf[12]), 1)
    print(json.dumps({'cycles_per_key_tile': {v: round(int(buf[i]) / nk, 1) for i, v in names.items() if i != 12}, 'key_tiles': nk}))
