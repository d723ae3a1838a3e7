"""Per-role cycle counters of the GEMM main loop (CTA 0): needs a library built with ABX_GEMM_PROFILE=1 and ABX_GEMM_PROF=1 set."""
import ctypes, json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abx_b200 import lib, ops
L = lib.load()
names = {0: 'tma_wait_empty', 1: 'tma_total', 2: 'slabs', 4: 'mma_wait_accfree', 5: 'mma_wait_conv', 6: 'mma_total', 7: 'mma_issue',
         8: 'conv_wait_full', 9: 'conv_total', 12: 'acc0_wait_ready', 13: 'acc0_drain', 14: 'acc0_epilogue', 15: 'acc0_total',
         16: 'acc1_wait_ready', 17: 'acc1_drain', 18: 'acc1_epilogue', 19: 'acc1_total',
         20: 'mma1_wait_accfree', 21: 'mma1_wait_conv', 22: 'mma1_total', 23: 'mma1_issue'}
for m, n, k, tn, mode in ((490000, 768, 192, 128, 'plain'), (980000, 192, 128, 128, 'plain'), (980000, 192, 128, 128, 'res'),
                          (980000, 192, 128, 128, 'gate_res'), (980000, 192, 768, 128, 'res')):
    x = torch.randn(m, k, device='cuda'); w = torch.randn(n, k, device='cuda'); y = torch.empty(m, n, device='cuda')
    b = torch.randn(n, device='cuda')
    g = torch.randn(m, n, device='cuda') if 'gate' in mode else None
    r = torch.randn(m, n, device='cuda') if 'res' in mode else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(3):
        if it == 2: e0.record()
        ops.linear(x, w, b, act='gate' if g is not None else None, gate=g, residual=r, out=y, tile_n=tn)
    e1.record(); torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 32)()
    lib.check(L.abx_gemm_profile(buf))
    d = {v: int(buf[i]) for i, v in names.items()}
    per = {kk: round(vv / max(d['slabs'], 1), 1) for kk, vv in d.items() if kk != 'slabs'}
    print(json.dumps({'shape': [m, n, k], 'mode': mode, 'ms': round(e0.elapsed_time(e1), 4), 'slabs': d['slabs'], 'cycles_per_slab': per}))
