"""Stress driver of the fused IPA kernel: many calls on fresh random inputs (wide coordinate spreads so that softmax reference
points move), L2 flushed in between, checking for faults / watchdog records / non-finite outputs and fused-vs-itself
determinism.   python tools/ipa_stress.py B N iters [spread]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abx_b200 import lib  # noqa: E402
from abx_b200.model.quat_affine import quat_to_rot  # noqa: E402
from tests.test_gpu_ipa import make_ipa  # noqa: E402


def main():
    B, N, iters = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    spread = float(sys.argv[4]) if len(sys.argv) > 4 else 2.0
    pmask = float(sys.argv[5]) if len(sys.argv) > 5 else 0.1
    ipa, _ = make_ipa()
    L = lib.load()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
    g = torch.Generator(device='cuda').manual_seed(1)
    for it in range(iters):
        x = torch.randn(B, N, 256, device='cuda', generator=g)
        z = torch.randn(B, N, N, 128, device='cuda', generator=g)
        q = torch.randn(B, N, 4, device='cuda', generator=g); q = q / q.norm(dim=-1, keepdim=True)
        rots, trans = quat_to_rot(q), torch.randn(B, N, 3, device='cuda', generator=g) * spread
        mask = (torch.rand(B, N, device='cuda', generator=g) >= pmask).float()
        with torch.no_grad():
            bias = ipa.pair_bias(z)
            flush.zero_()
            a = ipa(x, z, mask, (rots, trans), pair_bias=bias)
            flush.zero_()
            b = ipa(x, z, mask, (rots, trans), pair_bias=bias)
        torch.cuda.synchronize()
        buf = (ctypes.c_ulonglong * 8)()
        lib.check(L.abx_ipa_watchdog_read(buf))
        ok = bool(torch.isfinite(a).all()) and torch.equal(a, b) and buf[0] == 0
        print(f'iter {it}: finite={bool(torch.isfinite(a).all())} deterministic={torch.equal(a, b)} watchdog={list(buf)[:5]} rescales={buf[6]}', flush=True)
        if not ok:
            sys.exit(1)
    print('stress ok')


if __name__ == '__main__':
    main()
