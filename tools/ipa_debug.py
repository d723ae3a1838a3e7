"""Debug driver of the fused IPA kernel: one attention_features call per size against the CPU oracle, error per feature
segment, and the kernel's watchdog record (which wait timed out, if any).   python tools/ipa_debug.py 2,37 8,350"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abx_b200 import lib  # noqa: E402
from abx_b200.utils.weights import np_randn  # noqa: E402
from oracle import model as M  # noqa: E402
from oracle import quat as Q  # noqa: E402
from tests.test_gpu_ipa import make_ipa  # noqa: E402


def watchdog():
    buf = (ctypes.c_ulonglong * 8)()
    lib.check(lib.load().abx_ipa_watchdog_read(buf))
    return list(buf)


def main():
    sizes = [tuple(int(v) for v in a.split(',')) for a in sys.argv[1:]] or [(2, 37)]
    ipa, P = make_ipa()
    failed = []
    for B, N in sizes:
        gen = torch.Generator().manual_seed(100 + N)
        x, z = np_randn(300 + N, B, N, 256), np_randn(400 + N, B, N, N, 128)
        q = torch.randn(B, N, 4, generator=gen); q = q / q.norm(dim=-1, keepdim=True)
        rots, trans = Q.quat_to_rot(q), torch.randn(B, N, 3, generator=gen) * 2.0
        mask = torch.ones(B, N)
        if N > 8:
            mask[-1, -(N // 5):] = 0
            mask[0, 3] = 0
        ref, parts = M.ipa_forward(P, x, z, mask, rots, trans, return_parts=True)
        with torch.no_grad():
            feats = ipa.attention_features(x.cuda(), z.cuda(), mask.cuda(), (rots.cuda(), trans.cuda()))
        wd = watchdog()
        try:
            with torch.no_grad():
                bias = ipa.pair_bias(z.cuda())
                for _ in range(3):
                    out = ipa(x.cuda(), z.cuda(), mask.cuda(), (rots.cuda(), trans.cuda()), pair_bias=bias)
                torch.cuda.synchronize()
            print(f'  forward x3: rel.err {float((out.cpu() - ref).abs().max()) / float(ref.abs().max()):.2e} watchdog={watchdog()[:5]}', flush=True)
        except Exception as e:
            print('  forward FAILED:', str(e)[-300:], flush=True)
            raise
        f, r = feats.cpu(), parts['feats']
        seg = {'o_scalar': (0, 192), 'o_point': (192, 480), 'o_norm': (480, 576), 'o_pair': (576, 2112)}
        line = {k: float((f[..., a:b] - r[..., a:b]).abs().max()) / float(r[..., a:b].abs().max()) for k, (a, b) in seg.items()}
        print(f'B={B} N={N} watchdog={wd[:5]} rel.err={line} finite={bool(torch.isfinite(f).all())}', flush=True)
        if wd[0] != 0 or not max(line.values()) < 3e-5:
            failed.append((B, N))
        if wd[0] == 0 and max(line.values()) > 1e-3:
            d = (f - r).abs()
            bad = torch.nonzero(d > 1e-3 * r.abs().max())
            print('  first bad entries (b, i, col):', bad[:8].tolist(), ' of', len(bad))
            b0, i0, c0 = bad[0].tolist()
            print('  got', f[b0, i0, c0:c0 + 4].tolist(), 'want', r[b0, i0, c0:c0 + 4].tolist())
    if failed:
        print('FAILED:', failed)
        sys.exit(1)


if __name__ == '__main__':
    main()
