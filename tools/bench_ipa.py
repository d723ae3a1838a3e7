"""Micro-benchmark of one IPA layer-call (abx_ipa_forward) on cuda:0: CUDA-event time per call and the
achieved fraction of the measured HBM peak for SURVEY §8d's algorithmic bytes.

    python tools/bench_ipa.py [--B 8] [--N 350] [--iters 20]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--B', type=int, default=8)
    ap.add_argument('--N', type=int, default=350)
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--precomputed-bias', type=int, default=1)
    ap.add_argument('--profile', type=int, default=0)
    ap.add_argument('--prof', type=int, default=0, help='per-role stall profile of the fused kernel (abx_ipa_profile) over 5 calls')
    ap.add_argument('--graph', type=int, default=0, help='time G back-to-back layer-calls (same z, as IpaScore issues them) replayed from a CUDA graph')
    a = ap.parse_args()
    from abx_b200.model.folding import InvariantPointAttention
    from abx_b200.utils.weights import load_seeded_
    conf = dict(num_head=12, num_channel=256, num_scalar_qk=16, num_scalar_v=16, num_point_qk=4, num_point_v=8)
    ipa = load_seeded_(InvariantPointAttention(conf, 128), 0).cuda().eval()
    B, N = a.B, a.N
    g = torch.Generator(device='cuda').manual_seed(0)
    x = torch.randn(B, N, 256, device='cuda', generator=g)
    z = torch.randn(B, N, N, 128, device='cuda', generator=g)
    q = torch.randn(B, N, 4, device='cuda', generator=g); q = q / q.norm(dim=-1, keepdim=True)
    from abx_b200.model.quat_affine import quat_to_rot
    rots, trans = quat_to_rot(q), torch.randn(B, N, 3, device='cuda', generator=g) * 2
    mask = torch.ones(B, N, device='cuda')
    bias = ipa.pair_bias(z) if a.precomputed_bias else None
    bias_ms = None
    if a.precomputed_bias:                                    # the once-per-IpaScore pair-bias pass, timed alone (cold L2)
        bt = []
        flush0 = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
        for _ in range(a.iters):
            flush0.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ipa.pair_bias(z); e1.record()
            torch.cuda.synchronize()
            bt.append(e0.elapsed_time(e1))
        bt.sort(); bias_ms = bt[len(bt) // 2]
        del flush0
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
    with torch.no_grad():
        for _ in range(3):
            ipa(x, z, mask, (rots, trans), pair_bias=bias)
        times = []
        for _ in range(a.iters):
            flush.zero_()                                     # evict z from the 126 MB L2
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ipa(x, z, mask, (rots, trans), pair_bias=bias)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
    graph_ms = None
    if a.graph:
        side = torch.cuda.Stream()
        with torch.no_grad(), torch.cuda.stream(side):
            cg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cg, stream=side):
                keep = [ipa(x, z, mask, (rots, trans), pair_bias=bias) for _ in range(a.graph)]
        torch.cuda.synchronize()
        gt = []
        for _ in range(a.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            cg.replay()
            e1.record()
            torch.cuda.synchronize()
            gt.append(e0.elapsed_time(e1) / a.graph)
        gt.sort()
        graph_ms = gt[len(gt) // 2]
    if a.profile:
        from torch.profiler import ProfilerActivity, profile
        with torch.no_grad(), profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(5):
                flush.zero_()
                ipa(x, z, mask, (rots, trans), pair_bias=bias)
            torch.cuda.synchronize()
        for e in sorted(prof.key_averages(), key=lambda e: -e.self_device_time_total)[:10]:
            print(f'{e.self_device_time_total / e.count:9.1f} us x{e.count:3d}  {e.key[:90]}')
    if a.prof:
        import ctypes
        from abx_b200 import lib
        L = lib.load()
        lib.check(L.abx_ipa_profile(1, None))
        with torch.no_grad():
            for _ in range(5):
                flush.zero_()
                ipa(x, z, mask, (rots, trans), pair_bias=bias)
        buf = (ctypes.c_ulonglong * 64)()
        lib.check(L.abx_ipa_profile(0, buf))
        names = ['z producer (wait: slot free)', 'kv producer (wait: buffer free)', 'mma issuers (waits: p_full, a_full)',
                 'converters (waits: z_full, a_empty)', 'logits (waits: kv_full, p_empty)', 'values (waits: kv_full, p_full)']
        for r, nm in enumerate(names):
            loop, w0, w1, n = (buf[8 * r + k] for k in range(4))
            if n and (buf[8 * r + 4] or buf[8 * r + 5]):
                print(f'      extra: t2 {100 * buf[8 * r + 4] / max(loop, 1):5.1f}%  t3 {100 * buf[8 * r + 5] / max(loop, 1):5.1f}% of the loop')
            if n:
                print(f'  {nm:40s} loop {loop / n / 1.965e3:8.1f} us/call  wait0 {100 * w0 / max(loop, 1):5.1f}%  wait1 {100 * w1 / max(loop, 1):5.1f}%  (n={n})')
    times.sort()
    ms = times[len(times) // 2]
    alg = B * 4 * (128 * N * N + 2 * 256 * N + 12 * N + N) + 4 * 838552
    peak = 6550.7
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs']
    except Exception:
        pass
    gbs = alg / (ms * 1e-3) / 1e9
    print(json.dumps(dict(B=B, N=N, ms_per_layer_call=ms, us_per_element=ms * 1e3 / B, algorithmic_bytes=alg,
                          achieved_gbs=gbs, peak_gbs=peak, frac=gbs / peak, precomputed_bias=bool(a.precomputed_bias),
                          graph_calls=a.graph, graph_ms_per_layer_call=graph_ms,
                          graph_frac=(alg / (graph_ms * 1e-3) / 1e9 / peak) if graph_ms else None,
                          pair_bias_ms=bias_ms,
                          frac_with_bias_amortised=(alg / ((ms + bias_ms / 8) * 1e-3) / 1e9 / peak) if bias_ms else None,
                          pdl=os.environ.get('ABX_IPA_PDL', '1'), fused=os.environ.get('ABX_IPA_FUSED', '1'))))


if __name__ == '__main__':
    main()
