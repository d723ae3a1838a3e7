"""Host vs device time of one ScoreNetwork.forward + reverse step (torch.profiler), to spot launch-bound
stretches and host synchronisations.   python tools/profile_step.py [--B 4] [--n-antigen 120]"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--B', type=int, default=4)
    ap.add_argument('--n-antigen', type=int, default=120)
    ap.add_argument('--num-t', type=int, default=3)
    ap.add_argument('--linear-breakdown', type=int, default=1)
    ap.add_argument('--glue', type=int, default=0, help='attribute the torch (non-abx) device time to python source lines')
    a = ap.parse_args()
    import __graft_entry__
    __graft_entry__.build()
    from abx_b200 import sampler
    from abx_b200.data.synthetic import synthetic_complex
    from abx_b200.diffuser.full_diffuser import FullDiffuser
    from abx_b200.model import features as F_
    from abx_b200.model.abx import ScoreNetwork
    from abx_b200.utils.weights import load_seeded_
    dev = torch.device('cuda:0')
    cfg = bench.model_config()
    fd = FullDiffuser(cfg['diffuser'])
    model = load_seeded_(ScoreNetwork(cfg['model'], fd), 0).to(dev).eval()
    raw = synthetic_complex(n_antigen=a.n_antigen, seed=0, batch_size=a.B)
    dev_fields = ['seq', 'mask', 'chain_id', 'atom14_gt_positions', 'atom14_gt_exists', 'cdr_def', 'residx', 'anchor_flag']
    b = {k: (v.to(dev) if k in dev_fields else v) for k, v in raw.items()}
    feat_cfg = bench.feature_config(cfg, dev, fd)
    batch = F_.FeatureBuilder(feat_cfg).build(b)

    def run():
        torch.manual_seed(0)
        return sampler.sample_loop(dict(batch), cfg, fd, model, mode='design', num_t=a.num_t)

    run()
    torch.cuda.synchronize()
    if a.linear_breakdown:
        from abx_b200 import ops
        orig = ops.linear
        recs = []

        def timed_linear(x, weight, bias=None, act=None, residual=None, gate=None, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            y = orig(x, weight, bias, act, residual, gate, **kw)
            e1.record()
            M = x.numel() // x.shape[-1]
            recs.append(((M, weight.shape[0], x.shape[-1], act, residual is not None, gate is not None, x.is_contiguous()), e0, e1))
            return y

        ops.linear = timed_linear
        run()
        torch.cuda.synchronize()
        ops.linear = orig
        agg = {}
        for key, e0, e1 in recs:
            c = agg.setdefault(key, [0, 0.0])
            c[0] += 1
            c[1] += e0.elapsed_time(e1)
        tot = sum(v[1] for v in agg.values())
        print(f'linear calls: {len(recs)}, total {tot:.1f} ms (includes input .contiguous() copies)')
        for key, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
            print(f'  {ms:8.2f} ms  {n:4d} x {ms / n:7.3f}  (M,N,K,act,res,gate,contig)={key}')
    if a.glue:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
            run()
            torch.cuda.synchronize()
        rows = {}
        for e in prof.key_averages(group_by_stack_n=12):
            if e.self_device_time_total <= 0 or 'abx::' in e.key or 'Memcpy' in e.key:
                continue
            frames = [f for f in e.stack if '/abx_b200/' in f or '/bench.py' in f]
            where = frames[0].split('/abx_b200/')[-1] if frames else (e.stack[0] if e.stack else '?')
            r = rows.setdefault((e.key, where), [0, 0.0])
            r[0] += e.count
            r[1] += e.self_device_time_total
        tot = sum(v[1] for v in rows.values())
        print(f'torch-op device time: {tot / 1e3:.2f} ms over {a.num_t + 1} forwards (B={a.B})')
        for (op, where), (n, us) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:40]:
            print(f'  {us / 1e3:8.3f} ms {n:5d} x {us / n:8.1f} us  {op[:36]:36s} {where[:110]}')
        return
    t0 = time.perf_counter()
    run()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        run()
        torch.cuda.synchronize()
    ev = prof.key_averages()
    dev_us = sum(e.self_device_time_total for e in ev)
    n_fw = a.num_t + 1
    print(json.dumps(dict(B=a.B, wall_ms=wall * 1e3, forwards=n_fw, wall_ms_per_forward=wall * 1e3 / n_fw,
                          device_ms_total=dev_us / 1e3, device_ms_per_forward=dev_us / 1e3 / n_fw)))
    print(ev.table(sort_by='self_cpu_time_total', row_limit=25, max_name_column_width=60))
    print(ev.table(sort_by='self_device_time_total', row_limit=30, max_name_column_width=90))


if __name__ == '__main__':
    main()
