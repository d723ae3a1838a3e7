"""SURVEY section 7 hard part (b) / north_star "match on identical RNG seeds": the product's reverse-diffusion loop against
the REFERENCE's own module stack (oracle/_ref, materialised by oracle/build_ref.py) ON THE SAME GPU with the same
torch.manual_seed — both arms start from the same t=1 prior draw and consume the default CUDA generator in the same order
(randn, randn, poisson per step: full_diffuser.py:174-227), so they follow the same noise path until float32 differences
in the scores flip a Poisson count.  Writes the divergence curve (per step: residue types that differ, frame translation
and rotation differences over the diffused residues) as JSON + a markdown summary.

    python tools/same_seed.py [--n-antigen 120] [--num-t 100] [--seed 7] [--tag r02]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def clone(batch):
    out = {}
    for k, v in batch.items():
        if torch.is_tensor(v):
            out[k] = v.clone()
        elif isinstance(v, tuple) and v and torch.is_tensor(v[0]):
            out[k] = tuple(x.clone() for x in v)
        else:
            out[k] = v
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n-antigen', type=int, default=120)
    ap.add_argument('--num-t', type=int, default=100)
    ap.add_argument('--seed', type=int, default=7)
    ap.add_argument('--tag', default='r02')
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    import bench
    from abx_b200 import sampler as S
    from abx_b200.data.synthetic import synthetic_complex
    from abx_b200.diffuser.full_diffuser import FullDiffuser
    from abx_b200.model import features as F_
    from abx_b200.model.abx import ScoreNetwork, get_prev
    from abx_b200.utils.weights import load_seeded_
    cfg = bench.model_config()
    fd = FullDiffuser(cfg['diffuser'])
    model = load_seeded_(ScoreNetwork(cfg['model'], fd), 0).to(dev).eval()
    raw = synthetic_complex(n_antigen=a.n_antigen, seed=0, batch_size=1)
    # ---------------- reference arm: the reference's own ScoreNetwork / get_prev / FullDiffuser on the GPU ----------------
    from oracle import ref_harness, ref_runner
    assert ref_runner.available(), 'oracle/_ref is missing (python oracle/build_ref.py in the build container)'
    ref_harness.REFERENCE_ROOT = ref_runner.REF
    ref_harness.install()
    from abx.model.abx import ScoreNetwork as RefNet, get_prev as ref_get_prev
    from diffuser.full_diffuser import FullDiffuser as RefDiffuser
    import inference as ref_inf
    rcfg, _raw = ref_harness.load_config(cache_dir=os.path.join(ref_runner.REF, 'igso3_cache') + os.sep)
    rfd = RefDiffuser.get(rcfg.diffuser)
    rmodel = load_seeded_(RefNet(rcfg.model, rfd), 0).to(dev).eval()
    # the start state (features + t = 1 prior draw) comes from the REFERENCE's FeatureBuilder (it also carries the keys only the
    # reference's unused heads read, e.g. pseudo_beta); both arms start from clones of it
    from abx.model.features import FeatureBuilder as RefFeatureBuilder
    with open(os.path.join(ref_runner.REF, 'config', 'config_data_feature.json')) as f:
        rfeats = json.load(f)
    for name, args in rfeats:
        if 'device' in args:
            args['device'] = 'cpu'
        if 'diffuse' in name:
            args['diff_conf'] = _raw['diffuser']
            args.pop('optimize_steps', None)
            args['generate_area'] = 'H3'
    torch.manual_seed(a.seed)
    cpu_batch = RefFeatureBuilder(rfeats, is_training=False).build(raw)

    def to_dev(v):
        if torch.is_tensor(v):
            return v.to(dev)
        if isinstance(v, tuple) and v and torch.is_tensor(v[0]):
            return tuple(x.to(dev) for x in v)
        return v
    batch0 = {k: to_dev(v) for k, v in cpu_batch.items()}

    N = batch0['seq'].shape[1]
    grid = np.linspace(1.0 / a.num_t, 1.0, a.num_t)[::-1]
    dt = torch.tensor(1.0 / a.num_t)
    ones = torch.ones(1, device=dev)
    diffuse_mask = (1 - batch0['fixed_mask']) * batch0['atom14_gt_exists'][..., 0]

    def run(arm):
        b = clone(batch0)
        states = []
        torch.manual_seed(a.seed + 1)
        t0 = time.time()
        with torch.no_grad():
            if arm == 'reference':
                b = ref_inf._set_t_feats(b, rfd, grid[0], ones)
                b = ref_inf._self_conditioning(b, rmodel, rcfg.model)
            else:
                b = S._set_t_feats(b, fd, grid[0], ones, with_scalings=False)
                b = S._self_conditioning(b, model, cfg['model'])
            for k, t in enumerate(grid):
                if t > grid[-1]:
                    t_ = torch.tile(torch.tensor(t, device=dev), (1,))                       # float64 (inference.py:216)
                    if arm == 'reference':
                        b = ref_inf._set_t_feats(b, rfd, t_, ones)
                        out = rmodel(b)
                        b.update(ref_get_prev(b, out, rcfg.model))
                        d = rfd
                    else:
                        b = S._set_t_feats(b, fd, t_, ones, with_scalings=False)
                        out = model(b)
                        b.update(get_prev(b, out, cfg['model']))
                        d = fd
                    h = out['heads']
                    rig, seq = d.reverse(rigid_t=b['rigids_t'], seq_t=b['seq_t'], rot_score=h['folding']['rot_score'],
                                         trans_score=h['folding']['trans_score'], logits_t=h['sequence_module']['logits'],
                                         diffuse_mask=diffuse_mask, t=t_, dt=dt, center=True, noise_scale=1.0)
                else:                                                                         # final x0 call (:244-247)
                    out = (rmodel if arm == 'reference' else model)(b)
                    rig, seq = out['heads']['folding']['rigids'], out['heads']['sequence_module']['seq_0']
                b['rigids_t'], b['seq_t'] = rig, seq
                states.append((rig.double().cpu(), seq.long().cpu()))
        torch.cuda.synchronize()
        return states, time.time() - t0

    ref_states, ref_s = run('reference')
    prod_states, prod_s = run('product')
    m = diffuse_mask[0].bool().cpu()
    curve = []
    for k, ((rr, rs), (pr, ps)) in enumerate(zip(ref_states, prod_states)):
        dq = (rr[0, :, :4] * pr[0, :, :4]).sum(-1).abs().clamp(max=1.0)
        ang = 2 * torch.acos(dq)                                                              # rotation between the two frames (rad)
        dx = (rr[0, :, 4:] - pr[0, :, 4:]).norm(dim=-1)
        curve.append(dict(step=k, t=float(grid[k]), seq_mismatch=int((rs[0] != ps[0]).sum()),
                          trans_max_all=float(dx.max()), trans_rms_diffused=float(dx[m].pow(2).mean().sqrt()),
                          trans_max_diffused=float(dx[m].max()), rot_max_diffused_rad=float(ang[m].max())))
    first_seq = next((c['step'] for c in curve if c['seq_mismatch'] > 0), None)
    first_1e4 = next((c['step'] for c in curve if c['trans_max_all'] > 1e-4), None)
    out = dict(n_res=N, num_t=a.num_t, seed=a.seed, diffused_residues=int(m.sum()), reference_arm_seconds=ref_s,
               product_arm_seconds=prod_s, first_step_with_a_different_residue_type=first_seq,
               first_step_with_translation_diff_above_1e_4_A=first_1e4, curve=curve)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', f'same_seed_{a.tag}.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != 'curve'}))
    for c in curve[:3] + curve[9::10]:
        print(c)


if __name__ == '__main__':
    main()
