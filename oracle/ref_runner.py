"""Time the REFERENCE's own sampler modules on the host CPU (measurement infrastructure; used by bench.py's
`--impl reference` arm and `cpu_baseline` leg only).

Runs the unmodified reference code materialised in oracle/_ref (oracle/build_ref.py) through oracle/ref_harness.py:
`ScoreNetwork.forward` (abx/model/abx.py:75-104), `get_prev` (:17-26), `FullDiffuser.reverse`
(diffuser/full_diffuser.py:174-227) and the loop helpers of inference.py:166-251, on the synthetic complex of the
benchmark with the seeded weights every other arm uses, batch 1 (the reference's operating point).
"""
import json
import os
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, '_ref')


def available():
    return os.path.exists(os.path.join(REF, '.complete'))


class ReferenceLoop:
    """The reference's reverse loop, advanced one iteration per `step()` (inference.py:213-251)."""

    def __init__(self, n_antigen, num_t, threads):
        import torch
        from oracle import ref_harness
        ref_harness.REFERENCE_ROOT = REF
        ref_harness.install()
        torch.set_num_threads(threads)
        from abx.model.abx import ScoreNetwork, get_prev
        from abx.model.features import FeatureBuilder
        from diffuser.full_diffuser import FullDiffuser
        import inference as ref_inference
        from abx_b200.data.synthetic import synthetic_complex
        from abx_b200.utils.weights import load_seeded_
        self.torch, self.inf, self.get_prev = torch, ref_inference, get_prev
        self.cfg, raw = ref_harness.load_config(cache_dir=os.path.join(REF, 'igso3_cache') + os.sep)
        self.fd = FullDiffuser.get(self.cfg.diffuser)
        self.model = load_seeded_(ScoreNetwork(self.cfg.model, self.fd), 0).eval()
        with open(os.path.join(REF, 'config', 'config_data_feature.json')) as f:
            feats = json.load(f)
        for name, args in feats:
            if 'device' in args:
                args['device'] = 'cpu'
            if 'diffuse' in name:
                args['diff_conf'] = raw['diffuser']
                args.pop('optimize_steps', None)
                args['generate_area'] = 'H3'
        torch.manual_seed(0)
        self.batch = FeatureBuilder(feats, is_training=False).build(synthetic_complex(n_antigen=n_antigen, seed=0, batch_size=1))
        self.n_res = self.batch['seq'].shape[1]
        bb_mask = self.batch['atom14_gt_exists'][..., 0]
        self.diffuse_mask = (1 - self.batch['fixed_mask']) * bb_mask
        self.ones = torch.ones(1, dtype=torch.float32)
        self.grid = np.linspace(1.0 / num_t, 1.0, num_t)[::-1]                      # inference.py:196-199
        self.dt = torch.tensor(1.0 / num_t)
        self.k = 0
        t0 = time.perf_counter()
        with torch.no_grad():                                                       # self-conditioning warm-up call (:209-211)
            self.batch = self.inf._set_t_feats(self.batch, self.fd, self.grid[0], self.ones)
            self.batch = self.inf._self_conditioning(self.batch, self.model, self.cfg.model)
        self.t_warm = time.perf_counter() - t0

    def step(self):
        """One reverse iteration; returns (seconds in ScoreNetwork.forward + get_prev, seconds in FullDiffuser.reverse)."""
        torch = self.torch
        t = self.grid[self.k % (len(self.grid) - 1)]
        self.k += 1
        with torch.no_grad():
            t0 = time.perf_counter()
            t_ = torch.tile(torch.tensor(t), (1,))
            self.batch = self.inf._set_t_feats(self.batch, self.fd, t_, self.ones)
            out = self.model(self.batch)
            self.batch.update(self.get_prev(self.batch, out, self.cfg.model))
            t1 = time.perf_counter()
            h = out['heads']
            rig, seq = self.fd.reverse(rigid_t=self.batch['rigids_t'], seq_t=self.batch['seq_t'],
                                       rot_score=h['folding']['rot_score'], trans_score=h['folding']['trans_score'],
                                       logits_t=h['sequence_module']['logits'], diffuse_mask=self.diffuse_mask, t=t_, dt=self.dt,
                                       center=True, noise_scale=1.0)
            t2 = time.perf_counter()
            self.batch['rigids_t'], self.batch['seq_t'] = rig, seq
        return t1 - t0, t2 - t1
