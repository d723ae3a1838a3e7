"""Generate tests/golden/*.npz by running the REFERENCE's own modules (imported read-only from
/root/reference through oracle/ref_harness.py) on seeded synthetic inputs.

Build-container only; committed together with its outputs.  The reference has no tests or golden
vectors of its own, so these files are what pins the oracle — and, through the
oracle, the CUDA path — to the reference.

    python oracle/make_golden.py [geometry igso3 scores reverse reverse_edges prior marginal ipa ipascore model sampler]
"""
import contextlib
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402

ref_harness.install()
import torch  # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
CACHE = os.environ.get('ABX_IGSO3_CACHE', '/tmp/abx_cache/')


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if torch.is_tensor(v):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(GOLDEN, name + '.npz')
    np.savez_compressed(path, **out)
    print(f'  wrote {path}  {os.path.getsize(path) / 1024:.1f} KiB  keys={list(out)}')


@contextlib.contextmanager
def record_rng():
    """Record every torch.randn / rand / randint / poisson draw made inside the block, in order."""
    log = []
    orig = {n: getattr(torch, n) for n in ('randn', 'rand', 'randint', 'poisson', 'normal')}

    def wrap(n):
        def f(*a, **k):
            out = orig[n](*a, **k)
            log.append((n, out.clone()))
            return out
        return f
    for n in orig:
        setattr(torch, n, wrap(n))
    try:
        yield log
    finally:
        for n, f in orig.items():
            setattr(torch, n, f)


_diffuser = None


def get_diffuser():
    global _diffuser
    if _diffuser is None:
        from diffuser.full_diffuser import FullDiffuser
        cfg, _ = ref_harness.load_config(cache_dir=CACHE)
        t0 = time.time()
        _diffuser = FullDiffuser.get(cfg.diffuser)
        print(f'  reference FullDiffuser ready ({time.time() - t0:.1f}s)')
    return _diffuser


# ------------------------------------------------------------------------------------------------
def gen_geometry():
    from abx.model import quat_affine as qa, r3
    g = torch.Generator().manual_seed(11)
    q = torch.randn(64, 4, generator=g)
    q = q / q.norm(dim=-1, keepdim=True)
    q[:4] = torch.tensor([[1., 0, 0, 0], [-1., 0, 0, 0], [0.9999999, 1e-8, 0, 0], [0., 1, 0, 0]])
    q2 = torch.randn(64, 4, generator=g)
    q2 = q2 / q2.norm(dim=-1, keepdim=True)
    v = torch.randn(64, 3, generator=g)
    v[:3] = torch.tensor([[0., 0, 0], [1e-8, 0, 0], [3.1, 0.2, 0]])
    rots = qa.quat_to_rot(q)
    pts = torch.randn(64, 5, 3, generator=g)
    trans = torch.randn(64, 3, generator=g)
    p3 = torch.randn(3, 64, 3, generator=g)
    fr = r3.rigids_from_3_points(p3[0], p3[1], p3[2])
    save('geometry',
         q=q, q2=q2, v=v, pts=pts, trans=trans, p3=p3,
         quat_to_rot=rots, quat_multiply=qa.quat_multiply(q, q2), quat_multiply_by_vec=qa.quat_multiply_by_vec(q, v),
         quat_precompose_vec=qa.quat_precompose_vec(q, v * 0.3), invert_quat=qa.invert_quat(q * 1.7),
         quat_to_rotvec=qa.quat_to_rotvec(q), rotvec_to_quat=qa.rotvec_to_quat(v),
         quat_to_rotvec_f64=qa.quat_to_rotvec(q.double()), rotvec_to_quat_f64=qa.rotvec_to_quat(v.double()),
         rot_to_quat=qa.rot_to_quat(rots),
         rigids_apply=r3.rigids_apply((rots, trans), pts),
         invert_rots=r3.invert_rigids((rots, trans))[0], invert_trans=r3.invert_rigids((rots, trans))[1],
         frame_rots=fr[0], frame_trans=fr[1], tensor7=r3.rigids_to_tensor7(fr))


ROWS = [0, 1, 9, 19, 100, 250, 499, 750, 989, 999]
T_USED = [1.0, 0.99, 0.98, 0.5, 0.02, 0.01, 0.31, 0.7, 0.2, 0.9, 0.3]      # every t any golden uses


def gen_igso3():
    d = get_diffuser()._so3_diffuser
    global ROWS
    ROWS = sorted(set(ROWS + d.t_to_idx(torch.tensor(T_USED, dtype=torch.float64))))
    save('igso3', rows=np.array(ROWS), discrete_sigma=d.discrete_sigma, discrete_omega=d.discrete_omega,
         pdf=d._pdf[ROWS], cdf=d._cdf[ROWS], score_norms=d._score_norms[ROWS], score_scaling=d._score_scaling)


def gen_scores():
    fd = get_diffuser()
    g = torch.Generator().manual_seed(12)
    B, N = 3, 40
    q0 = torch.randn(B, N, 4, generator=g); q0 = q0 / q0.norm(dim=-1, keepdim=True)
    dq = torch.randn(B, N, 4, generator=g) * torch.tensor([1.0, 0.5, 0.05])[:, None, None]
    dq[..., 0] = 1.0
    dq = dq / dq.norm(dim=-1, keepdim=True)
    from abx.model.quat_affine import quat_multiply
    qt = quat_multiply(q0, dq)
    qt[0, :3] = q0[0, :3]                       # identical frames -> zero rotvec branch
    x0 = torch.randn(B, N, 3, generator=g) * 10
    xt = x0 + torch.randn(B, N, 3, generator=g) * 3
    out = {}
    for tag, t in (('a', [1.0, 0.99, 0.5]), ('b', [0.02, 0.01, 0.31])):
        t32 = torch.tensor(t, dtype=torch.float32)
        t64 = torch.tensor(t, dtype=torch.float64)
        out[f't_{tag}'] = t64
        out[f'rot_score_{tag}'] = fd.calc_quat_score(qt, q0, t32)
        out[f'rot_score_t64_{tag}'] = fd.calc_quat_score(qt, q0, t64)
        out[f'trans_score_{tag}'] = fd.calc_trans_score(xt, x0, t32)
        out[f'trans_score_t64_{tag}'] = fd.calc_trans_score(xt, x0, t64)
        rs, ts = fd.score_scaling(t32)
        out[f'rot_score_scaling_{tag}'] = rs
        out[f'trans_score_scaling_{tag}'] = ts
        out[f'sigma_idx_{tag}'] = np.array(fd._so3_diffuser.t_to_idx(t64))
        out[f'diffusion_coef_{tag}'] = fd._so3_diffuser.diffusion_coef(t64)
    grid = np.linspace(0.01, 1.0, 100)
    out['grid_sigma_idx'] = np.array(fd._so3_diffuser.t_to_idx(torch.tensor(grid)))
    save('scores', q0=q0, qt=qt, x0=x0, xt=xt, **out)


def gen_reverse():
    fd = get_diffuser()
    g = torch.Generator().manual_seed(13)
    B, N = 2, 48
    out = {}
    for tag, tval, f64_state in (('t99', 0.99, False), ('t50', 0.5, True), ('t02', 0.02, True)):
        q = torch.randn(B, N, 4, generator=g); q = q / q.norm(dim=-1, keepdim=True)
        x = torch.randn(B, N, 3, generator=g) * 12
        rigid_t = torch.cat([q, x], dim=-1)
        if f64_state:
            rigid_t = rigid_t.double()          # state after the first reverse step is float64
        seq_t = torch.randint(0, 20, (B, N), generator=g)
        seq_t[0, 0] = 20                        # padding / unknown index gets clamped
        rot_score = torch.randn(B, N, 3, generator=g) * 0.7
        trans_score = (torch.randn(B, N, 3, generator=g) * 0.5).double()
        logits = torch.randn(B, N, 20, generator=g) * 2
        mask = (torch.rand(B, N, generator=g) < 0.4).to(torch.int32)
        t = torch.tile(torch.tensor(np.float64(tval)), (B,))           # float64, as inference.py:216
        dt = torch.tensor(1 / 100)
        torch.manual_seed(100 + int(tval * 100))
        with record_rng() as log:
            rigids_1, seq_1 = fd.reverse(rigid_t=rigid_t, seq_t=seq_t, rot_score=rot_score, trans_score=trans_score,
                                         logits_t=logits, t=t, dt=dt, diffuse_mask=mask, center=True, noise_scale=1.0)
        kinds = [k for k, _ in log]
        assert kinds == ['randn', 'randn', 'poisson'], kinds
        out.update({f'{tag}_rigid_t': rigid_t, f'{tag}_seq_t': seq_t, f'{tag}_rot_score': rot_score,
                    f'{tag}_trans_score': trans_score, f'{tag}_logits': logits, f'{tag}_mask': mask,
                    f'{tag}_t': t, f'{tag}_z_rot': log[0][1], f'{tag}_z_trans': log[1][1], f'{tag}_jumps': log[2][1],
                    f'{tag}_rigids_1': rigids_1, f'{tag}_seq_1': seq_1})
        # the Poisson rates themselves (reverse_rates * dt) for the rate-parity test
        sd = fd._seq_diffuser
        orig = torch.distributions.poisson.Poisson.__init__
        captured = {}

        def spy(self, rate, *a, **k):
            captured['rate'] = rate.clone()
            orig(self, rate, *a, **k)
        torch.distributions.poisson.Poisson.__init__ = spy
        try:
            sd.reverse(x_t=seq_t, logits_t=logits, t=t, dt=dt)
        finally:
            torch.distributions.poisson.Poisson.__init__ = orig
        out[f'{tag}_rate_dt'] = captured['rate']
    save('reverse', **out)


def gen_reverse_edges():
    """FullDiffuser.reverse on the corner cases of the quaternion <-> rotation-vector maps and of the tau-leap:
    identity / w < 0 / w = 0 (angle pi) / 1e-8 rad input rotations, zero rotation perturbation (zero score and zero
    noise: the |theta| < 1e-6 series of rotvec_to_quat, quat_affine.py:133-150), saturated logits, residue indices 0 /
    19 / 20, and multiple jumps per residue that leave [0, 19] before the clamp (discrete_diffuser.py:181-187)."""
    fd = get_diffuser()
    g = torch.Generator().manual_seed(31)
    B, N = 2, 16
    q = torch.randn(B, N, 4, generator=g, dtype=torch.float64); q = q / q.norm(dim=-1, keepdim=True)
    eps = 1e-8
    special = torch.tensor([[1.0, 0, 0, 0], [1.0, 0, 0, 0], [-0.6, 0.8, 0, 0], [-0.6, 0, 0.8, 0], [0, 1.0, 0, 0], [0, 0, 0.6, 0.8],
                            [math.cos(eps / 2), math.sin(eps / 2), 0, 0], [math.cos(eps / 2), 0, 0, -math.sin(eps / 2)]],
                           dtype=torch.float64)
    q[:, :8] = special
    x = torch.randn(B, N, 3, generator=g, dtype=torch.float64) * 12
    rigid_t = torch.cat([q, x], dim=-1)
    mask = torch.zeros(B, N, dtype=torch.int32)
    mask[:, 1::2] = 1                                   # odd residues diffused, even ones fixed
    mask[1, 8:] = 1
    seq_t = torch.randint(0, 20, (B, N), generator=g)
    seq_t[0, 0], seq_t[0, 1], seq_t[0, 2], seq_t[0, 3] = 20, 20, 0, 19
    rot_score = torch.randn(B, N, 3, generator=g) * 0.7
    rot_score[:, 8:12] = 0                              # with zero noise below: perturbation exactly zero
    trans_score = (torch.randn(B, N, 3, generator=g) * 0.5).double()
    logits = torch.randn(B, N, 20, generator=g) * 2
    logits[0, 4] = -60.0; logits[0, 4, 7] = 60.0        # saturated softmax
    logits[0, 5] = 0.0                                  # uniform
    t = torch.tile(torch.tensor(np.float64(0.5)), (B,))
    dt = torch.tensor(1 / 100)
    jumps = torch.zeros(B, N, 20)
    jumps[0, 3, 0] = 2; jumps[0, 3, 5] = 1              # 19 + 2 (0 - 19) + (5 - 19) < 0 -> clamp to 0
    jumps[0, 2, 19] = 3                                 # 0 + 3 * 19 > 19 -> clamp to 19
    jumps[1, 9, 4] = 1; jumps[1, 11, 11] = 1; jumps[0, 1, 6] = 1
    orig_randn, orig_poisson = torch.randn, torch.poisson
    drawn = []

    def randn(*a, **k):
        z = orig_randn(*a, **k)
        if len(drawn) == 0:
            z[:, 8:12] = 0                              # rotation noise of residues 8..11
        drawn.append(z.clone())
        return z
    torch.randn, torch.poisson = randn, (lambda rate, *a, **k: jumps.to(rate.dtype))
    try:
        torch.manual_seed(77)
        rigids_1, seq_1 = fd.reverse(rigid_t=rigid_t, seq_t=seq_t, rot_score=rot_score, trans_score=trans_score,
                                     logits_t=logits, t=t, dt=dt, diffuse_mask=mask, center=True, noise_scale=1.0)
    finally:
        torch.randn, torch.poisson = orig_randn, orig_poisson
    assert len(drawn) == 2 and torch.isfinite(rigids_1).all()
    save('reverse_edges', rigid_t=rigid_t, seq_t=seq_t, rot_score=rot_score, trans_score=trans_score, logits=logits, mask=mask,
         t=t, z_rot=drawn[0], z_trans=drawn[1], jumps=jumps, rigids_1=rigids_1, seq_1=seq_1)


def gen_prior():
    fd = get_diffuser()
    g = torch.Generator().manual_seed(14)
    B, N = 2, 40
    q = torch.randn(B, N, 4, generator=g); q = q / q.norm(dim=-1, keepdim=True)
    x = torch.randn(B, N, 3, generator=g) * 12
    rig = torch.cat([q, x], dim=-1)
    seq = torch.randint(0, 20, (B, N), generator=g)
    mask = (torch.rand(B, N, generator=g) < 0.5).to(torch.int32)
    torch.manual_seed(7)
    with record_rng() as log:
        ret = fd.sample_ref(n_samples=rig.shape[:2], impute_rigids=rig, impute_seq=seq, diffuse_mask=mask)
    kinds = [k for k, _ in log]
    assert kinds == ['randn', 'rand', 'randn', 'randint'], kinds
    save('prior', impute_rigids=rig, impute_seq=seq, mask=mask, z_rot=log[0][1], u_rot=log[1][1], z_trans=log[2][1],
         seq_rand=log[3][1], rigids_t=ret['rigids_t'], seq_t=ret['seq_t'])


def gen_marginal():
    """FullDiffuser.forward_marginal (full_diffuser.py:57-126): the optimize-mode start state.  Draw order of the
    reference: randn [B,N,3] + rand [B,N] (SO3Diffuser.sample), normal [B,N,3] (R3Diffuser.forward_marginal:101),
    three multinomial draws (DiscreteDiffuser.forward_marginal: x_t, the perturbed position, its new value)."""
    fd = get_diffuser()
    g = torch.Generator().manual_seed(21)
    B, N = 3, 37
    q = torch.randn(B, N, 4, generator=g); q = q / q.norm(dim=-1, keepdim=True)
    x = torch.randn(B, N, 3, generator=g) * 12
    rig = torch.cat([q, x], dim=-1)
    seq = torch.randint(0, 20, (B, N), generator=g)
    mask = (torch.rand(B, N, generator=g) < 0.5).to(torch.int32)
    t = torch.tensor([0.02, 0.2, 0.7])            # sigma rows 19, 199, 699: all held by tests/golden/igso3.npz

    log = []
    orig = {n: getattr(torch, n) for n in ('randn', 'rand', 'normal', 'multinomial')}

    def wrap(n):
        def f(*a, **k):
            if n == 'normal':                              # recover the unit draw behind mean + std * z
                state = torch.get_rng_state()
                out = orig[n](*a, **k)
                after = torch.get_rng_state()
                torch.set_rng_state(state)
                z = orig['randn'](out.shape)
                torch.set_rng_state(after)
                assert torch.allclose(k['mean'] + k['std'] * z, out, atol=1e-6), 'normal() is not mean + std * randn here'
                log.append((n, z))
                return out
            out = orig[n](*a, **k)
            log.append((n, out.clone()))
            return out
        return f
    for n in orig:
        setattr(torch, n, wrap(n))
    try:
        torch.manual_seed(11)
        ret = fd.forward_marginal(rigids_0=rig, seq_0=seq, t=t, diffuse_mask=mask)
        torch.manual_seed(12)
        ret_nomask = fd.forward_marginal(rigids_0=rig, seq_0=seq, t=t, diffuse_mask=None)
    finally:
        for n, f in orig.items():
            setattr(torch, n, f)
    kinds = [k for k, _ in log]
    assert kinds == ['randn', 'rand', 'normal', 'multinomial', 'multinomial', 'multinomial'] * 2, kinds
    out = dict(rigids_0=rig, seq_0=seq, mask=mask, t=t)
    for tag, r, lg in (('m', ret, log[:6]), ('n', ret_nomask, log[6:])):
        out.update({f'{tag}_z_rot': lg[0][1], f'{tag}_u_rot': lg[1][1], f'{tag}_z_trans': lg[2][1],
                    f'{tag}_x_t': lg[3][1].reshape(B, N), f'{tag}_dims': lg[4][1].reshape(B), f'{tag}_newval': lg[5][1].reshape(B)})
        out.update({f'{tag}_{k}': v for k, v in r.items()})
    save('marginal', **out)


# ------------------------------------------------------------------------------------------------
def _ref_features(n_antigen=9, batch_size=2, generate_area='H3', seed=0):
    from abx.model.features import FeatureBuilder
    from abx_b200.data.synthetic import small_complex
    _, raw = ref_harness.load_config(cache_dir=CACHE)
    get_diffuser()
    with open(os.path.join(ref_harness.REFERENCE_ROOT, 'config', 'config_data_feature.json')) as f:
        feats = json.load(f)
    for name, args in feats:
        if 'device' in args:
            args['device'] = 'cpu'
        if 'diffuse' in name:
            args['diff_conf'] = raw['diffuser']
            args.pop('optimize_steps', None)
            args['generate_area'] = generate_area
    torch.manual_seed(1000 + seed)
    return FeatureBuilder(feats, is_training=False).build(small_complex(n_antigen=n_antigen, seed=seed, batch_size=batch_size))


def _ref_model():
    from abx.model.abx import ScoreNetwork
    from abx_b200.utils.weights import load_seeded_
    cfg, _ = ref_harness.load_config(cache_dir=CACHE)
    model = ScoreNetwork(cfg.model, get_diffuser())
    load_seeded_(model, 0)
    return model.eval(), cfg


MODEL_BATCH_KEYS = ('seq', 'mask', 'atom14_gt_positions', 'atom14_gt_exists', 'cdr_def', 'chain_id', 'residx',
                    'anchor_flag', 'residx_atom37_to_atom14', 'torsion_angles_sin_cos', 'pseudo_beta', 'pseudo_beta_mask',
                    'rigids_t', 'rigids_0', 'seq_t', 't', 'fixed_mask', 'struc_loss_mask')


def _batch_arrays(batch, prefix='batch_'):
    out = {prefix + k: batch[k] for k in MODEL_BATCH_KEYS}
    out[prefix + 'gt_frame_rots'] = batch['rigidgroups_gt_frames'][0]
    out[prefix + 'gt_frame_trans'] = batch['rigidgroups_gt_frames'][1]
    return out


def gen_ipa():
    """InvariantPointAttention.forward alone (folding.py:47-132), B=2, N=37 (ragged: batch 1 has 5 masked keys)."""
    model, cfg = _ref_model()
    ipa = model.impl.diffusion_module.ScoreNetwork.attention_module
    from abx.model.quat_affine import quat_to_rot
    from abx_b200.utils.weights import np_randn
    g = torch.Generator().manual_seed(21)
    B, N = 2, 37
    x = np_randn(211, B, N, 256)              # regenerated from the seed by the tests (not stored)
    z = np_randn(212, B, N, N, 128)
    q = torch.randn(B, N, 4, generator=g); q = q / q.norm(dim=-1, keepdim=True)
    trans = torch.randn(B, N, 3, generator=g) * 1.5           # nm units (Å / position_scale)
    mask = torch.ones(B, N)
    mask[1, -5:] = 0
    with torch.no_grad():
        out = ipa(inputs_1d=x, inputs_2d=z, mask=mask, in_rigids=(quat_to_rot(q), trans))
    save('ipa', quat=q, rots=quat_to_rot(q), trans=trans, mask=mask, out=out)


def gen_ipascore():
    """IpaScore.forward (score_network.py:83-196) on trunk-shaped random inputs, N=64."""
    model, cfg = _ref_model()
    batch = _ref_features()
    B, N = batch['seq'].shape
    from abx_b200.utils.weights import np_randn
    rep = {'seq': np_randn(221, B, N, 544), 'pair': np_randn(222, B, N, N, 192)}   # regenerated by the tests
    batch['t'] = torch.tensor([0.7, 0.2])
    with torch.no_grad():
        out = model.impl.diffusion_module.ScoreNetwork(rep, batch)
    save('ipascore', **_batch_arrays(batch),
         rot_score=out['rot_score'], trans_score=out['trans_score'], rigids=out['rigids'],
         structure_module=out['representations']['structure_module'],
         angles_sin_cos=out['sidechains'][-1]['angles_sin_cos'],
         traj_trans=torch.stack([t for _, t in out['traj']]), traj_rots=torch.stack([r for r, _ in out['traj']]))


def gen_model():
    """ScoreNetwork.forward (abx.py:75-104: 2 recycles + final pass) on the small complex."""
    model, cfg = _ref_model()
    batch = _ref_features()
    arrays = _batch_arrays(batch)
    batch['t'] = torch.tensor([0.9, 0.3])
    arrays['batch_t'] = batch['t']
    batch['is_recycling'] = False
    with torch.no_grad():
        # one trunk pass alone first (seqformer.py:170-226), from zero self-conditioning
        B, N = batch['seq'].shape
        b1 = dict(batch)
        b1.update(prev_pos=torch.zeros(B, N, N, dtype=torch.int64), prev_seq=torch.zeros(B, N, 544),
                  prev_pair=torch.zeros(B, N, N, 192), is_recycling=True)
        seq_act, pair_act = model.impl.seqformer(b1)
        out = model(batch)
    h = out['heads']
    from abx.model.abx import get_prev
    prev = get_prev(batch, out, cfg.model)
    save('model', **arrays, trunk_seq=seq_act, trunk_pair=pair_act[:, :8, :8],
         rot_score=h['folding']['rot_score'], trans_score=h['folding']['trans_score'], rigids=h['folding']['rigids'],
         atom14=h['folding']['final_atom14_positions'], atom37=h['folding']['final_atom_positions'],
         logits=h['sequence_module']['logits'], seq_0=h['sequence_module']['seq_0'], pLDDT=h['predicted_lddt']['pLDDT'],
         rep_seq=out['representations']['seq'], rep_pair_corner=out['representations']['pair'][:, :8, :8],
         seq_t_after=batch['seq_t'], prev_pos=prev['prev_pos'])


def gen_sampler():
    """Three iterations of the reverse loop (inference.py:209-251) with recorded noise: warm-up
    self-conditioning call at t=1, then steps at t=1.0, 0.99 (model + reverse)."""
    from abx.model.abx import get_prev
    import inference as ref_inference
    model, cfg = _ref_model()
    fd = get_diffuser()
    batch = _ref_features(batch_size=1)
    arrays = _batch_arrays(batch)
    device = 'cpu'
    bb_mask = batch['atom14_gt_exists'][..., 0]
    diffuse_mask = (1 - batch['fixed_mask']) * bb_mask
    ones = torch.ones(batch['rigids_t'].shape[0], dtype=torch.float32)
    steps = np.linspace(0.01, 1.0, 100)[::-1]
    dt = torch.tensor(1 / 100)
    out = {}
    torch.manual_seed(5)
    with torch.no_grad():
        batch = ref_inference._set_t_feats(batch, fd, steps[0], ones)
        batch = ref_inference._self_conditioning(batch, model, cfg.model)
        for k, t in enumerate(steps[:2]):
            t_ = torch.tile(torch.tensor(t), (batch['rigids_t'].shape[0],))
            batch = ref_inference._set_t_feats(batch, fd, t_, ones)
            mo = model(batch)
            batch.update(get_prev(batch, mo, cfg.model))
            with record_rng() as log:
                rig, seq = fd.reverse(rigid_t=batch['rigids_t'], seq_t=batch['seq_t'],
                                      rot_score=mo['heads']['folding']['rot_score'],
                                      trans_score=mo['heads']['folding']['trans_score'],
                                      logits_t=mo['heads']['sequence_module']['logits'], diffuse_mask=diffuse_mask,
                                      t=t_, dt=dt, center=True, noise_scale=1.0)
            out.update({f's{k}_z_rot': log[0][1], f's{k}_z_trans': log[1][1], f's{k}_jumps': log[2][1],
                        f's{k}_rigids': rig, f's{k}_seq': seq, f's{k}_rot_score': mo['heads']['folding']['rot_score'],
                        f's{k}_trans_score': mo['heads']['folding']['trans_score'],
                        f's{k}_logits': mo['heads']['sequence_module']['logits'],
                        f's{k}_pLDDT': mo['heads']['predicted_lddt']['pLDDT'],
                        f's{k}_atom14': mo['heads']['folding']['final_atom14_positions']})
            batch['rigids_t'], batch['seq_t'] = rig, seq
    save('sampler', **arrays, diffuse_mask=diffuse_mask, **out)


def _to64(x):
    if torch.is_tensor(x):
        return x.double() if x.is_floating_point() else x
    if isinstance(x, tuple):
        return tuple(_to64(v) for v in x)
    return x


def _pair_digest(pair, pos):
    """Compact check data of a [1,N,N,C] pair tensor: the channel vectors at sampled (i,j) positions plus the sums
    over j, over i and over the diagonal band — any transposed / shifted / tile-local indexing error moves them."""
    p = pair[0]
    return dict(samples=p[pos[:, 0], pos[:, 1]], row_sum=p.sum(dim=1), col_sum=p.sum(dim=0))


def gen_big(n_antigen, name):
    """BASELINE-size parity data (N = 230 + n_antigen, B = 1, H3 diffused) with a float64 error budget:
      r32_*  the REFERENCE's own modules in float32 (what the product must match),
      o64_*  oracle/model.py evaluated in float64 on the same inputs and weights ("exact" value of the arithmetic),
    for (a) one trunk pass from zero self-conditioning (seqformer.py:170-226), (b) IpaScore on seeded random
    representations (score_network.py:83-196), (c) ScoreNetwork.forward (abx.py:75-104).  |r32 - o64| is the float32
    noise of the reference itself; the GPU tests bound |product - o64| by a small multiple of it."""
    from abx.model.features import FeatureBuilder
    from abx.model.abx import get_prev
    from abx_b200.data.synthetic import synthetic_complex
    from abx_b200.utils.weights import np_randn
    from oracle import model as OM
    from oracle.diffusers import OracleDiffuser
    model, cfg = _ref_model()
    _, raw = ref_harness.load_config(cache_dir=CACHE)
    fd = get_diffuser()
    with open(os.path.join(ref_harness.REFERENCE_ROOT, 'config', 'config_data_feature.json')) as f:
        feats = json.load(f)
    for fname, args in feats:
        if 'device' in args:
            args['device'] = 'cpu'
        if 'diffuse' in fname:
            args['diff_conf'] = raw['diffuser']
            args.pop('optimize_steps', None)
            args['generate_area'] = 'H3'
    torch.manual_seed(4000 + n_antigen)
    batch = FeatureBuilder(feats, is_training=False).build(synthetic_complex(n_antigen=n_antigen, seed=3, batch_size=1))
    batch['t'] = torch.tensor([0.6])
    batch['is_recycling'] = False
    arrays = _batch_arrays(batch)
    B, N = batch['seq'].shape
    pos = torch.from_numpy(np.random.default_rng(99).integers(0, N, size=(1536, 2)))
    out = dict(arrays, pair_pos=pos)
    P32 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    P64 = {k: _to64(v) for k, v in P32.items()}
    od = OracleDiffuser(fd._so3_diffuser._score_norms, cdf=fd._so3_diffuser._cdf, pdf=None)

    def zero_prev(b, dt):
        b = dict(b)
        b.update(prev_pos=torch.zeros(B, N, N, dtype=torch.int64), prev_seq=torch.zeros(B, N, 544, dtype=dt),
                 prev_pair=torch.zeros(B, N, N, 192, dtype=dt), is_recycling=True)
        return b

    rep_seq, rep_pair = np_randn(921, B, N, 544), np_randn(922, B, N, N, 192)      # regenerated by the tests
    t0 = time.time()
    with torch.no_grad():
        # ---- reference, float32
        s32, p32 = model.impl.seqformer(zero_prev(batch, torch.float32))
        ipa32 = model.impl.diffusion_module.ScoreNetwork({'seq': rep_seq, 'pair': rep_pair}, dict(batch))
        bm = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}
        m32 = model(bm)
        prev32 = get_prev(bm, m32, cfg.model)
        print(f'  reference float32 done ({time.time() - t0:.0f}s)')
        # ---- oracle, float64
        OM.set_compute_dtype(torch.float64)
        try:
            b64 = {k: _to64(v) for k, v in batch.items()}
            z = zero_prev(b64, torch.float64)
            s64, p64 = OM.embed_inputs(P64, z)
            s64, p64 = OM.seqformer_block(P64, s64, p64, z['mask'])
            ipa64 = OM.ipascore_forward(P64, od, rep_seq.double(), rep_pair.double(), dict(b64))
            bm64 = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in b64.items()}
            m64 = OM.score_network(P64, od, bm64)
        finally:
            OM.set_compute_dtype(torch.float32)
        print(f'  oracle float64 done ({time.time() - t0:.0f}s)')
    h = m32['heads']
    out.update(r32_trunk_seq=s32, o64_trunk_seq=s64.float())
    out.update({'r32_trunk_pair_' + k: v for k, v in _pair_digest(p32, pos).items()})
    out.update({'o64_trunk_pair_' + k: v.float() for k, v in _pair_digest(p64, pos).items()})
    for tag, o in (('r32', ipa32), ('o64', ipa64)):
        sm = o['representations']['structure_module'] if 'representations' in o else o['structure_module']
        ang = o['sidechains'][-1]['angles_sin_cos'] if 'sidechains' in o else o['angles_sin_cos']
        out.update({f'{tag}_ipa_structure_module': sm.float(), f'{tag}_ipa_rigids': o['rigids'].float(),
                    f'{tag}_ipa_angles': ang.float(), f'{tag}_ipa_trans_score': o['trans_score'].float(),
                    f'{tag}_ipa_rot_score': o['rot_score'].float()})
    out.update(r32_rigids=h['folding']['rigids'], r32_atom14=h['folding']['final_atom14_positions'],
               r32_trans_score=h['folding']['trans_score'], r32_rot_score=h['folding']['rot_score'],
               r32_logits=h['sequence_module']['logits'], r32_seq_0=h['sequence_module']['seq_0'],
               r32_pLDDT=h['predicted_lddt']['pLDDT'], r32_rep_seq=m32['representations']['seq'],
               r32_seq_t_after=bm['seq_t'], r32_prev_pos=prev32['prev_pos'])
    out.update({'r32_rep_pair_' + k: v for k, v in _pair_digest(m32['representations']['pair'], pos).items()})
    out.update(o64_rigids=m64['rigids'].float(), o64_atom14=m64['atom14'].float(), o64_trans_score=m64['trans_score'].float(),
               o64_rot_score=m64['rot_score'].float(), o64_logits=m64['logits'].float(), o64_seq_0=m64['seq_0'],
               o64_pLDDT=m64['pLDDT'].float(), o64_rep_seq=m64['rep_seq'].float())
    out.update({'o64_rep_pair_' + k: v.float() for k, v in _pair_digest(m64['rep_pair'], pos).items()})
    for k in ('trunk_seq', 'rigids', 'atom14', 'trans_score', 'logits', 'pLDDT', 'rep_seq', 'ipa_structure_module', 'ipa_rigids'):
        d = (out['r32_' + k].double() - out['o64_' + k].double()).abs().max()
        print(f'  |reference32 - oracle64| {k:22s} {float(d):.3e}   (scale {float(out["o64_" + k].abs().max()):.3g})')
    save(name, **out)


def gen_model_n350():
    gen_big(120, 'model_n350')


def gen_model_n262():
    gen_big(32, 'model_n262')


ALL = dict(geometry=gen_geometry, igso3=gen_igso3, scores=gen_scores, reverse=gen_reverse, reverse_edges=gen_reverse_edges, prior=gen_prior, marginal=gen_marginal,
           ipa=gen_ipa, ipascore=gen_ipascore, model=gen_model, sampler=gen_sampler,
           model_n350=gen_model_n350, model_n262=gen_model_n262)

if __name__ == '__main__':
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for name in (sys.argv[1:] or list(ALL)):
        print(f'[{name}]')
        ALL[name]()
