"""The oracle (oracle/quat.py, oracle/diffusers.py) against golden vectors produced by the reference's
own modules (oracle/make_golden.py).  CPU only."""
import numpy as np
import torch

from oracle import diffusers as D
from oracle import quat as Q
from tests.util import golden, maxabs, oracle_diffuser


def test_quaternion_algebra_matches_reference():
    g = golden('geometry')
    q, q2, v = g['q'], g['q2'], g['v']
    assert maxabs(Q.quat_to_rot(q), g['quat_to_rot']) < 1e-6
    assert maxabs(Q.quat_multiply(q, q2), g['quat_multiply']) < 1e-6
    assert maxabs(Q.quat_multiply_by_vec(q, v), g['quat_multiply_by_vec']) < 1e-6
    assert maxabs(Q.quat_precompose_vec(q, v * 0.3), g['quat_precompose_vec']) < 1e-6
    assert maxabs(Q.invert_quat(q * 1.7), g['invert_quat']) < 1e-6
    assert maxabs(Q.quat_to_rotvec(q), g['quat_to_rotvec']) < 1e-6
    assert maxabs(Q.rotvec_to_quat(v), g['rotvec_to_quat']) < 1e-6
    assert maxabs(Q.quat_to_rotvec(q.double()), g['quat_to_rotvec_f64']) < 1e-12
    assert maxabs(Q.rotvec_to_quat(v.double()), g['rotvec_to_quat_f64']) < 1e-12
    assert maxabs(Q.rot_to_quat(g['quat_to_rot']), g['rot_to_quat']) < 1e-6


def test_rigid_algebra_matches_reference():
    g = golden('geometry')
    rig = (g['quat_to_rot'], g['trans'])
    assert maxabs(Q.rigids_apply(rig, g['pts']), g['rigids_apply']) < 1e-5
    inv = Q.invert_rigids(rig)
    assert maxabs(inv[0], g['invert_rots']) == 0 and maxabs(inv[1], g['invert_trans']) < 1e-6
    fr = Q.rigids_from_3_points(*g['p3'])
    assert maxabs(fr[0], g['frame_rots']) < 1e-6 and maxabs(fr[1], g['frame_trans']) == 0
    assert maxabs(Q.rigids_to_tensor7(fr), g['tensor7']) < 1e-5


def test_igso3_grids_and_rows():
    g = golden('igso3')
    assert maxabs(D.so3_discrete_sigma(), g['discrete_sigma']) == 0
    assert maxabs(D.so3_discrete_omega(), g['discrete_omega']) == 0
    rows = g['rows'].long()
    # well-conditioned rows (sigma >= ~0.5): the series converges and the tables agree tightly
    for k, r in enumerate(rows.tolist()):
        if r < 250:
            continue
        _, pdf, cdf, sc = D.igso3_table_rows(D.so3_discrete_sigma()[r:r + 1])
        assert maxabs(pdf[0], g['pdf'][k]) < 1e-4 * float(g['pdf'][k].abs().max())
        assert maxabs(cdf[0], g['cdf'][k]) < 1e-4
        assert maxabs(sc[0], g['score_norms'][k]) < 2e-3 * float(g['score_norms'][k].abs().max())


def test_sigma_index_on_the_sampling_grid():
    g = golden('scores')
    grid = torch.tensor(np.linspace(0.01, 1.0, 100))
    idx = D.so3_t_to_idx(grid, D.so3_discrete_sigma())
    assert torch.equal(idx, g['grid_sigma_idx'].long())
    for tag in 'ab':
        assert torch.equal(D.so3_t_to_idx(g[f't_{tag}'], D.so3_discrete_sigma()), g[f'sigma_idx_{tag}'].long())
        assert maxabs(D.so3_diffusion_coef(g[f't_{tag}']), g[f'diffusion_coef_{tag}']) < 1e-12


def test_scores_match_reference():
    g = golden('scores')
    od = oracle_diffuser()
    for tag in 'ab':
        t64 = g[f't_{tag}']
        t32 = t64.float()
        assert maxabs(od.calc_quat_score(g['qt'], g['q0'], t32), g[f'rot_score_{tag}']) < 1e-6
        assert maxabs(od.calc_quat_score(g['qt'], g['q0'], t64), g[f'rot_score_t64_{tag}']) < 1e-6
        assert maxabs(od.calc_trans_score(g['xt'], g['x0'], t32), g[f'trans_score_{tag}']) < 1e-5
        r64 = od.calc_trans_score(g['xt'], g['x0'], t64)
        assert r64.dtype == torch.float64 and maxabs(r64, g[f'trans_score_t64_{tag}']) < 1e-10
        rs, ts = od.score_scaling(t32)
        assert maxabs(ts, g[f'trans_score_scaling_{tag}']) < 1e-5
        assert maxabs(rs, g[f'rot_score_scaling_{tag}']) < 1e-4 * float(g[f'rot_score_scaling_{tag}'].abs().max())


def test_reverse_step_matches_reference():
    g = golden('reverse')
    od = oracle_diffuser()
    for tag in ('t99', 't50', 't02'):
        p = {k[len(tag) + 1:]: v for k, v in g.items() if k.startswith(tag + '_')}
        rr, _ = od.reverse_rates(p['seq_t'], p['logits'], p['t'])
        dt = torch.tensor(1 / 100)
        assert maxabs(rr * dt, p['rate_dt']) < 1e-6
        rig, seq = od.reverse(p['rigid_t'], p['seq_t'], p['rot_score'], p['trans_score'], p['logits'], p['t'], dt,
                              p['mask'], p['z_rot'], p['z_trans'], p['jumps'])
        assert rig.dtype == p['rigids_1'].dtype == torch.float64
        assert torch.equal(seq.long(), p['seq_1'].long())
        assert maxabs(rig, p['rigids_1']) < 1e-9


def test_reverse_step_corner_cases_match_reference():
    """Identity / w < 0 / angle-pi / 1e-8 rad rotations, zero perturbation, saturated logits, clamped residue indices
    and multi-jump tau-leaps (tests/golden/reverse_edges.npz, written by the reference's FullDiffuser.reverse)."""
    p = golden('reverse_edges')
    od = oracle_diffuser()
    rig, seq = od.reverse(p['rigid_t'], p['seq_t'], p['rot_score'], p['trans_score'], p['logits'], p['t'], torch.tensor(1 / 100),
                          p['mask'], p['z_rot'], p['z_trans'], p['jumps'])
    assert torch.isfinite(rig).all() and rig.dtype == torch.float64
    assert torch.equal(seq.long(), p['seq_1'].long())
    assert maxabs(rig, p['rigids_1']) < 1e-9
    assert int(seq[0, 3]) == 0 and int(seq[0, 2]) == 0 and int(seq[0, 1]) == 6     # clamp (diffused), fixed keeps 0, one jump


def test_prior_sample_matches_reference():
    g = golden('prior')
    od = oracle_diffuser()
    rig, seq = od.sample_ref(g['impute_rigids'], g['impute_seq'], g['mask'], g['z_rot'], g['u_rot'], g['z_trans'],
                             g['seq_rand'])
    assert torch.equal(seq.long(), g['seq_t'].long())
    assert maxabs(rig, g['rigids_t']) < 1e-5


def test_forward_marginal_matches_reference():
    """Optimize-mode start state (full_diffuser.py:57-126), masked and unmasked, draws replayed."""
    g = golden('marginal')
    od = oracle_diffuser()
    for tag, mask in (('m', g['mask']), ('n', None)):
        out = od.forward_marginal(g['rigids_0'], g['seq_0'], g['t'], mask, g[f'{tag}_z_rot'], g[f'{tag}_u_rot'],
                                  g[f'{tag}_z_trans'], g[f'{tag}_x_t'], g[f'{tag}_dims'], g[f'{tag}_newval'])
        assert torch.equal(out['seq_t'].long(), g[f'{tag}_seq_t'].long())
        assert maxabs(out['rigids_t'], g[f'{tag}_rigids_t']) < 1e-5
        assert maxabs(out['rot_score'], g[f'{tag}_rot_score']) < 1e-4 * max(1.0, float(g[f'{tag}_rot_score'].abs().max()))
        assert maxabs(out['trans_score'], g[f'{tag}_trans_score']) < 1e-4
        assert maxabs(out['q_t0'], g[f'{tag}_q_t0']) < 1e-6 and maxabs(out['rate_t'], g[f'{tag}_rate_t']) == 0
        assert maxabs(out['rot_score_scaling'], g[f'{tag}_rot_score_scaling']) < 1e-5
        assert maxabs(out['trans_score_scaling'], g[f'{tag}_trans_score_scaling']) < 1e-5
        # the categorical draws are possible under the rows the oracle derives (positive probability)
        c = out['checks']
        assert float(torch.gather(c['p_xt'], 2, g[f'{tag}_x_t'].long()[..., None]).min()) > 0
        assert float(torch.gather(c['p_dims'], 1, g[f'{tag}_dims'].long()[:, None]).min()) > 0
        assert float(torch.gather(c['p_new'], 1, g[f'{tag}_newval'].long()[:, None]).min()) > 0
