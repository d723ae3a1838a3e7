"""GPU parity of the SE(3)/categorical diffuser kernels (through the C-ABI, via the reference-shaped
`FullDiffuser`) against the reference's golden vectors and the CPU oracle."""
import numpy as np
import pytest
import torch

from tests.util import golden, maxabs, oracle_diffuser

pytestmark = pytest.mark.gpu


def test_igso3_tables_match_reference_rows(cuda_device):
    from tests.gpu_util import built_diffuser
    g = golden('igso3')
    so3 = built_diffuser()._so3_diffuser
    assert maxabs(so3.discrete_sigma, g['discrete_sigma']) == 0
    assert maxabs(so3.discrete_omega, g['discrete_omega']) == 0
    assert torch.isfinite(so3._score_norms).all() and torch.isfinite(so3._cdf).all()
    # rows with sigma >~ 0.5 are well conditioned in the reference's float32 series
    for k, r in enumerate(g['rows'].tolist()):
        if r < 250:
            continue
        assert maxabs(so3._pdf[r], g['pdf'][k]) < 1e-4 * float(g['pdf'][k].abs().max())
        assert maxabs(so3._cdf[r], g['cdf'][k]) < 1e-4
        assert maxabs(so3._score_norms[r], g['score_norms'][k]) < 2e-3 * float(g['score_norms'][k].abs().max())
    # low-sigma rows: the reference is float32-noisy at large omega; compare where its series is still accurate
    for k, r in enumerate(g['rows'].tolist()):
        if r >= 250:
            continue
        ok = g['pdf'][k] > 1e-3 * g['pdf'][k].max()
        assert maxabs(so3._pdf[r][ok], g['pdf'][k][ok]) < 2e-3 * float(g['pdf'][k].abs().max())
    # cdf is a proper distribution function: monotone, ends at ~1
    d = so3._cdf[:, 1:] - so3._cdf[:, :-1]
    assert float(d.min()) > -1e-6
    assert float((so3._cdf[:, -1] - 1).abs().max()) < 2e-2
    assert maxabs(so3._score_scaling[g['rows'].long()][-4:], g['score_scaling'][g['rows'].long()][-4:]) < 1e-3


def test_tables_cache_roundtrip(cuda_device):
    """A second construction on the same cache_dir loads the .npy files (reference layout) instead of rebuilding."""
    import os
    from abx_b200.diffuser.so3_diffuser import SO3Diffuser
    from tests.gpu_util import built_diffuser
    d = built_diffuser()
    conf = d._diff_conf['so3']
    sub = os.listdir(conf['cache_dir'])
    assert sub == ['eps_1000_omega_1000_min_sigma_0_1_max_sigma_1_5_schedule_logarithmic']
    assert sorted(os.listdir(os.path.join(conf['cache_dir'], sub[0]))) == ['cdf_vals.npy', 'pdf_vals.npy', 'score_norms.npy']
    again = SO3Diffuser(conf, consts=d._consts)
    assert torch.equal(again._score_norms, d._so3_diffuser._score_norms)


def test_scores_match_reference(cuda_device):
    from tests.gpu_util import reference_table_diffuser
    g = golden('scores')
    fd = reference_table_diffuser()
    qt, q0, xt, x0 = (g[k].cuda() for k in ('qt', 'q0', 'xt', 'x0'))
    for tag in 'ab':
        t64 = g[f't_{tag}'].cuda()
        t32 = t64.float()
        # entries [0, :3] hold IDENTICAL frames: their true score is 0 and both implementations return
        # float32 rounding noise of q0^-1 (x) q0 (~1e-8) amplified by score_norm / 1e-6 — compare magnitudes only
        for t_, key in ((t32, f'rot_score_{tag}'), (t64, f'rot_score_t64_{tag}')):
            out = fd.calc_quat_score(qt, q0, t_).cpu()
            ref = g[key].clone()
            noise = max(1e-4, 3 * float(ref[0, :3].abs().max()))
            assert float(out[0, :3].abs().max()) < noise
            out[0, :3] = 0
            ref[0, :3] = 0
            assert maxabs(out, ref) < 2e-5 * max(1.0, float(ref.abs().max()))
        s32 = fd.calc_trans_score(xt, x0, t32)
        assert s32.dtype == torch.float32 and maxabs(s32.cpu(), g[f'trans_score_{tag}']) < 1e-5 * float(g[f'trans_score_{tag}'].abs().max())
        s64 = fd.calc_trans_score(xt, x0, t64)
        assert s64.dtype == torch.float64 and maxabs(s64.cpu(), g[f'trans_score_t64_{tag}']) < 1e-9
        rs, ts = fd.score_scaling(t32)
        assert maxabs(ts.cpu(), g[f'trans_score_scaling_{tag}']) < 1e-5
        assert maxabs(rs.cpu(), g[f'rot_score_scaling_{tag}']) < 1e-4 * float(g[f'rot_score_scaling_{tag}'].abs().max())
        assert fd._so3_diffuser.t_to_idx(t64) == g[f'sigma_idx_{tag}'].tolist()
    grid = torch.tensor(np.linspace(0.01, 1.0, 100)).cuda()
    assert fd._so3_diffuser.t_to_idx(grid) == g['grid_sigma_idx'].tolist()


def test_scores_bucket_lookup_is_exact(cuda_device):
    """Lookup semantics (sigma index + torch.bucketize) agree with the oracle on 10^5 random rotations."""
    from tests.gpu_util import built_diffuser
    from oracle import diffusers as D
    fd = built_diffuser()
    so3 = fd._so3_diffuser
    od = D.OracleDiffuser(so3._score_norms)
    gen = torch.Generator().manual_seed(3)
    B, N = 100, 1000
    q0 = torch.randn(B, N, 4, generator=gen); q0 = q0 / q0.norm(dim=-1, keepdim=True)
    qt = torch.randn(B, N, 4, generator=gen); qt = qt / qt.norm(dim=-1, keepdim=True)
    t = torch.tensor(np.linspace(0.01, 1.0, 100))
    ref = od.calc_quat_score(qt, q0, t)
    out = fd.calc_quat_score(qt.cuda(), q0.cuda(), t.cuda()).cpu()
    bad = (out - ref).abs().max(dim=-1)[0] > 1e-4 * (1 + ref.abs().max(dim=-1)[0])
    # a 1-ulp difference in |rotvec| can move a value across a bucket edge: allow a handful
    assert int(bad.sum()) <= 5, int(bad.sum())


def test_live_series_score(cuda_device):
    """use_cached_score=False branch (so3_diffuser.py:290-295) against the oracle's series."""
    import copy
    from oracle import diffusers as D
    from tests.gpu_util import built_diffuser
    fd = copy.copy(built_diffuser())
    so3 = copy.copy(fd._so3_diffuser)
    so3.use_cached_score = False
    gen = torch.Generator().manual_seed(5)
    v = torch.randn(3, 50, 3, generator=gen) * torch.tensor([1.5, 0.7, 0.2])[:, None, None]
    t = torch.tensor([1.0, 0.6, 0.3], dtype=torch.float64)
    out = so3.score(v.cuda(), t.cuda()).cpu()
    sig = D.so3_discrete_sigma()[D.so3_t_to_idx(t, D.so3_discrete_sigma())]
    om = torch.linalg.norm(v, dim=-1) + 1e-6
    ref = torch.stack([D.igso3_score_norm(D.igso3_expansion(om[b].double(), sig[b].double()), om[b].double(), sig[b].double())
                       for b in range(3)])
    ref = ref[..., None] * v / (om[..., None] + 1e-6)
    assert maxabs(out, ref) < 1e-3 * float(ref.abs().max())


def test_reverse_step_matches_reference(cuda_device):
    from tests.gpu_util import reference_table_diffuser
    g = golden('reverse')
    fd = reference_table_diffuser()
    od = oracle_diffuser()
    dt = torch.tensor(1 / 100)
    for tag in ('t99', 't50', 't02'):
        p = {k[len(tag) + 1:]: v for k, v in g.items() if k.startswith(tag + '_')}
        c = {k: v.cuda() for k, v in p.items()}
        rates = fd.reverse_rates(c['seq_t'], c['logits'], c['t'], dt)
        # closed-form transition matrix vs the reference's float32 eigendecomposition: ~1e-5 relative
        assert maxabs(rates.cpu(), p['rate_dt']) < 1e-7 + 5e-5 * float(p['rate_dt'].abs().max())
        rig, seq = fd.reverse(c['rigid_t'], c['seq_t'], c['rot_score'], c['trans_score'], c['logits'], c['t'], dt,
                              diffuse_mask=c['mask'], noise=(c['z_rot'], c['z_trans'], c['jumps']))
        assert rig.dtype == torch.float64 and seq.dtype == torch.int64
        assert torch.equal(seq.cpu(), p['seq_1'].long())                      # bit-exact residue types
        # float32 input state (first step): the reference evaluates quat->rotvec in float32
        tol = 1e-9 if p['rigid_t'].dtype == torch.float64 else 2e-5
        assert maxabs(rig.cpu(), p['rigids_1']) < tol
        # and against the oracle with no mask / no centring
        rig_o, seq_o = od.reverse(p['rigid_t'], p['seq_t'], p['rot_score'], p['trans_score'], p['logits'], p['t'], dt,
                                  torch.ones_like(p['mask']), p['z_rot'], p['z_trans'], p['jumps'], center=False)
        rig_c, seq_c = fd.reverse(c['rigid_t'], c['seq_t'], c['rot_score'], c['trans_score'], c['logits'], c['t'], dt,
                                  diffuse_mask=None, center=False, noise=(c['z_rot'], c['z_trans'], c['jumps']))
        assert torch.equal(seq_c.cpu(), seq_o.long()) and maxabs(rig_c.cpu(), rig_o) < tol


def test_reverse_step_corner_cases(cuda_device):
    """The reverse-step kernel on the corner cases of tests/golden/reverse_edges.npz (identity / w < 0 / angle-pi /
    1e-8 rad rotations, zero perturbation, saturated logits, clamped indices, multi-jump tau-leaps): residue types
    bit-exact, frames within 1e-9 of the reference's float64 result."""
    from tests.gpu_util import reference_table_diffuser
    p = golden('reverse_edges')
    fd = reference_table_diffuser()
    c = {k: v.cuda() for k, v in p.items()}
    rig, seq = fd.reverse(c['rigid_t'], c['seq_t'], c['rot_score'], c['trans_score'], c['logits'], c['t'], torch.tensor(1 / 100),
                          diffuse_mask=c['mask'], noise=(c['z_rot'], c['z_trans'], c['jumps']))
    assert rig.dtype == torch.float64 and bool(torch.isfinite(rig).all())
    assert torch.equal(seq.cpu(), p['seq_1'].long())
    assert maxabs(rig.cpu(), p['rigids_1']) < 1e-9


def test_reverse_draws_noise_in_reference_order(cuda_device):
    """Default path draws randn, randn, poisson (same shapes/order as the reference) from the CUDA generator."""
    from tests.gpu_util import reference_table_diffuser
    g = golden('reverse')
    fd = reference_table_diffuser()
    p = {k[4:]: v.cuda() for k, v in g.items() if k.startswith('t50_')}
    dt = torch.tensor(1 / 100)
    torch.manual_seed(123)
    a = fd.reverse(p['rigid_t'], p['seq_t'], p['rot_score'], p['trans_score'], p['logits'], p['t'], dt, diffuse_mask=p['mask'])
    torch.manual_seed(123)
    z_rot = torch.randn(2, 48, 3, device='cuda')
    z_trans = torch.randn(2, 48, 3, device='cuda')
    jumps = torch.poisson(fd.reverse_rates(p['seq_t'], p['logits'], p['t'], dt))
    b = fd.reverse(p['rigid_t'], p['seq_t'], p['rot_score'], p['trans_score'], p['logits'], p['t'], dt,
                   diffuse_mask=p['mask'], noise=(z_rot, z_trans, jumps))
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_reverse_step_large_batch_properties(cuda_device):
    """Full-size (N=350, B=64) step: centred translations, unit quaternions, fixed residues untouched."""
    from tests.gpu_util import built_diffuser
    fd = built_diffuser()
    gen = torch.Generator(device='cuda').manual_seed(9)
    B, N = 64, 350
    q = torch.randn(B, N, 4, device='cuda', generator=gen, dtype=torch.float64); q = q / q.norm(dim=-1, keepdim=True)
    x = torch.randn(B, N, 3, device='cuda', generator=gen, dtype=torch.float64) * 15
    rig = torch.cat([q, x], -1)
    seq = torch.randint(0, 20, (B, N), device='cuda', generator=gen)
    mask = (torch.rand(B, N, device='cuda', generator=gen) < 0.1).int()
    t = torch.full((B,), 0.37, dtype=torch.float64, device='cuda')
    out, seq1 = fd.reverse(rig, seq, torch.randn(B, N, 3, device='cuda', generator=gen),
                           torch.randn(B, N, 3, device='cuda', generator=gen, dtype=torch.float64),
                           torch.randn(B, N, 20, device='cuda', generator=gen), t, torch.tensor(0.01), diffuse_mask=mask,
                           generator=gen)
    fixed = mask == 0
    assert float((out[..., :4].norm(dim=-1) - 1).abs().max()) < 1e-12
    assert maxabs(out[..., 4:][fixed].cpu(), x[fixed].cpu()) == 0
    assert torch.equal(seq1[fixed], seq[fixed])
    same_rot = (out[..., :4] - q).abs().max(-1)[0].minimum((out[..., :4] + q).abs().max(-1)[0])
    assert float(same_rot[fixed].max()) < 1e-12
    assert int(seq1.min()) >= 0 and int(seq1.max()) <= 19


def test_prior_sample_matches_oracle(cuda_device):
    """sample_ref (full_diffuser.py:229-290) with the CUDA generator's draws replayed through the oracle."""
    from tests.gpu_util import reference_table_diffuser
    g = golden('prior')
    fd = reference_table_diffuser()
    od = oracle_diffuser()
    rig, seq, mask = g['impute_rigids'].cuda(), g['impute_seq'].cuda(), g['mask'].cuda()
    torch.manual_seed(7)
    out = fd.sample_ref(n_samples=tuple(rig.shape[:2]), impute_rigids=rig, impute_seq=seq, diffuse_mask=mask)
    torch.manual_seed(7)
    z_rot = torch.randn(2, 40, 3, device='cuda'); u = torch.rand(2, 40, device='cuda')
    z_tr = torch.randn(2, 40, 3, device='cuda'); sr = torch.randint(0, 20, (2, 40), device='cuda')
    r_o, s_o = od.sample_ref(g['impute_rigids'], g['impute_seq'], g['mask'], z_rot.cpu(), u.cpu(), z_tr.cpu(), sr.cpu())
    assert torch.equal(out['seq_t'].cpu().long(), s_o.long())
    assert maxabs(out['rigids_t'].cpu(), r_o) < 1e-4


def test_forward_marginal_matches_oracle(cuda_device):
    """forward_marginal (full_diffuser.py:57-126, optimize-mode start state) on the device; the draws the product
    makes (randn, rand, normal, 3 x multinomial — the reference's order) are captured and replayed through the oracle."""
    from tests.gpu_util import reference_table_diffuser
    g = golden('marginal')
    fd = reference_table_diffuser()
    od = oracle_diffuser()
    rig, seq, t = g['rigids_0'].cuda(), g['seq_0'].cuda(), g['t'].cuda()
    B, N = seq.shape
    for mask in (g['mask'], None):
        log = []
        orig = {n: getattr(torch, n) for n in ('randn', 'rand', 'normal', 'multinomial')}

        def wrap(n):
            def f(*a, **k):
                out = orig[n](*a, **k)
                if n == 'normal':
                    log.append((n, ((out.double() - k['mean'].double()) / k['std'].double()).float().cpu()))
                else:
                    log.append((n, out.clone().cpu()))
                return out
            return f
        for n in orig:
            setattr(torch, n, wrap(n))
        try:
            torch.manual_seed(3)
            out = fd.forward_marginal(rig, seq, t, diffuse_mask=None if mask is None else mask.cuda())
        finally:
            for n, f in orig.items():
                setattr(torch, n, f)
        assert [k for k, _ in log] == ['randn', 'rand', 'normal', 'multinomial', 'multinomial', 'multinomial']
        ref = od.forward_marginal(g['rigids_0'], g['seq_0'], g['t'], mask, log[0][1], log[1][1], log[2][1],
                                  log[3][1].reshape(B, N), log[4][1].reshape(B), log[5][1].reshape(B))
        assert torch.equal(out['seq_t'].cpu().long(), ref['seq_t'].long())
        assert maxabs(out['rigids_t'].cpu(), ref['rigids_t']) < 1e-4
        assert maxabs(out['rot_score'].cpu(), ref['rot_score']) < 1e-4 * max(1.0, float(ref['rot_score'].abs().max()))
        assert maxabs(out['trans_score'].cpu(), ref['trans_score']) < 1e-4
        assert maxabs(out['q_t0'].cpu(), ref['q_t0']) < 1e-6 and maxabs(out['rate_t'].cpu(), ref['rate_t']) < 1e-7
        assert maxabs(out['trans_score_scaling'].cpu(), ref['trans_score_scaling']) < 1e-5
        assert maxabs(out['rot_score_scaling'].cpu(), ref['rot_score_scaling']) < 1e-3
        # draws are admissible: every categorical draw has positive probability under the oracle's rows
        c = ref['checks']
        assert float(torch.gather(c['p_xt'], 2, log[3][1].reshape(B, N, 1).long()).min()) > 0
        assert float(torch.gather(c['p_dims'], 1, log[4][1].reshape(B, 1).long()).min()) > 0
        assert float(torch.gather(c['p_new'], 1, log[5][1].reshape(B, 1).long()).min()) > 0
        if mask is not None:                                       # fixed residues keep x0 exactly (translation, type)
            fixed = mask == 0
            assert maxabs(out['rigids_t'][..., 4:].cpu()[fixed], g['rigids_0'][..., 4:][fixed]) == 0
            assert torch.equal(out['seq_t'].cpu().long()[fixed], g['seq_0'].long()[fixed])


def test_no_cpu_fallback(cuda_device):
    from abx_b200.lib import AbxError
    from tests.gpu_util import built_diffuser
    fd = built_diffuser()
    g = golden('scores')
    with pytest.raises(AbxError):
        fd.calc_quat_score(g['qt'], g['q0'], g['t_a'])            # CPU tensors


@pytest.mark.parametrize('center', [True, False])
def test_reverse_step_baseline_size_matches_oracle(cuda_device, center):
    """FullDiffuser.reverse at the benchmark's size (B=8, N=350, the centre-of-mass block reduction over 350 residues)
    against the CPU oracle with injected noise: residue types bit-exact, frames <= 1e-9 (float64 on both sides)."""
    from tests.gpu_util import reference_table_diffuser
    fd = reference_table_diffuser()
    od = oracle_diffuser()
    gen = torch.Generator().manual_seed(31)
    B, N = 8, 350
    q = torch.randn(B, N, 4, generator=gen, dtype=torch.float64); q = q / q.norm(dim=-1, keepdim=True)
    rig = torch.cat([q, torch.randn(B, N, 3, generator=gen, dtype=torch.float64) * 15], -1)
    seq = torch.randint(0, 20, (B, N), generator=gen)
    mask = (torch.rand(B, N, generator=gen) < 0.1).int()
    mask[0] = 0; mask[0, 100:112] = 1                                           # an H3-like window
    mask[1] = 1                                                                 # everything diffused
    t = torch.full((B,), 0.37, dtype=torch.float64)
    dt = torch.tensor(1 / 100)
    rot_score, trans_score = torch.randn(B, N, 3, generator=gen), torch.randn(B, N, 3, generator=gen, dtype=torch.float64)
    logits = torch.randn(B, N, 20, generator=gen) * 3
    z_rot, z_trans = torch.randn(B, N, 3, generator=gen), torch.randn(B, N, 3, generator=gen)
    jumps = torch.poisson(torch.rand(B, N, 20, generator=gen) * 0.2, generator=gen)
    ref_rig, ref_seq = od.reverse(rig, seq, rot_score, trans_score, logits, t, dt, mask, z_rot, z_trans, jumps, center=center)
    c = lambda x: x.cuda()
    out, seq1 = fd.reverse(c(rig), c(seq), c(rot_score), c(trans_score), c(logits), c(t), dt, diffuse_mask=c(mask), center=center,
                           noise=(c(z_rot), c(z_trans), c(jumps)))
    assert out.dtype == torch.float64
    assert torch.equal(seq1.cpu(), ref_seq.long())
    assert maxabs(out.cpu(), ref_rig) < 1e-9
