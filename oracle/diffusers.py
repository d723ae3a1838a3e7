"""ORACLE (test infrastructure, not product code) — CPU restatement of the reference's SE(3) +
categorical diffusers: IGSO(3) tables and score lookup, R^3 VP-SDE, uniform-rate CTMC, and the
`FullDiffuser` reverse step.  Pinned by tests/golden/{igso3,scores,reverse,prior}.npz, which were
produced by the reference's own modules (oracle/make_golden.py).

Randomness is never drawn here: every function takes its noise as arguments, in the order the
reference draws it, so results are device- and generator-independent.
"""
import math

import numpy as np
import torch

from oracle import quat as Q

DEFAULT_CONF = {
    'inference_step': 100,
    'diffuse': {'diffuse_trans': True, 'diffuse_rot': True, 'diffuse_seq': True},
    'r3': {'min_b': 0.1, 'max_b': 20.0, 'coordinate_scaling': 0.1},
    'so3': {'num_omega': 1000, 'num_sigma': 1000, 'min_sigma': 0.1, 'max_sigma': 1.5,
            'schedule': 'logarithmic', 'cache_dir': '.cache/', 'use_cached_score': True},
    'seq': {'rate_const': 0.3},
}


# ---------------------------------------------------------------------------------------------------
# IGSO(3)  (diffuser/so3_diffuser.py)
# ---------------------------------------------------------------------------------------------------

def so3_sigma(t, min_sigma=0.1, max_sigma=1.5):
    """so3_diffuser.py:198-205 (logarithmic schedule); dtype follows t."""
    return torch.log(t * torch.exp(torch.tensor(max_sigma)) + (1 - t) * torch.exp(torch.tensor(min_sigma)))


def so3_discrete_sigma(num_sigma=1000, min_sigma=0.1, max_sigma=1.5):
    """so3_diffuser.py:183-187."""
    return so3_sigma(torch.linspace(0.0, 1.0, num_sigma), min_sigma, max_sigma)


def so3_discrete_omega(num_omega=1000):
    """so3_diffuser.py:128."""
    return torch.linspace(0, np.pi, num_omega + 1)[1:]


def so3_sigma_idx(sigma, discrete_sigma):
    """so3_diffuser.py:189-196: #{k: grid[k] <= sigma + 1e-5} - 1."""
    return torch.sum(discrete_sigma[None, ...] <= sigma[..., None] + 1e-5, -1) - 1


def so3_t_to_idx(t, discrete_sigma, min_sigma=0.1, max_sigma=1.5):
    """so3_diffuser.py:218-220."""
    return so3_sigma_idx(so3_sigma(t, min_sigma, max_sigma), discrete_sigma)


def igso3_expansion(omega, eps, L=1000):
    """so3_diffuser.py:15-49 (1-D omega): Σ_l (2l+1) e^{-l(l+1)ε²/2} sin((l+½)ω)/sin(ω/2)."""
    ls = torch.arange(L)[None]
    om = omega[..., None]
    p = (2 * ls + 1) * torch.exp(-ls * (ls + 1) * eps ** 2 / 2) * torch.sin(om * (ls + 1 / 2)) / torch.sin(om / 2)
    return p.sum(dim=-1)


def igso3_score_norm(exp, omega, eps, L=1000):
    """so3_diffuser.py:72-112: d/dω log(expansion) via the quotient rule, denominator exp + 1e-4."""
    ls = torch.arange(L)[None]
    om = omega[..., None]
    hi = torch.sin(om * (ls + 1 / 2))
    dhi = (ls + 1 / 2) * torch.cos(om * (ls + 1 / 2))
    lo = torch.sin(om / 2)
    dlo = 1 / 2 * torch.cos(om / 2)
    d = (2 * ls + 1) * torch.exp(-ls * (ls + 1) * eps ** 2 / 2) * (lo * dhi - hi * dlo) / lo ** 2
    return d.sum(dim=-1) / (exp + 1e-4)


def igso3_table_rows(sigma_rows, num_omega=1000, L=1000):
    """Rows of the three cached tables (so3_diffuser.py:150-166) for the given σ values.
    Returns exp_vals, pdf, cdf, score_norms, each [len(sigma_rows), num_omega] float32."""
    omega = so3_discrete_omega(num_omega)
    exps, pdfs, cdfs, scs = [], [], [], []
    for s in sigma_rows:
        e = igso3_expansion(omega, s, L)
        pdf = e * (1 - torch.cos(omega)) / torch.tensor(np.pi)          # density(), :52-69
        cdf = torch.cumsum(pdf, dim=0) / num_omega * torch.tensor(np.pi)
        exps.append(e); pdfs.append(pdf); cdfs.append(cdf)
        scs.append(igso3_score_norm(e, omega, s, L))
    return torch.stack(exps), torch.stack(pdfs), torch.stack(cdfs), torch.stack(scs)


def so3_score_scaling(score_norms, pdf):
    """so3_diffuser.py:176-181."""
    return torch.sqrt(torch.abs(torch.sum(score_norms ** 2 * pdf, dim=-1) / torch.sum(pdf, dim=-1))) / math.sqrt(3)


def so3_diffusion_coef(t, min_sigma=0.1, max_sigma=1.5):
    """so3_diffuser.py:207-216: g(t) = sqrt(2 (e^{σmax} - e^{σmin}) σ(t) / e^{σ(t)})."""
    s = so3_sigma(t, min_sigma, max_sigma)
    return torch.sqrt(2 * (torch.exp(torch.tensor(max_sigma)) - torch.exp(torch.tensor(min_sigma))) * s / torch.exp(s))


def so3_score_cached(vec, t, score_norms, discrete_sigma, discrete_omega, eps=1e-6):
    """so3_diffuser.py:264-297, cached branch (inference.py:99 forces use_cached_score=True).
    vec [B,N,3], t [B]; score_norms is the full [num_sigma, num_omega] table."""
    omega = torch.linalg.norm(vec, dim=-1) + eps
    rows = score_norms[so3_t_to_idx(t, discrete_sigma)]                      # [B, num_omega]
    idx = torch.bucketize(omega, discrete_omega[:-1])
    val = torch.gather(rows, 1, idx)
    return val[..., None] * vec / (omega[..., None] + eps)


def so3_reverse(rot_t, score_t, t, dt, z, noise_scale=1.0):
    """so3_diffuser.py:328-361 with the randn draw `z` injected."""
    g = so3_diffusion_coef(t)[:, None, None]
    z = noise_scale * z
    perturb = (g ** 2) * score_t * dt + g * torch.sqrt(dt) * z
    q1 = Q.quat_multiply(Q.rotvec_to_quat(rot_t), Q.rotvec_to_quat(perturb))
    return Q.quat_to_rotvec(q1)


def torch_interp(x_new, x, y):
    """abx/utils.py:31-59 batched 1-D linear interpolation."""
    order = x.argsort(dim=1)
    x = torch.gather(x, -1, order)
    y = torch.gather(y, -1, order)
    b = torch.sum(x.unsqueeze(2) < x_new.unsqueeze(1), dim=1)
    b = torch.clamp(b, 0, x.shape[1] - 2)
    x_lo, x_hi = torch.gather(x, -1, b), torch.gather(x, -1, b + 1)
    y_lo, y_hi = torch.gather(y, -1, b), torch.gather(y, -1, b + 1)
    w = (x_new - x_lo) / (x_hi - x_lo + 1e-8)
    w = torch.where(x_new > x[:, -1].unsqueeze(1), torch.ones_like(w), w)
    w = torch.where(x_new < x[:, 0].unsqueeze(1), torch.zeros_like(w), w)
    return y_lo * (1 - w) + y_hi * w


def so3_sample(t, z, u, cdf, discrete_sigma, discrete_omega):
    """so3_diffuser.py:222-258: axis = z/|z| (randn [B,N,3]), angle = inverse-CDF(u) (rand [B,N])."""
    axis = z / torch.linalg.norm(z, dim=-1, keepdim=True)
    rows = cdf[so3_t_to_idx(t, discrete_sigma)]
    om = discrete_omega[None].expand(t.shape[0], -1)
    return axis * torch_interp(u, rows, om)[..., None]


# ---------------------------------------------------------------------------------------------------
# R^3 VP-SDE  (diffuser/r3_diffuser.py)
# ---------------------------------------------------------------------------------------------------

def r3_b_t(t, min_b=0.1, max_b=20.0):
    """r3_diffuser.py:29-32."""
    return torch.tensor(min_b) + t * torch.tensor(max_b - min_b)


def r3_marginal_b_t(t, min_b=0.1, max_b=20.0):
    """r3_diffuser.py:45-46."""
    return t * torch.tensor(min_b) + (1 / 2) * (t ** 2) * torch.tensor(max_b - min_b)


def r3_score(x_t, x_0, t, scale=True, coordinate_scaling=0.1):
    """r3_diffuser.py:158-164 (FullDiffuser.calc_trans_score passes scale=True, full_diffuser.py:131-133)."""
    if scale:
        x_t = x_t * torch.tensor(coordinate_scaling)
        x_0 = x_0 * torch.tensor(coordinate_scaling)
    t = t[:, None, None]
    return -(x_t - torch.exp(-1 / 2 * r3_marginal_b_t(t)) * x_0) / (1 - torch.exp(-r3_marginal_b_t(t)))


def r3_reverse(x_t, score_t, t, dt, z, center=True, noise_scale=1.0, coordinate_scaling=0.1):
    """r3_diffuser.py:110-148 with the randn draw `z` injected.  Note the `g·dt·z` noise term (:137)
    and the centre of mass over ALL residues (mask=None -> ones, :141-146)."""
    x = x_t * torch.tensor(coordinate_scaling)
    b = r3_b_t(t)
    g = torch.sqrt(b)[:, None, None]
    f = -1 / 2 * b[:, None, None] * x
    z = noise_scale * z
    perturb = (f - g ** 2 * score_t) * dt + g * dt * z
    mask = torch.ones(x.shape[:-1])
    x1 = x - perturb
    if center:
        com = torch.sum(x1, dim=-2) / torch.sum(mask, dim=-1, keepdim=True)
        x1 = x1 - com[..., None, :]
    return x1 / torch.tensor(coordinate_scaling)


# ---------------------------------------------------------------------------------------------------
# categorical CTMC  (diffuser/discrete_diffuser.py)
# ---------------------------------------------------------------------------------------------------

def seq_rate_matrix(rate_const=0.3, S=20):
    """discrete_diffuser.py:15-26: off-diagonal rate_const, rows sum to zero; eigh in float32."""
    rate = rate_const * torch.ones((S, S))
    rate = rate - torch.diag(torch.diag(rate))
    rate = rate - torch.diag(torch.sum(rate, dim=1))
    eigvals, eigvecs = torch.linalg.eigh(rate)
    return rate.float(), eigvals.float(), eigvecs.float()


def seq_transition(t, eigvals, eigvecs):
    """discrete_diffuser.py:53-67: Q_t = V diag(e^{λt}) Vᵀ, entries < 1e-8 zeroed."""
    S = eigvals.shape[0]
    t = t.float()
    tr = eigvecs.reshape(1, S, S) @ torch.diag_embed(torch.exp(eigvals.reshape(1, S) * t.reshape(-1, 1))) \
        @ eigvecs.T.reshape(1, S, S)
    tr = torch.where(tr < 1e-8, torch.zeros_like(tr), tr)
    return tr


def seq_reverse_rates(x_t, logits_t, t, rate, eigvals, eigvecs, eps_ratio=1e-9):
    """discrete_diffuser.py:150-179: reverse rates [B,N,S] of the τ-leap (before × dt)."""
    B, N = x_t.shape
    S = rate.shape[0]
    x = torch.clamp(x_t, min=0, max=S - 1).long()
    p0t = torch.softmax(logits_t, dim=2)
    qt0 = seq_transition(t * torch.ones((B,)), eigvals, eigvecs)             # [B, S, S]
    bidx = torch.arange(B)[:, None, None]
    sidx = torch.arange(S)[None, None, :]
    denom = qt0[bidx, sidx, x[..., None]] + torch.tensor(eps_ratio)          # qt0[b, s, x_n]
    fwd = rate[None].expand(B, S, S)[bidx, sidx, x[..., None]]               # rate[s, x_n]
    inner = (p0t / denom) @ qt0
    rr = fwd * inner
    rr = rr.scatter(2, x[..., None], 0.0)
    return rr, x


def seq_apply_jumps(x, jumps):
    """discrete_diffuser.py:181-188: x + Σ_s k_s (s - x), clamp [0, S-1], int32."""
    S = jumps.shape[-1]
    diffs = torch.arange(S).view(1, 1, S) - x.view(*x.shape, 1)
    xp = x + torch.sum(jumps * diffs, dim=2)
    return torch.clamp(xp, min=0, max=S - 1).to(dtype=torch.int32)


# ---------------------------------------------------------------------------------------------------
# FullDiffuser  (diffuser/full_diffuser.py)
# ---------------------------------------------------------------------------------------------------

class OracleDiffuser:
    """Holds the cached IGSO(3) tables + CTMC eigendecomposition and exposes the reference
    `FullDiffuser` methods used on the sampling path, with noise injected instead of drawn."""

    def __init__(self, score_norms, cdf=None, pdf=None, conf=None):
        conf = conf or DEFAULT_CONF
        so3 = conf['so3']
        self.conf = conf
        self.discrete_sigma = so3_discrete_sigma(so3['num_sigma'], so3['min_sigma'], so3['max_sigma'])
        self.discrete_omega = so3_discrete_omega(so3['num_omega'])
        self.score_norms = torch.as_tensor(score_norms)
        self.cdf = None if cdf is None else torch.as_tensor(cdf)
        self.pdf = None if pdf is None else torch.as_tensor(pdf)
        self.rate, self.eigvals, self.eigvecs = seq_rate_matrix(conf['seq']['rate_const'])

    # full_diffuser.py:135-142
    def calc_quat_score(self, quat_t, quat_0, t):
        q0t = Q.quat_multiply(Q.invert_quat(quat_0), quat_t)
        return so3_score_cached(Q.quat_to_rotvec(q0t), t, self.score_norms, self.discrete_sigma, self.discrete_omega)

    # full_diffuser.py:131-133
    def calc_trans_score(self, trans_t, trans_0, t, scale=True):
        return r3_score(trans_t, trans_0, t, scale=scale)

    # full_diffuser.py:169-172 (so3_diffuser.py:299-301, r3_diffuser.py:107-108,150-156)
    def score_scaling(self, t):
        assert self.pdf is not None
        rot = so3_score_scaling(self.score_norms, self.pdf)[so3_t_to_idx(t, self.discrete_sigma)]
        trans = 1 / torch.sqrt(1 - torch.exp(-r3_marginal_b_t(t)))
        return rot, trans

    def reverse_rates(self, seq_t, logits_t, t):
        return seq_reverse_rates(seq_t, logits_t, t, self.rate, self.eigvals, self.eigvecs)

    def reverse(self, rigid_t, seq_t, rot_score, trans_score, logits_t, t, dt, diffuse_mask,
                z_rot, z_trans, jumps, center=True, noise_scale=1.0):
        """full_diffuser.py:174-227.  `jumps` = the Poisson(rate·dt) draw [B,N,20] (or a callable
        rate_dt -> counts, used when generating goldens)."""
        rot_t = Q.quat_to_rotvec(rigid_t[..., :4])                          # _extract_trans_rots :12-18
        trans_t = rigid_t[..., 4:]
        rot_1 = so3_reverse(rot_t, rot_score, t, dt, z_rot, noise_scale)
        trans_1 = r3_reverse(trans_t, trans_score, t, dt, z_trans, center, noise_scale)
        rr, x = self.reverse_rates(seq_t, logits_t, t)
        if callable(jumps):
            jumps = jumps(rr * dt)
        seq_1 = seq_apply_jumps(x, jumps)
        m = diffuse_mask
        trans_1 = m[..., None] * trans_1 + (1 - m[..., None]) * trans_t      # _apply_mask :54-55
        rot_1 = m[..., None] * rot_1 + (1 - m[..., None]) * rot_t
        seq_1 = m * seq_1 + (1 - m) * seq_t
        rigids = torch.cat([Q.rotvec_to_quat(rot_1).to(trans_1.dtype), trans_1], dim=-1)   # _assemble_rigid :20-26
        return rigids, seq_1

    def sample_ref(self, impute_rigids, impute_seq, diffuse_mask, z_rot, u_rot, z_trans, seq_rand):
        """full_diffuser.py:229-290 with the four draws injected in reference order: randn[B,N,3]
        (rotation axis), rand[B,N] (angle quantile), randn[B,N,3] (translation), randint[B,N]."""
        assert self.cdf is not None
        B = impute_rigids.shape[0]
        rot_imp = Q.quat_to_rotvec(impute_rigids[..., :4])
        trans_imp = impute_rigids[..., 4:] * torch.tensor(self.conf['r3']['coordinate_scaling'])
        rot_ref = so3_sample(torch.ones(B), z_rot, u_rot, self.cdf, self.discrete_sigma, self.discrete_omega)
        m = diffuse_mask
        rot_ref = m[..., None] * rot_ref + (1 - m[..., None]) * rot_imp
        trans_ref = m[..., None] * z_trans + (1 - m[..., None]) * trans_imp
        seq_ref = m * seq_rand + (1 - m) * impute_seq
        trans_ref = trans_ref / torch.tensor(self.conf['r3']['coordinate_scaling'])
        return torch.cat([Q.rotvec_to_quat(rot_ref), trans_ref], dim=-1), seq_ref

    def forward_marginal(self, rigids_0, seq_0, t, diffuse_mask, z_rot, u_rot, z_trans, x_t, dims, newval):
        """full_diffuser.py:57-126 (optimize-mode start state) with the draws injected in reference order:
        randn [B,N,3] / rand [B,N] (SO3Diffuser.sample, so3_diffuser.py:222-258), the unit normal behind
        torch.normal(mean, std) (r3_diffuser.py:101), and the three categorical draws of
        discrete_diffuser.py:72-127 (x_t [B,N], the perturbed position [B], its new value [B]).
        Returns the reference's dict; `checks` holds the probability rows the categorical draws came from."""
        assert self.cdf is not None and self.pdf is not None
        B, N = seq_0.shape
        cs = torch.tensor(self.conf['r3']['coordinate_scaling'])
        rot_0 = Q.quat_to_rotvec(rigids_0[..., :4])                        # _extract_trans_rots :12-18
        trans_0 = rigids_0[..., 4:]
        # so3_diffuser.py:303-326
        sampled = so3_sample(t, z_rot, u_rot, self.cdf, self.discrete_sigma, self.discrete_omega)
        rot_score = so3_score_cached(sampled, t, self.score_norms, self.discrete_sigma, self.discrete_omega)
        rot_t = Q.quat_to_rotvec(Q.quat_multiply(Q.rotvec_to_quat(rot_0), Q.rotvec_to_quat(sampled)))
        # r3_diffuser.py:80-105
        x0 = trans_0 * cs
        lmc = (-0.5 * r3_marginal_b_t(t)).view(B, 1, 1)
        mean, std = torch.exp(lmc) * x0, torch.sqrt(1.0 - torch.exp(2.0 * lmc))
        xt = mean + std * z_trans
        trans_score = r3_score(xt, x0, t, scale=False)
        trans_t = xt / cs
        # discrete_diffuser.py:72-127
        S = self.rate.shape[0]
        qt0 = seq_transition(t, self.eigvals, self.eigvecs)
        rate = self.rate[None].expand(B, S, S)
        x0c = torch.clamp(seq_0, min=0, max=S - 1).long()
        bidx = torch.arange(B)[:, None].expand(B, N)
        p_xt = qt0[bidx, x0c]                                               # rows the x_t draw came from [B,N,S]
        rv = rate[bidx, x_t.long()].clone()
        rv.scatter_(2, x_t.long()[..., None], 0.0)
        p_dims = rv.sum(dim=2)                                              # [B,N] (unnormalised)
        p_new = rv[torch.arange(B), dims.long()]                            # [B,S]
        seq_t = x_t.clone()
        seq_t[torch.arange(B), dims.long()] = newval.to(seq_t.dtype)
        rot_scaling, trans_scaling = self.score_scaling(t)
        if diffuse_mask is not None:
            m = diffuse_mask
            rot_t = m[..., None] * rot_t + (1 - m[..., None]) * rot_0
            trans_t = m[..., None] * trans_t + (1 - m[..., None]) * trans_0
            trans_score = m[..., None] * trans_score
            rot_score = m[..., None] * rot_score
            seq_t = m * seq_t + (1 - m) * seq_0
        return {'rigids_t': torch.cat([Q.rotvec_to_quat(rot_t), trans_t], dim=-1), 'trans_score': trans_score,
                'rot_score': rot_score, 'trans_score_scaling': trans_scaling, 'rot_score_scaling': rot_scaling,
                'seq_t': seq_t, 'q_t0': qt0, 'rate_t': rate,
                'checks': {'p_xt': p_xt, 'p_dims': p_dims, 'p_new': p_new}}
