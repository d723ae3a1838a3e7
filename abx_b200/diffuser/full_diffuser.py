"""`FullDiffuser` — the reference's diffuser/full_diffuser.py facade over the SO(3), R^3 and categorical
diffusers, with the per-step methods (`calc_quat_score`, `calc_trans_score`, `reverse`) running as
single launches of the sm_100a kernels in libabx_b200 instead of ~120 eager ops and >= 8 host syncs.

Same constructor (`diff_conf` with keys diffuse / r3 / so3 / seq), same `get` singleton, same method
names, argument meaning and return conventions (rigids are [qw,qx,qy,qz,tx,ty,tz], translations in
Angstrom; `reverse` returns float64 rigids as the reference does once t is float64).  Two keyword-only
extensions: `noise=(z_rot, z_trans, jumps)` injects the random draws (teacher-forced parity tests),
and `generator=` draws them from a given torch generator; by default they are drawn with torch.randn /
torch.randn / torch.poisson in the reference's order and shapes.
"""
import ctypes

import numpy as np
import torch

from abx_b200 import lib
from abx_b200.diffuser import discrete_diffuser, r3_diffuser, so3_diffuser

diffuser_obj_dict = {}


def _f32(v):
    return float(np.float32(v))


def make_consts(diff_conf):
    """Schedule scalars exactly as torch evaluates them on 0-d float32 tensors (promoted against t later)."""
    so3, r3, seq = diff_conf['so3'], diff_conf['r3'], diff_conf['seq']
    e_max = torch.exp(torch.tensor(so3['max_sigma']))
    e_min = torch.exp(torch.tensor(so3['min_sigma']))
    c = lib.DiffuserConsts()
    c.so3_exp_max, c.so3_exp_min = float(e_max), float(e_min)
    c.so3_g2_coef = float(2 * (e_max - e_min))
    c.r3_min_b = float(torch.tensor(r3['min_b']))
    c.r3_delta_b = float(torch.tensor(r3['max_b'] - r3['min_b']))
    c.r3_coord_scale = float(torch.tensor(r3['coordinate_scaling']))
    c.seq_rate = float(seq['rate_const'])
    c.num_sigma, c.num_omega = int(so3['num_sigma']), int(so3['num_omega'])
    return c


def _host_scalar(x):
    """python float of a scalar argument (0-d tensor or number); float32 tensors keep their rounding."""
    if torch.is_tensor(x):
        return float(x)          # a CUDA tensor here costs a sync: pass CPU scalars on the hot path
    return float(x)


def _extract_trans_rots(rigid):
    from abx_b200.model import quat_affine as qa
    assert rigid.dim() == 3 and rigid.shape[-1] == 7
    return rigid[..., 4:], qa.quat_to_rotvec(rigid[..., :4])


def _assemble_rigid(rotvec, trans):
    from abx_b200.model import quat_affine as qa
    return torch.cat([qa.rotvec_to_quat(rotvec).to(trans.dtype), trans], dim=-1)


class FullDiffuser:

    def __init__(self, diff_conf):
        self._diff_conf = diff_conf
        self._diffuser = diff_conf['diffuse']
        self._diffuse_rot = self._diffuser['diffuse_rot']
        self._diffuse_trans = self._diffuser['diffuse_trans']
        self._diffuse_seq = self._diffuser['diffuse_seq']
        self._consts = make_consts(diff_conf)
        self._so3_diffuser = so3_diffuser.SO3Diffuser(diff_conf['so3'], consts=self._consts)
        self._r3_diffuser = r3_diffuser.R3Diffuser(diff_conf['r3'])
        self._seq_diffuser = discrete_diffuser.DiscreteDiffuser(diff_conf['seq'])
        self._lib = lib.load()

    @staticmethod
    def get(diff_conf):
        if 'diffuser' not in diffuser_obj_dict:
            diffuser_obj_dict['diffuser'] = FullDiffuser(diff_conf)
        return diffuser_obj_dict['diffuser']

    def _apply_mask(self, x_diff, x_fixed, diff_mask):
        return diff_mask * x_diff + (1 - diff_mask) * x_fixed

    @staticmethod
    def _t64(t):
        return t.to(torch.float64).contiguous(), int(t.dtype != torch.float64)

    # ---- scores of a predicted x0 (full_diffuser.py:131-142) -------------------------------------------
    def calc_scores(self, quat_t, quat_0, trans_t, trans_0, t):
        """Both scores in one launch of abx_se3_scores; either pair may be None."""
        ref = quat_t if quat_t is not None else trans_t
        B, N = ref.shape[:2]
        dev = ref.device
        t64, is32 = self._t64(t)
        tab, sig, om, _, _ = self._so3_diffuser.tables_on(dev)
        rot = trans = None
        args = [None] * 4
        if quat_t is not None:
            args[0], args[1] = quat_t.float().contiguous(), quat_0.float().reshape(B, N, 4).contiguous()
            rot = torch.empty(B, N, 3, device=dev, dtype=torch.float32)
        if trans_t is not None:
            args[2], args[3] = trans_t.float().contiguous(), trans_0.float().contiguous()
            trans = torch.empty(B, N, 3, device=dev, dtype=torch.float32 if is32 else torch.float64)
        with lib.device_guard(ref):
            lib.check(self._lib.abx_se3_scores(lib.stream(), B, N, ctypes.byref(self._consts),
                                               *(lib.ptr(a) for a in args), lib.ptr(t64), is32, lib.ptr(tab),
                                               lib.ptr(sig), lib.ptr(om), lib.ptr(rot), lib.ptr(trans)))
        return rot, trans

    def calc_quat_score(self, quat_t, quat_0, t):
        if not self._so3_diffuser.use_cached_score:
            from abx_b200.model import quat_affine as qa
            q0t = qa.quat_multiply(qa.invert_quat(quat_0.reshape(quat_t.shape)), quat_t)
            return self._so3_diffuser.score(qa.quat_to_rotvec(q0t), t)
        return self.calc_scores(quat_t, quat_0, None, None, t)[0]

    def calc_trans_score(self, trans_t, trans_0, t, scale=True):
        if not scale:
            return self._r3_diffuser.score(trans_t, trans_0, t, scale=False)
        return self.calc_scores(None, None, trans_t, trans_0, t)[1]

    def calc_trans_0(self, trans_score, trans_t, t):
        return self._r3_diffuser.calc_trans_0(trans_score, trans_t, t)

    def trans_parameters(self, trans_t, score_t, t, dt, mask):
        return self._r3_diffuser.distribution(trans_t, score_t, t, dt, mask)

    def score(self, rigid_0, rigid_t, t):
        tran_0, rot_0 = _extract_trans_rots(rigid_0)
        tran_t, rot_t = _extract_trans_rots(rigid_t)
        rot_score = self._so3_diffuser.score(rot_t, t) if self._diffuse_rot else torch.zeros_like(rot_0)
        trans_score = self._r3_diffuser.score(tran_t, tran_0, t) if self._diffuse_trans else torch.zeros_like(tran_0)
        return trans_score, rot_score

    def score_scaling(self, t):
        return self._so3_diffuser.score_scaling(t), self._r3_diffuser.score_scaling(t)

    # ---- reverse step (full_diffuser.py:174-227) ----------------------------------------------------------
    def reverse_rates(self, seq_t, logits_t, t, dt):
        """rate*dt of the categorical tau-leap (discrete_diffuser.py:150-180): what torch.poisson is fed."""
        B, N = seq_t.shape
        t64, _ = self._t64(t)
        out = torch.empty(B, N, 20, device=logits_t.device, dtype=torch.float32)
        with lib.device_guard(logits_t):
            lib.check(self._lib.abx_seq_reverse_rates(lib.stream(), B, N, ctypes.byref(self._consts),
                                                      lib.ptr(seq_t.long().contiguous()),
                                                      lib.ptr(logits_t.float().contiguous()), lib.ptr(t64),
                                                      _f32(_host_scalar(dt)), lib.ptr(out)))
        return out

    def reverse(self, rigid_t, seq_t, rot_score, trans_score, logits_t, t, dt, diffuse_mask=None, center=True,
                noise_scale=1.0, *, noise=None, generator=None):
        B, N = rigid_t.shape[:2]
        dev = rigid_t.device
        if rigid_t.dtype not in (torch.float32, torch.float64):
            rigid_t = rigid_t.float()
        rigid_t = rigid_t.contiguous()
        seq_in = seq_t.long().contiguous()
        t64, _ = self._t64(t)
        dt32 = _f32(_host_scalar(dt))
        sqrt_dt32 = float(np.sqrt(np.float32(dt32)))
        flags = (1 if self._diffuse_rot else 0) | (2 if self._diffuse_trans else 0) | \
                (4 if self._diffuse_seq else 0) | (8 if center else 0)
        rot_score = None if rot_score is None else rot_score.float().contiguous()
        trans_score = None if trans_score is None else trans_score.to(torch.float64).contiguous()
        if noise is not None:
            z_rot, z_trans, jumps = (None if z is None else z.to(dev, torch.float32).contiguous() for z in noise)
        else:   # reference draw order: randn (rotation), randn (translation), poisson (so3:351, r3:136, discrete:180)
            z_rot = torch.randn(B, N, 3, device=dev, generator=generator) if self._diffuse_rot else None
            z_trans = torch.randn(B, N, 3, device=dev, generator=generator) if self._diffuse_trans else None
            jumps = None
            if self._diffuse_seq:
                jumps = torch.poisson(self.reverse_rates(seq_in, logits_t, t, dt), generator=generator)
        mask = None if diffuse_mask is None else diffuse_mask.to(torch.int32).contiguous()
        rigids = torch.empty(B, N, 7, device=dev, dtype=torch.float64)
        seq_out = torch.empty(B, N, device=dev, dtype=torch.int64)
        with lib.device_guard(rigid_t):
            lib.check(self._lib.abx_se3_reverse_step(
                lib.stream(), B, N, ctypes.byref(self._consts), lib.ptr(rigid_t), int(rigid_t.dtype == torch.float64),
                lib.ptr(seq_in), lib.ptr(rot_score), lib.ptr(trans_score), lib.ptr(mask), lib.ptr(t64), dt32, sqrt_dt32,
                float(noise_scale), lib.ptr(z_rot), lib.ptr(z_trans), lib.ptr(jumps), flags, lib.ptr(rigids),
                lib.ptr(seq_out)))
        return rigids, seq_out

    # ---- initial states (full_diffuser.py:57-126, 229-290): once per sample, torch ops ------------------------
    def forward_marginal(self, rigids_0, seq_0, t, diffuse_mask=None):
        trans_0, rot_0 = _extract_trans_rots(rigids_0)
        if self._diffuse_rot:
            rot_t, rot_score = self._so3_diffuser.forward_marginal(rot_0, t)
            rot_score_scaling = self._so3_diffuser.score_scaling(t)
        else:
            rot_t, rot_score, rot_score_scaling = rot_0, torch.zeros_like(rot_0), torch.ones_like(t)
        if self._diffuse_trans:
            trans_t, trans_score = self._r3_diffuser.forward_marginal(trans_0, t)
            trans_score_scaling = self._r3_diffuser.score_scaling(t)
        else:
            trans_t, trans_score, trans_score_scaling = trans_0, torch.zeros_like(trans_0), torch.ones_like(t)
        if self._diffuse_seq:
            seq_t, q_t0, rate_t = self._seq_diffuser.forward_marginal(seq_0, t)
        else:
            S = self._seq_diffuser.residue_num
            seq_t = seq_0
            q_t0 = torch.eye(S, device=t.device).unsqueeze(0).expand(t.shape[0], -1, -1)
            rate_t = torch.zeros((t.shape[0], S, S), device=t.device)
        if diffuse_mask is not None:
            m = diffuse_mask[..., None]
            rot_t = self._apply_mask(rot_t, rot_0, m)
            trans_t = self._apply_mask(trans_t, trans_0, m)
            trans_score = self._apply_mask(trans_score, torch.zeros_like(trans_score), m)
            rot_score = self._apply_mask(rot_score, torch.zeros_like(rot_score), m)
            seq_t = self._apply_mask(seq_t, seq_0, diffuse_mask)
        return {'rigids_t': _assemble_rigid(rot_t, trans_t), 'trans_score': trans_score, 'rot_score': rot_score,
                'trans_score_scaling': trans_score_scaling, 'rot_score_scaling': rot_score_scaling, 'seq_t': seq_t,
                'q_t0': q_t0, 'rate_t': rate_t}

    def sample_ref(self, n_samples, impute_rigids=None, impute_seq=None, diffuse_mask=None):
        if impute_rigids is not None:
            device = impute_rigids.device
            assert tuple(impute_rigids.shape[:2]) == tuple(n_samples)
            trans_impute, rot_impute = _extract_trans_rots(impute_rigids)
            trans_impute = self._r3_diffuser._scale(trans_impute.reshape((*n_samples, 3)))
            rot_impute = rot_impute.reshape((*n_samples, 3))
        if diffuse_mask is not None and (impute_rigids is None or impute_seq is None):
            raise ValueError('Must provide imputation values.')
        if ((not self._diffuse_rot) or (not self._diffuse_trans)) and impute_rigids is None:
            raise ValueError('Must provide imputation values.')
        if (not self._diffuse_seq) and impute_seq is None:
            raise ValueError('Must provide imputation values.')
        if impute_rigids is None:
            device = 'cpu'
        # draw order of the reference: rotation (randn, rand), translation (randn), sequence (randint)
        rot_ref = self._so3_diffuser.sample_ref(n_samples=n_samples, device=device) if self._diffuse_rot else rot_impute
        trans_ref = self._r3_diffuser.sample_ref(n_samples=n_samples, device=device) if self._diffuse_trans else trans_impute
        seq_ref = self._seq_diffuser.sample_ref(n_samples=n_samples, device=device) if self._diffuse_seq else impute_seq
        if diffuse_mask is not None:
            rot_ref = self._apply_mask(rot_ref, rot_impute, diffuse_mask[..., None])
            trans_ref = self._apply_mask(trans_ref, trans_impute, diffuse_mask[..., None])
            seq_ref = self._apply_mask(seq_ref, impute_seq, diffuse_mask)
        trans_ref = self._r3_diffuser._unscale(trans_ref)
        return {'rigids_t': _assemble_rigid(rot_ref, trans_ref), 'seq_t': seq_ref}
