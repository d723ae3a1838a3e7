"""Feature pipeline at the model boundary (reference: abx/model/features.py + abx/common/geometry.py).

Same registry (`FeatureBuilder` over the `[name, kwargs]` list of config_data_feature.json) and the same
output keys for the entries the sampling path reads.  Everything here runs once
per batch, before the reverse-diffusion loop, as plain torch ops on whatever device the batch lives on.
Training-only outputs of the reference pipeline (alt/ambiguous ground truth, C-alpha triplet frames,
pseudo-beta targets) are not produced.
"""
import functools

import torch
from torch.nn import functional as F

from abx_b200.data import residue_tables as rt
from abx_b200.model import r3
from abx_b200.model.head import batched_select

_feats_fn = {}
_TABLES = {}


def _table(name, device):
    key = (name, str(device))
    if key not in _TABLES:
        _TABLES[key] = torch.from_numpy(rt.table(name)).to(device)
    return _TABLES[key]


def take1st(fn):
    """Registry decorator of the reference (features.py:40-50): supply all arguments but the first."""
    @functools.wraps(fn)
    def fc(*args, **kwargs):
        return lambda x: fn(x, *args, **kwargs)
    _feats_fn[fn.__name__] = fc
    return fc


@take1st
def make_to_device(protein, fields, device, is_training=True):
    if callable(device):
        device = device()
    for k in fields:
        if k in protein:
            protein[k] = protein[k].to(device)
    return protein


@take1st
def make_restype_atom_constants(batch, is_training=False):
    """features.py:52-66."""
    dev = batch['seq'].device
    seq = batch['seq']
    batch['atom14_atom_exists'] = _table('restype_atom14_mask', dev)[seq]
    batch['atom14_atom_is_ambiguous'] = _table('restype_atom14_is_ambiguous', dev)[seq]
    batch.setdefault('residx_atom37_to_atom14', _table('restype_atom37_to_atom14', dev)[seq])
    batch.setdefault('atom37_atom_exists', _table('restype_atom37_mask', dev)[seq])
    return batch


def make_atom37_positions(batch):
    """features.py:118-128."""
    idx = batch['residx_atom37_to_atom14']
    batch['atom37_gt_positions'] = batched_select(batch['atom14_gt_positions'], idx, batch_dims=2)
    batch['atom37_gt_exists'] = torch.logical_and(batched_select(batch['atom14_gt_exists'], idx, batch_dims=2),
                                                  batch['atom37_atom_exists'])
    return batch


def atom37_to_frames(aatype, all_atom_positions, all_atom_mask):
    """abx/common/geometry.py:9-66: the 8 rigid-group frames of every residue from its atoms."""
    dev = aatype.device
    base_idx = _table('restype_rigidgroup_base_atom37_idx', dev)[aatype].long()                      # [B,N,8,3]
    flat = base_idx.reshape(base_idx.shape[:2] + (24,))
    base = batched_select(all_atom_positions, flat, batch_dims=2).reshape(base_idx.shape + (3,))      # [B,N,8,3,3]
    rots, trans = r3.rigids_from_3_points(base[..., 0, :], base[..., 1, :], base[..., 2, :])
    group_exists = _table('restype_rigidgroup_mask', dev)[aatype]
    atoms_exist = batched_select(all_atom_mask, flat, batch_dims=2).reshape(base_idx.shape)
    gt_exists = torch.logical_and(torch.all(atoms_exist, dim=-1), group_exists)
    flip = torch.eye(3, dtype=rots.dtype, device=dev).repeat(8, 1, 1)
    flip[0, 0, 0] = -1
    flip[0, 2, 2] = -1
    rots = r3.rots_mul_rots(rots, flip)
    return {'rigidgroups_gt_frames': (rots, trans), 'rigidgroups_gt_exists': gt_exists,
            'rigidgroups_group_exists': group_exists}


def atom37_to_torsion_angles(aatype, all_atom_pos, all_atom_mask):
    """abx/common/geometry.py:115-212: (pre-omega, phi, psi, chi1..4) as (sin, cos)."""
    dev = aatype.device
    prev_pos = F.pad(all_atom_pos[:, :-1], [0, 0, 0, 0, 1, 0])
    prev_mask = F.pad(all_atom_mask[:, :-1], [0, 0, 1, 0])
    pre_omega = torch.cat([prev_pos[:, :, 1:3], all_atom_pos[:, :, 0:2]], dim=-2)        # prev CA, C, this N, CA
    phi = torch.cat([prev_pos[:, :, 2:3], all_atom_pos[:, :, 0:3]], dim=-2)              # prev C, this N, CA, C
    psi = torch.cat([all_atom_pos[:, :, 0:3], all_atom_pos[:, :, 4:5]], dim=-2)          # this N, CA, C, O
    pre_omega_mask = torch.all(prev_mask[:, :, 1:3], dim=-1) & torch.all(all_atom_mask[:, :, 0:2], dim=-1)
    phi_mask = prev_mask[:, :, 2] & torch.all(all_atom_mask[:, :, 0:3], dim=-1)
    psi_mask = torch.all(all_atom_mask[:, :, 0:3], dim=-1) & all_atom_mask[:, :, 4]
    chi_idx = _table('chi_angles_atom_indices', dev)[aatype].long()                       # [B,N,4,4]
    flat = chi_idx.reshape(chi_idx.shape[:2] + (16,))
    chis = batched_select(all_atom_pos, flat, batch_dims=2).reshape(chi_idx.shape + (3,))
    chi_atoms_ok = torch.all(batched_select(all_atom_mask, flat, batch_dims=2).reshape(chi_idx.shape), dim=-1)
    chis_mask = torch.logical_and(_table('chi_angles_mask', dev)[aatype], chi_atoms_ok)
    pos = torch.cat([pre_omega[:, :, None], phi[:, :, None], psi[:, :, None], chis], dim=2)   # [B,N,7,4,3]
    mask = torch.cat([pre_omega_mask[:, :, None], phi_mask[:, :, None], psi_mask[:, :, None], chis_mask], dim=2)
    frames = r3.rigids_from_3_points(pos[..., 1, :], pos[..., 2, :], pos[..., 0, :])
    rel = r3.rigids_mul_vecs(r3.invert_rigids(frames), pos[..., 3, :])
    sc = torch.stack([rel[..., 2], rel[..., 1]], dim=-1)
    sc = sc / torch.sqrt(torch.sum(torch.square(sc), dim=-1, keepdim=True) + 1e-8)
    sc = sc * torch.tensor([1.0, 1.0, -1.0, 1.0, 1.0, 1.0, 1.0], device=dev)[..., None]
    return {'torsion_angles_sin_cos': sc, 'torsion_angles_mask': mask}


@take1st
def make_atom14_alt_gt_positions(batch, is_training=True):
    return batch            # training-loss targets only (features.py:69-79)


@take1st
def make_pseudo_beta(batch, is_training=True):
    return batch            # metric-head targets only (features.py:81-89)


@take1st
def make_calpha3_frames(batch, is_training=True):
    return batch            # training-loss targets only (features.py:101-108)


@take1st
def make_gt_frames(batch, is_training=True):
    if 'atom37_gt_positions' not in batch:
        batch = make_atom37_positions(batch)
    batch.update(atom37_to_frames(batch['seq'], batch['atom37_gt_positions'], batch['atom37_gt_exists']))
    return batch


@take1st
def make_torsion_angles(batch, is_training=True):
    if 'atom37_gt_positions' not in batch:
        batch = make_atom37_positions(batch)
    batch.update(atom37_to_torsion_angles(batch['seq'], batch['atom37_gt_positions'], batch['atom37_gt_exists']))
    return batch


def design_mask(batch, generate_area):
    """features.py:142-170 at inference: 1 on the residues to (re)design.  `generate_area` is one of H1..L3, or
    'cdr' (all six); 'cdrs' (README spelling, which the reference leaves undefined) is accepted as 'cdr'."""
    anchor_flag = batch['anchor_flag'].int()
    if generate_area in ('cdr', 'cdrs'):
        cdrs = anchor_flag[anchor_flag > 0].unique().tolist()
    elif generate_area in rt.cdr_str_to_enum:
        cdrs = [rt.cdr_str_to_enum[generate_area]]
    else:
        raise ValueError(f'unknown generate_area {generate_area!r}')
    diffused = torch.zeros_like(batch['mask'], dtype=torch.int32)
    struc = torch.zeros_like(batch['anchor_flag'], dtype=torch.int32)
    for cdr in cdrs:
        idx = torch.nonzero(anchor_flag == cdr).tolist()
        for i in range(0, len(idx) - 1, 2):
            b, right, left = idx[i][0], idx[i][1], idx[i + 1][1]
            diffused[b, right + 1: left - 1] = 1        # (sic) the last CDR residue stays fixed
            struc[b, max(right - 1, 0): min(left + 1, diffused.shape[1] - 1)] = 1
    return diffused, struc


@take1st
def make_diffuser_features(batch, generate_area, diff_conf, shrink_limit=1, extend_limit=2, is_training=False,
                           diffuser=None):
    """features.py:130-212 (inference branches): design mask, t = 1 prior draw (`sample_ref`) or, in optimize
    mode (diff_conf['opt_step']), the forward marginal at t = opt_step / inference_step."""
    assert not is_training, 'training features are not part of the sampling path'
    from abx_b200.diffuser.full_diffuser import FullDiffuser
    full_diffuser = diffuser if diffuser is not None else FullDiffuser.get(diff_conf)
    device = batch['seq'].device
    B = batch['seq'].shape[0]
    n_ab = batch['anchor_flag'].shape[1]
    gt_rots, gt_trans = batch['rigidgroups_gt_frames']
    rigids_0 = r3.rigids_to_tensor7((gt_rots[:, :, 0], gt_trans[:, :, 0]))
    seq_0 = batch['seq']
    diffused_mask, ab_struc = design_mask(batch, generate_area)
    struc_loss_mask = batch['mask'].type(torch.int32)
    struc_loss_mask[:, :n_ab] = ab_struc
    if 'opt_step' not in diff_conf:
        t = torch.ones((B,), device=device, dtype=torch.float32)
        feats = full_diffuser.sample_ref(n_samples=tuple(rigids_0.shape[:2]), impute_rigids=rigids_0, impute_seq=seq_0,
                                         diffuse_mask=diffused_mask)
    else:
        t = torch.full((B,), diff_conf['opt_step'] / diff_conf['inference_step'], device=device, dtype=torch.float32)
        feats = full_diffuser.forward_marginal(rigids_0=rigids_0, seq_0=seq_0, t=t, diffuse_mask=diffused_mask)
    batch.update(feats)
    batch.update(t=t, struc_loss_mask=struc_loss_mask, fixed_mask=1 - diffused_mask, rigids_0=rigids_0)
    return batch


class FeatureBuilder:
    """features.py:229-242."""

    def __init__(self, config, is_training=False):
        self.config = config
        self.training = is_training

    def build(self, protein):
        for fn, kwargs in (self.config or []):
            protein = _feats_fn[fn](is_training=self.training, **kwargs)(protein)
        return protein

    __call__ = build
