"""Torsion angles -> rigid groups -> atom14 coordinates (reference: abx/model/atom.py), torch ops."""
import torch
import torch.nn.functional as F

from abx_b200.data import residue_tables as rt
from abx_b200.model import r3

_TABLES = {}


def _table(name, device, dtype=None):
    key = (name, str(device))
    if key not in _TABLES:
        t = torch.from_numpy(rt.table(name)).to(device)
        _TABLES[key] = t if dtype is None else t.to(dtype)
    return _TABLES[key]


def torsion_angles_to_frames(aatype, backb_to_global, torsion_angles_sin_cos):
    """atom.py:9-58: the 8 rigid groups of every residue in the global frame, (rots [B,N,8,3,3], trans [B,N,8,3])."""
    rots, trans = backb_to_global
    m = _table('restype_rigid_group_default_frame', aatype.device)[aatype]          # [B,N,8,4,4]
    d_rot, d_trans = m[..., :3, :3], m[..., :3, 3]
    sin = F.pad(torsion_angles_sin_cos[..., 0], (1, 0), value=0.)                     # backbone group: identity
    cos = F.pad(torsion_angles_sin_cos[..., 1], (1, 0), value=1.)
    zeros, ones = torch.zeros_like(sin), torch.ones_like(sin)
    rx = torch.stack([ones, zeros, zeros, zeros, cos, -sin, zeros, sin, cos], dim=-1).reshape(sin.shape + (3, 3))
    f_rot = r3.rots_mul_rots(d_rot, rx)
    frames = [(f_rot[:, :, g], d_trans[:, :, g]) for g in range(8)]
    chi2 = r3.rigids_mul_rigids(frames[4], frames[5])
    chi3 = r3.rigids_mul_rigids(chi2, frames[6])
    chi4 = r3.rigids_mul_rigids(chi3, frames[7])
    to_bb = frames[:5] + [chi2, chi3, chi4]
    bb_rot = torch.stack([f[0] for f in to_bb], dim=2)
    bb_trans = torch.stack([f[1] for f in to_bb], dim=2)
    return r3.rigids_mul_rigids((rots[:, :, None], trans[:, :, None]), (bb_rot, bb_trans))


def frames_and_literature_positions_to_atom14_pos(aatype, all_frames_to_global):
    """atom.py:60-76."""
    f_rot, f_trans = all_frames_to_global
    grp = _table('restype_atom14_to_rigid_group', aatype.device)[aatype].long()      # [B,N,14]
    a_rot = torch.gather(f_rot, 2, grp[..., None, None].expand(grp.shape + (3, 3)))
    a_trans = torch.gather(f_trans, 2, grp[..., None].expand(grp.shape + (3,)))
    lit = _table('restype_atom14_rigid_group_positions', aatype.device)[aatype]
    return a_trans + r3.rots_mul_vecs(a_rot, lit)
