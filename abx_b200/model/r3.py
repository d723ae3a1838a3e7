"""Rigid-frame helpers with the reference's names (abx/model/r3.py); rigids are (rots [...,3,3], trans [...,3])."""
import torch

from abx_b200.model.quat_affine import rot_to_quat


def rots_mul_vecs(rots, vecs):
    return torch.einsum('...rd,...d->...r', rots, vecs)


def rigids_apply(rigids, points):
    """r3.py:9-16."""
    rots, trans = rigids
    if points.ndim == trans.ndim:
        return trans + rots_mul_vecs(rots, points)
    return trans[..., None, :] + torch.einsum('...rd,...md->...mr', rots, points)


rigids_mul_vecs = rigids_apply


def rots_mul_rots(a, b):
    return torch.einsum('...rd,...dm->...rm', a, b)


def rigids_mul_rigids(a, b):
    """r3.py:37-46."""
    return rots_mul_rots(a[0], b[0]), rots_mul_vecs(a[0], b[1]) + a[1]


def rigids_mul_rots(rigids, rots):
    return rots_mul_rots(rigids[0], rots), rigids[1]


def invert_rigids(rigids):
    """r3.py:54-59."""
    inv = rigids[0].transpose(-1, -2)
    return inv, -rots_mul_vecs(inv, rigids[1])


def vecs_robust_normalize(v, dim=-1, eps=1e-8):
    return v / torch.sqrt(torch.sum(v * v, dim=dim, keepdim=True) + eps)


def rigids_from_3_points(point_on_neg_x_axis, origin, point_on_xy_plane):
    """r3.py:89-112: Gram-Schmidt frame with columns (e0, e1, e2)."""
    e0 = vecs_robust_normalize(origin - point_on_neg_x_axis)
    e1 = point_on_xy_plane - origin
    e1 = vecs_robust_normalize(e1 - torch.sum(e1 * e0, dim=-1, keepdim=True) * e0)
    e2 = torch.cross(e0, e1, dim=-1)
    return torch.stack([e0, e1, e2], dim=-1), origin


def rigids_to_tensor7(rigids):
    """r3.py:114-121."""
    return torch.cat([rot_to_quat(rigids[0]), rigids[1]], dim=-1)


def rigids_op(rigids, op):
    return tuple(op(x) for x in rigids)
