"""Quaternion helpers with the reference's names (abx/model/quat_affine.py), as component-wise torch ops.

Used off the per-step path only (prior / forward-marginal draws, API helpers); inside the sampler loop
the same algebra runs as device functions of the C-ABI kernels (abx_b200/csrc/common.cuh).
Quaternions are real-first [w, x, y, z].
"""
import torch


def l2_normalize(v, dim=-1, eps=1e-12):
    """abx/model/utils.py:12-14."""
    return v / torch.sqrt(torch.sum(v ** 2, dim=dim, keepdim=True) + eps)


def make_identity(out_shape, device):
    """quat_affine.py:53-58: (identity quaternions, zero translations)."""
    quats = torch.zeros(tuple(out_shape) + (4,), device=device)
    quats[..., 0] = 1.0
    return quats, torch.zeros(tuple(out_shape) + (3,), device=device)


def quat_to_rot(q):
    """quat_affine.py:60-67."""
    w, x, y, z = q.unbind(-1)
    ww, xx, yy, zz = w * w, x * x, y * y, z * z
    m = torch.stack([ww + xx - yy - zz, 2 * (x * y - w * z), 2 * (x * z + w * y),
                     2 * (x * y + w * z), ww - xx + yy - zz, 2 * (y * z - w * x),
                     2 * (x * z - w * y), 2 * (y * z + w * x), ww - xx - yy + zz], dim=-1)
    return m.reshape(q.shape[:-1] + (3, 3))


def quat_multiply(a, b):
    """quat_affine.py:76-82."""
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw], dim=-1)


def quat_multiply_by_vec(q, v):
    """quat_affine.py:69-74."""
    zero = torch.zeros_like(v[..., :1])
    return quat_multiply(q, torch.cat([zero, v], dim=-1))


def quat_precompose_vec(q, v):
    """quat_affine.py:84-92."""
    return l2_normalize(q + quat_multiply_by_vec(q, v))


def invert_quat(q):
    """quat_affine.py:234-238."""
    sign = torch.tensor([1.0, -1.0, -1.0, -1.0], device=q.device, dtype=q.dtype)
    return q * sign / torch.linalg.norm(q, dim=-1, keepdim=True)


def _half_sinc(half, angle):
    small = angle.abs() < 1e-6
    safe = torch.where(small, torch.ones_like(angle), angle)
    return torch.where(small, 0.5 - angle * angle / 48, torch.sin(half) / safe)


def quat_to_rotvec(q):
    """quat_affine.py:113-131."""
    q = torch.where(q[..., :1] < 0, -q, q)
    half = torch.atan2(torch.linalg.norm(q[..., 1:], dim=-1, keepdim=True), q[..., :1])
    return q[..., 1:] / _half_sinc(half, 2 * half)


def rotvec_to_quat(v):
    """quat_affine.py:133-150."""
    angle = torch.linalg.norm(v, dim=-1, keepdim=True)
    half = 0.5 * angle
    return torch.cat([torch.cos(half), v * _half_sinc(half, angle)], dim=-1)


def rot_to_quat(m):
    """quat_affine.py:180-231: rotation matrix -> quaternion via the best-conditioned of four candidates."""
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = m.reshape(m.shape[:-2] + (9,)).unbind(-1)
    q_abs = torch.sqrt(torch.clamp(torch.stack([1 + m00 + m11 + m22, 1 + m00 - m11 - m22, 1 - m00 + m11 - m22,
                                                1 - m00 - m11 + m22], dim=-1), min=0))
    rows = [[q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01],
            [m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20],
            [m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21],
            [m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2]]
    cand = torch.stack([torch.stack(r, dim=-1) for r in rows], dim=-2) / (2.0 * q_abs[..., None].clamp(min=0.1))
    pick = q_abs.argmax(dim=-1)[..., None, None].expand(q_abs.shape[:-1] + (1, 4))
    return torch.gather(cand, -2, pick).squeeze(-2)
