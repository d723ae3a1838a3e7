"""ORACLE (test infrastructure, not product code) — the reverse-diffusion loop of the reference
(`sample_fn`, inference.py:180-273) restated over oracle/model.py + oracle/diffusers.py, with the
noise injected per step so it is deterministic and device independent.
"""
import numpy as np
import torch

from oracle import model as M


def reverse_grid(num_t=100, min_t=0.01):
    """inference.py:197-199."""
    return np.linspace(min_t, 1.0, num_t)[::-1]


def self_condition(P, diffuser, batch, t):
    """inference.py:209-211 (`_set_t_feats` + `_self_conditioning`): warm-up call at float32 t."""
    B = batch['rigids_t'].shape[0]
    batch['t'] = t * torch.ones(B, dtype=torch.float32)                      # :167 (numpy float * f32 tensor)
    out = M.score_network(P, diffuser, batch)
    batch.update(M.get_prev(batch, out))
    return batch


def diffuse_mask_of(batch):
    """inference.py:192-193."""
    return (1 - batch['fixed_mask']) * batch['atom14_gt_exists'][..., 0]


def sample_step(P, diffuser, batch, t, dt, noise, last=False):
    """One iteration of the loop body (inference.py:213-258).  `noise` = (z_rot, z_trans, jumps) or
    a callable rate_dt -> jumps in the third slot.  Returns the model outputs; mutates `batch`."""
    B = batch['rigids_t'].shape[0]
    if not last:
        t_ = torch.tile(torch.tensor(np.float64(t)), (B,))                   # float64 (:216)
        batch['t'] = t_ * torch.ones(B, dtype=torch.float32)                 # :167 -> float64
        out = M.score_network(P, diffuser, batch)
        batch.update(M.get_prev(batch, out))
        rig, seq = diffuser.reverse(batch['rigids_t'], batch['seq_t'], out['rot_score'], out['trans_score'],
                                    out['logits'], t_, dt, diffuse_mask_of(batch), *noise)
    else:
        out = M.score_network(P, diffuser, batch)                            # :244-247
        rig, seq = out['rigids'], out['seq_0']
    batch['rigids_t'], batch['seq_t'] = rig, seq
    return out
