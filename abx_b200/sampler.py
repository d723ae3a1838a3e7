"""Reverse-diffusion sampling loop (reference: `sample_fn`, inference.py:180-273 == design.py:182-275).

`sample_fn` keeps the reference's signature and per-step semantics — self-conditioning warm-up at t = 1
(float32 t), 99 model + reverse steps with float64 t, final x0 call at t = min_t that re-uses the previous
step's t features — but removes the per-step host round trips: nothing is copied to the host inside the
loop unless `args.mode == 'trajectory'`, the step-invariant trunk embeddings are computed once per complex,
and every batch element is an independent sample of the same complex (the reference runs batch_size 1).
"""
import copy

import numpy as np
import torch

from abx_b200.model.abx import get_prev


def _set_t_feats(feats, diffuser, t, t_placeholder, with_scalings=True):
    """inference.py:166-171.  The score scalings only feed the training loss; the loop skips them."""
    feats['t'] = t * t_placeholder
    if with_scalings:
        rot, trans = diffuser.score_scaling(feats['t'])
        feats['rot_score_scaling'] = rot * t_placeholder
        feats['trans_score_scaling'] = trans * t_placeholder
    return feats


def _self_conditioning(batch, model, config):
    """inference.py:173-177."""
    batch.update(get_prev(batch, model(batch), config))
    return batch


def reverse_grid(num_t=100, min_t=0.01):
    """inference.py:197-199."""
    return np.linspace(min_t, 1.0, num_t)[::-1]


def _frame_output(batch, model_out, seq_t, diffuse_mask, antibody_len, t, to_host):
    plddt = model_out['heads']['predicted_lddt']['pLDDT']
    item = torch.sum(plddt * diffuse_mask, dim=1) / torch.sum(diffuse_mask, dim=1)           # inference.py:253-255
    plddt = torch.tile(item[:, None], (1, antibody_len))
    atom14 = model_out['heads']['folding']['final_atom14_positions'][:, :antibody_len]
    seq = torch.clamp(seq_t[:, :antibody_len], min=0, max=19).long()
    if to_host:
        plddt, seq = plddt.to('cpu').numpy(), seq.to('cpu').numpy()
    return {'seq': seq, 'atom14_results': atom14, 'pLDDT': plddt, 'time': t}


def sample_loop(data_init, config, diffuser, model, mode='design', num_t=100, min_t=0.01, center=True,
                self_condition=True, noise_scale=1.0, eps=1e-8, noise_fn=None, generator=None, cache_static=True):
    """The loop of `sample_fn` without the PDB writing; returns (trajectory list, final batch).
    `noise_fn(k) -> (z_rot, z_trans, jumps)` injects the k-th step's draws (teacher-forced parity)."""
    config_model = config['model'] if isinstance(config, dict) else config.model
    embed_sc = config_model['heads']['diffusion_module']['embed']['embed_self_conditioning']
    batch = {k: (v.clone() if torch.is_tensor(v) else copy.deepcopy(v)) for k, v in data_init.items()}
    device = batch['rigids_t'].device
    diffuse_mask = (1 - batch['fixed_mask']) * batch['atom14_gt_exists'][..., 0]             # :192-193
    antibody_len = batch['anchor_flag'].shape[1]
    B = batch['rigids_t'].shape[0]
    t_placeholder = torch.ones(B, device=device, dtype=torch.float32)
    reverse_steps = reverse_grid(num_t, min_t)
    dt = torch.tensor(1 / num_t)                          # host scalar: its float32 value is what the kernels get
    if mode == 'optimize':
        opt_step = float(batch['t'][0])
        if opt_step < 1.0:
            reverse_steps = reverse_steps[reverse_steps <= opt_step + eps]
    trajectory = mode == 'trajectory'
    trunk = model.impl.seqformer
    traj = []
    with torch.no_grad():
        if cache_static:
            trunk.cache_static(batch)
        try:
            if embed_sc and self_condition and len(reverse_steps) > 0:
                batch = _set_t_feats(batch, diffuser, reverse_steps[0], t_placeholder, with_scalings=False)
                batch = _self_conditioning(batch, model, config_model)
            data = None
            for k, t in enumerate(reverse_steps):
                if t > min_t:
                    t_ = torch.full((B,), float(t), device=device, dtype=torch.float64)      # float64 (:216)
                    batch = _set_t_feats(batch, diffuser, t_, t_placeholder, with_scalings=False)
                    model_out = model(batch)
                    fold = model_out['heads']['folding']
                    if embed_sc:
                        batch.update(get_prev(batch, model_out, config_model))
                    rigids_t, seq_t = diffuser.reverse(
                        rigid_t=batch['rigids_t'], seq_t=batch['seq_t'], rot_score=fold['rot_score'],
                        trans_score=fold['trans_score'], logits_t=model_out['heads']['sequence_module']['logits'],
                        diffuse_mask=diffuse_mask, t=t_, dt=dt, center=center, noise_scale=noise_scale,
                        noise=None if noise_fn is None else noise_fn(k), generator=generator)
                else:
                    model_out = model(batch)                                                  # :244-247
                    rigids_t = model_out['heads']['folding']['rigids']
                    seq_t = model_out['heads']['sequence_module']['seq_0']
                batch['rigids_t'], batch['seq_t'] = rigids_t, seq_t
                if trajectory or k == len(reverse_steps) - 1:
                    data = _frame_output(batch, model_out, seq_t, diffuse_mask, antibody_len, t, to_host=trajectory)
                    traj.append(data)
        finally:
            if cache_static:
                trunk.clear_static()
    return traj, batch


def sample_fn(data_init, config, diffuser, model, args, num_t=100, min_t=0.01, center=True, self_condition=True,
              noise_scale=1.0, eps=1e-8):
    """Drop-in for the reference `sample_fn`: runs the loop and writes the PDB file(s) (postprocess_trajectory)."""
    from abx_b200.data.pdb_io import postprocess_trajectory
    traj, batch = sample_loop(data_init, config, diffuser, model, mode=args.mode, num_t=num_t, min_t=min_t, center=center,
                              self_condition=self_condition, noise_scale=noise_scale, eps=eps)
    for d in traj:
        if torch.is_tensor(d['seq']):
            d['seq'], d['pLDDT'] = d['seq'].to('cpu').numpy(), d['pLDDT'].to('cpu').numpy()
    postprocess_trajectory(batch, traj, args)
    return traj
