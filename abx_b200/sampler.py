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


STATE_KEYS = ('rigids_t', 'seq_t', 'prev_pos', 'prev_seq', 'prev_pair')
T_KEYS = ('t', 'rot_score_scaling', 'trans_score_scaling')      # rewritten by _set_t_feats inside every iteration


class GraphedReverseStep:
    """One reverse iteration (model forward = 3 trunk passes, get_prev, SE(3)/categorical reverse step) captured
    in a CUDA graph and replayed for every t of the loop: ~1100 kernel launches per iteration become one graph
    launch, so the host never gates the GPU.  The iteration reads and writes five state tensors
    (STATE_KEYS) plus the scalar t; they live in static buffers that the graph's outputs are copied back into."""

    replayed_launches = 0     # abx kernel launches executed through graph replays (bench.py's `gpu_launches`)

    def __init__(self, batch, step_fn, generator=None):
        dev = batch['rigids_t'].device
        self.batch = batch
        self.generator = generator
        self.step_fn = step_fn            # keeps the closure's tensors (masks, dt, placeholders) alive: the graph reads them by address
        # everything else the captured iteration reads: a later batch may reuse this graph only if these are unchanged
        self.features = {k: v for k, v in batch.items() if torch.is_tensor(v) and k not in STATE_KEYS and k not in T_KEYS}
        batch['rigids_t'] = batch['rigids_t'].to(torch.float64).contiguous()     # exact widening; reverse() returns float64
        batch['seq_t'] = batch['seq_t'].long().contiguous()
        self.state = {k: batch[k] for k in STATE_KEYS}
        self.t = torch.ones(batch['rigids_t'].shape[0], device=dev, dtype=torch.float64)
        snapshot = {k: v.clone() for k, v in self.state.items()}
        rng = generator.get_state() if generator is not None else torch.cuda.get_rng_state(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                      # eager run: lazy initialisations happen outside the capture
            step_fn(batch, self.t)
        torch.cuda.current_stream(dev).wait_stream(side)
        self._restore(snapshot)
        if generator is not None:                          # the warm-up must not shift the random stream
            generator.set_state(rng)
        else:
            torch.cuda.set_rng_state(rng, dev)
        self.graph = torch.cuda.CUDAGraph()
        if generator is not None:
            self.graph.register_generator_state(generator)
        from abx_b200 import lib
        n0 = lib.launch_count()
        with torch.cuda.graph(self.graph):
            self.out = step_fn(batch, self.t)
        GraphedReverseStep.last_captured_launches = lib.launch_count() - n0     # abx kernels replayed per iteration
        self._restore(snapshot)

    def _restore(self, snapshot):
        for k, v in snapshot.items():
            self.state[k].copy_(v)
            self.batch[k] = self.state[k]

    def matches(self, batch):
        """True when `batch` holds the same step-invariant features (same keys, shapes, dtypes and values) as the batch this
        graph was captured on — samples of the same complex: the graph (which reads the captured tensors) can serve it."""
        feats = {k: v for k, v in batch.items() if torch.is_tensor(v) and k not in STATE_KEYS and k not in T_KEYS}
        if feats.keys() != self.features.keys():
            return False
        for k, v in feats.items():
            c = self.features[k]
            if v.shape != c.shape or v.dtype != c.dtype or v.device != c.device:
                return False
        return all(v is self.features[k] or bool(torch.equal(v, self.features[k])) for k, v in feats.items())

    def rebind(self, batch, generator=None):
        """Serve a new batch of the same complex with the captured graph: its state is copied into the static buffers, which
        the batch then refers to; the per-iteration entries the graph writes (`t`) are the captured batch's tensors (the loop's
        last, eager forward reads the t of the last replayed iteration, as in the reference); the captured generator continues
        the new generator's stream."""
        for k in STATE_KEYS:
            self.state[k].copy_(batch[k])
            batch[k] = self.state[k]
        for k in T_KEYS:
            if k in self.batch:
                batch[k] = self.batch[k]
        if generator is not None and self.generator is not None and generator is not self.generator:
            self.generator.set_state(generator.get_state())

    def __call__(self, t):
        self.t.fill_(float(t))
        self.graph.replay()
        GraphedReverseStep.replayed_launches += getattr(GraphedReverseStep, 'last_captured_launches', 0)
        new_state, model_out = self.out
        for k in STATE_KEYS:
            self.state[k].copy_(new_state[k])
            self.batch[k] = self.state[k]
        return model_out


def sample_loop(data_init, config, diffuser, model, mode='design', num_t=100, min_t=0.01, center=True,
                self_condition=True, noise_scale=1.0, eps=1e-8, noise_fn=None, generator=None, cache_static=True,
                cuda_graph=False):
    """The loop of `sample_fn` without the PDB writing; returns (trajectory list, final batch).
    `noise_fn(k) -> (z_rot, z_trans, jumps)` injects the k-th step's draws (teacher-forced parity).
    `cuda_graph=True` replays the reverse iteration as a CUDA graph (needs self-conditioning features, i.e.
    the default config; not combined with `noise_fn` or trajectory mode)."""
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

            def reverse_iteration(batch, t_, k=None):
                """inference.py:215-242 for one t > min_t; returns (new state tensors, model_out)."""
                _set_t_feats(batch, diffuser, t_, t_placeholder, with_scalings=False)
                model_out = model(batch)
                fold = model_out['heads']['folding']
                if embed_sc:
                    batch.update(get_prev(batch, model_out, config_model))
                rigids_t, seq_t = diffuser.reverse(
                    rigid_t=batch['rigids_t'], seq_t=batch['seq_t'], rot_score=fold['rot_score'],
                    trans_score=fold['trans_score'], logits_t=model_out['heads']['sequence_module']['logits'],
                    diffuse_mask=diffuse_mask, t=t_, dt=dt, center=center, noise_scale=noise_scale,
                    noise=None if (noise_fn is None or k is None) else noise_fn(k), generator=generator)
                new_state = {'rigids_t': rigids_t, 'seq_t': seq_t}
                new_state.update({key: batch[key] for key in STATE_KEYS[2:] if key in batch})
                return new_state, model_out

            graphed = None
            if cuda_graph and len(reverse_steps) > 2:
                if noise_fn is not None or trajectory or not (embed_sc and self_condition):
                    raise ValueError('cuda_graph=True needs the default self-conditioned design/optimize loop without noise_fn')
                # One graph per (model, diffuser, loop settings, complex): chunks of samples of the same complex (bench.py's steps,
                # the CLI's sample chunks) replay the graph captured for the first one instead of re-capturing ~3300 launches
                # (eager warm-up + capture + instantiation: 0.3 s on a fast host, several times that on a busy one).
                key = (id(diffuser), id(config_model), mode, float(dt), bool(center), float(noise_scale), generator is None,
                       sum(p._version for p in model.parameters()))
                cached = getattr(model, '_abx_graph_cache', None)
                if cached is not None and cached[0] == key and cached[1].matches(batch):
                    graphed = cached[1]
                    graphed.rebind(batch, generator)
                else:
                    model._abx_graph_cache = None           # free the old graph's memory pool before capturing a new one
                    graphed = GraphedReverseStep(batch, reverse_iteration, generator=generator)
                    # the captured graph reads the static-embedding tensors of this call: they must outlive `clear_static()`
                    model._abx_graph_cache = (key, graphed, trunk._static)
            for k, t in enumerate(reverse_steps):
                if t > min_t:
                    if graphed is not None:
                        model_out = graphed(t)
                        rigids_t, seq_t = batch['rigids_t'], batch['seq_t']
                    else:
                        t_ = torch.full((B,), float(t), device=device, dtype=torch.float64)  # float64 (:216)
                        new_state, model_out = reverse_iteration(batch, t_, k)
                        rigids_t, seq_t = new_state['rigids_t'], new_state['seq_t']
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
