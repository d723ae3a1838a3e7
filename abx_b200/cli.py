"""Command-line drivers shared by `inference.py` and `design.py` at the repository root (reference:
inference.py:59-418, design.py).  Same flags, same output layout:

    <output_dir>/<mode>/reference/<name>.pdb, <output_dir>/<mode>/0000..NNNN/<name>.pdb
    optimize: <output_dir>/optimize/OPT-<step>/<k>/<name>.pdb;  trajectory: <name>@<t:.4f>.pdb per step

Differences from the reference, all on the host side: samples of one complex are batched along B
(`--samples_per_batch`) instead of re-running the loader once per sample; with several GPUs sample k goes to
rank k mod world (the reference's spawn skeleton would repeat all work on every rank);
`--model random[:seed]` runs with seeded random weights when no checkpoint is available; errors are not
swallowed (inference.py:301-302 logs and continues).
"""
import argparse
import copy
import json
import logging
import os
import time

import numpy as np
import torch


def build_parser(single_pdb):
    p = argparse.ArgumentParser()
    p.add_argument('--gpu_list', type=int, nargs='+', default=[0])
    p.add_argument('--device', type=str, choices=['gpu', 'cpu'], default='gpu')
    p.add_argument('--model', type=str, required=True, help="checkpoint ({'model_state_dict': ...}) or random[:seed]")
    p.add_argument('--model_features', type=str, required=True)
    p.add_argument('--model_config', type=str, required=True)
    if single_pdb:
        p.add_argument('--pdb_file', type=str, required=True)
    else:
        p.add_argument('--name_idx', type=str, required=True)
        p.add_argument('--data_dir', type=str, required=True)
    p.add_argument('--output_dir', type=str, required=True)
    p.add_argument('--mode', type=str, choices=['design', 'optimize', 'trajectory'], default='design')
    p.add_argument('--batch_size', type=int, default=1)
    p.add_argument('--num_samples', type=int, default=100)
    p.add_argument('--verbose', action='store_true')
    # additions
    p.add_argument('--samples_per_batch', type=int, default=8, help='independent samples of one complex batched per forward')
    p.add_argument('--num_t', type=int, default=None, help='override diffuser.inference_step')
    p.add_argument('--seed', type=int, default=0,
                   help='base seed; each batch of samples draws from one generator seeded with a hash of (seed, sample '
                        'indices of the batch), so results depend on --samples_per_batch and on the number of GPUs')
    p.add_argument('--unsafe_pickle_checkpoint', action='store_true',
                   help='load the checkpoint with torch.load(weights_only=False) (arbitrary unpickling; trusted files only)')
    p.add_argument('--no_cuda_graph', action='store_true')
    return p


def worker_load(args, device):
    """inference.py:84-125."""
    from abx_b200.diffuser.full_diffuser import FullDiffuser
    from abx_b200.model.abx import ScoreNetwork
    from abx_b200.utils.weights import load_seeded_
    with open(args.model_features) as f:
        feats = json.load(f)
    with open(args.model_config) as f:
        config = json.load(f)
    config['diffuser']['so3']['use_cached_score'] = True                                   # inference.py:99
    if config['model']['embeddings_and_seqformer'].get('esm', {}).get('enabled'):
        if not args.model.startswith('random'):
            raise RuntimeError(
                f'{args.model_config} enables ESM2 conditioning (embeddings_and_seqformer.esm.enabled=true), which this '
                'implementation does not provide (fair-esm and the esm2_t36_3B weights are not available offline). A '
                'checkpoint trained with that config carries esm_embed_weights / proj_esm_embed parameters and would run '
                'without the conditioning it was trained with. Use a checkpoint trained with esm.enabled=false together '
                'with a config that says so.')
        logging.warning('ESM2 conditioning is not supported: running the seeded random model with esm.enabled=false')
        config['model']['embeddings_and_seqformer']['esm']['enabled'] = False
    diffuser = FullDiffuser.get(config['diffuser'])
    model = ScoreNetwork(config['model'], diffuser)
    if args.model.startswith('random'):
        seed = int(args.model.split(':')[1]) if ':' in args.model else 0
        load_seeded_(model, seed)
        logging.warning('running with seeded random weights (seed %d), not a trained checkpoint', seed)
    else:
        ckpt = torch.load(args.model, map_location='cpu', weights_only=not getattr(args, 'unsafe_pickle_checkpoint', False))
        model.load_state_dict(ckpt['model_state_dict'], strict=True)                           # inference.py:105
    model = model.to(device).eval()
    optimize_steps = None
    for name, kw in feats:
        if 'device' in kw:
            kw['device'] = device
        if name == 'make_diffuser_features':
            kw['diff_conf'] = config['diffuser']
            kw['diffuser'] = diffuser
            optimize_steps = kw.pop('optimize_steps', None)
            if args.mode != 'optimize':
                kw['diff_conf'] = dict(kw['diff_conf'])
    return feats, model, diffuser, config, optimize_steps


def _repeat_batch(batch, n):
    """n copies of a collated single-complex batch along B (independent samples of the same complex)."""
    out = {}
    for k, v in batch.items():
        if torch.is_tensor(v):
            out[k] = v.repeat(n, *([1] * (v.dim() - 1)))
        elif isinstance(v, (tuple, list)):
            out[k] = type(v)(list(v) * n) if isinstance(v, list) else tuple(list(v) * n)
        else:
            out[k] = v
    return out


def _reference_batch(batch, args, out_dir):
    from abx_b200.data.pdb_io import postprocess_trajectory
    n_ab = batch['anchor_flag'].shape[1]
    ref = [{'atom14_results': batch['atom14_gt_positions'][:, :n_ab], 'seq': batch['seq'][:, :n_ab],
            'pLDDT': np.full((batch['seq'].shape[0], n_ab), 100.0), 'time': None}]
    a = copy.copy(args)
    a.output_dir = out_dir
    os.makedirs(out_dir, exist_ok=True)
    postprocess_trajectory(batch, ref, a)


def run(args, raw_batches, rank=0, world=1):
    """Sample every raw (collated, un-featurised) batch `num_samples` times and write the PDB files."""
    from abx_b200 import parallel, sampler
    from abx_b200.data.pdb_io import postprocess_trajectory
    from abx_b200.model.features import FeatureBuilder
    if args.device != 'gpu' or not torch.cuda.is_available():
        raise RuntimeError('abx_b200 runs on CUDA devices only (sm_100a kernels; there is no CPU path)')
    device = torch.device('cuda', args.gpu_list[rank] if rank < len(args.gpu_list) else rank)
    torch.cuda.set_device(device)
    feats, model, diffuser, config, optimize_steps = worker_load(args, device)
    num_t = args.num_t or config['diffuser']['inference_step']
    if args.num_t:                                   # optimize mode derives its start time from inference_step (features.py:194-203)
        config['diffuser']['inference_step'] = num_t
        for name, kw in feats:
            if name == 'make_diffuser_features':
                kw['diff_conf'] = dict(kw['diff_conf'], inference_step=num_t)
    root = os.path.join(args.output_dir, args.mode)
    os.makedirs(root, exist_ok=True)
    plans = [(None, root)] if args.mode != 'optimize' else [(s, os.path.join(root, f'OPT-{s}')) for s in (optimize_steps or [])]
    mine = parallel.shard_samples(args.num_samples, rank, world)
    for raw in raw_batches:
        if rank == 0:
            _reference_batch(raw, args, os.path.join(root, 'reference'))
        for opt_step, out_root in plans:
            fcfg = [[name, dict(kw)] for name, kw in feats]          # shallow: kw holds the diffuser object
            for name, kw in fcfg:
                if name == 'make_diffuser_features':
                    kw['diffuser'] = diffuser
                    if opt_step is not None:
                        kw['diff_conf'] = dict(kw['diff_conf'], opt_step=opt_step)
            for c0 in range(0, len(mine), args.samples_per_batch):
                ks = mine[c0:c0 + args.samples_per_batch]
                t0 = time.time()
                chunk_seed = parallel.chunk_seed(args.seed, ks)
                torch.manual_seed(chunk_seed)
                gen = torch.Generator(device=device).manual_seed(chunk_seed)
                batch = FeatureBuilder(fcfg).build(_repeat_batch(raw, len(ks)))
                traj, final = sampler.sample_loop(batch, config, diffuser, model, mode=args.mode, num_t=num_t, generator=gen,
                                                  cuda_graph=not args.no_cuda_graph and args.mode != 'trajectory')
                for d in traj:
                    for key in ('seq', 'pLDDT', 'atom14_results'):
                        if torch.is_tensor(d[key]):
                            d[key] = d[key].detach().to('cpu').numpy()
                n_complex = len(raw['name'])
                for j, k in enumerate(ks):                                   # sample k -> <out_root>/<k:04d>/<name>.pdb
                    a = copy.copy(args)
                    a.output_dir = os.path.join(out_root, f'{k:04d}')
                    os.makedirs(a.output_dir, exist_ok=True)
                    sl = slice(j * n_complex, (j + 1) * n_complex)
                    one = {key: (v[sl] if isinstance(v, (tuple, list)) else v) for key, v in final.items()
                           if key in ('name', 'str_heavy_seq', 'str_light_seq') or key.startswith('antigen_origin')}
                    postprocess_trajectory(one, [{**d, 'seq': d['seq'][sl], 'pLDDT': d['pLDDT'][sl],
                                                  'atom14_results': d['atom14_results'][sl]} for d in traj], a)
                logging.info('names %s samples %s: %.2f s', ','.join(raw['name']), ks, time.time() - t0)


def _spawn_entry(rank, args, loader_fn):
    import torch.distributed as dist
    world = len(args.gpu_list)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29512')
        dist.init_process_group('nccl', rank=rank, world_size=world)
    try:
        run(args, loader_fn(args), rank, world)
    finally:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()


def main(args, loader_fn):
    logging.basicConfig(level=logging.DEBUG if args.verbose else logging.INFO,
                        format='%(asctime)s [%(levelname)s] %(message)s')
    os.makedirs(os.path.join(args.output_dir, args.mode), exist_ok=True)
    logging.info('Arguments: %s', args)
    if len(args.gpu_list) > 1:
        import torch.multiprocessing as mp
        mp.spawn(_spawn_entry, args=(args, loader_fn), nprocs=len(args.gpu_list), join=True)
    else:
        _spawn_entry(0, args, loader_fn)
