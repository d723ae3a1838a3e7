"""Synthetic antibody–antigen complexes: the stand-in for the SAbDab `.npz` inputs,
which are not available offline.

`synthetic_complex` returns the dict the reference's `collate_fn` hands to its FeatureBuilder
(abx/data/dataset.py:400-461): heavy + light Fv chains with IMGT-like region ids, CDR anchor flags
(dataset.py:497-508), and an antigen patch.  All randomness is numpy PCG64 (platform independent).
"""
import numpy as np
import torch

from abx_b200.data import residue_tables as rt

# IMGT-like region lengths (FR1,CDR1,FR2,CDR2,FR3,CDR3,FR4); heavy ids 0-6, light ids 7-13, antigen id 14.
HEAVY_REGIONS = (25, 8, 17, 8, 38, 13, 11)      # 120 residues
LIGHT_REGIONS = (26, 6, 17, 3, 36, 9, 13)       # 110 residues
SMALL_HEAVY = (4, 4, 4, 3, 5, 6, 3)             # 29 residues (test-sized)
SMALL_LIGHT = (4, 3, 4, 3, 5, 4, 3)             # 26 residues


def _region_ids(lengths, first_id):
    return np.concatenate([np.full(n, first_id + k, dtype=np.int64) for k, n in enumerate(lengths)])


def synthetic_complex(n_antigen=120, heavy=HEAVY_REGIONS, light=LIGHT_REGIONS, seed=0, batch_size=1,
                      name='synt_H_L_A'):
    """Collate-shaped batch of `batch_size` copies of one synthetic complex (N = |H| + |L| + n_antigen)."""
    rng = np.random.default_rng(seed)
    n_h, n_l = int(sum(heavy)), int(sum(light))
    n_ab = n_h + n_l
    n = n_ab + n_antigen

    seq = rng.integers(0, rt.restype_num, size=n).astype(np.int64)
    steps = rng.standard_normal((n, 3))
    steps = 3.8 * steps / np.linalg.norm(steps, axis=-1, keepdims=True)
    ca = np.cumsum(steps, axis=0)
    atoms = ca[:, None, :] + 1.5 * rng.standard_normal((n, 14, 3))
    atoms[:, 1] = ca
    exists = rt.table('restype_atom14_mask')[seq].astype(bool)          # [n, 14]
    # centre on the antibody CA centroid (dataset.py:354-366) and zero the absent atoms
    atoms = atoms - ca[:n_ab].mean(axis=0)[None, None, :]
    atoms = (atoms * exists[..., None]).astype(np.float32)

    cdr_def = np.concatenate([_region_ids(heavy, 0), _region_ids(light, 7),
                              np.full(n_antigen, rt.num_ab_regions, dtype=np.int64)])
    chain_id = np.concatenate([np.zeros(n_h), np.ones(n_l), np.full(n_antigen, 2)]).astype(np.int64)
    residx = np.concatenate([np.arange(n_h), np.arange(n_l) + rt.residue_chain_index_offset,
                             np.arange(n_antigen)]).astype(np.int64)

    anchor = np.zeros(n_ab, dtype=np.int64)
    for cdr in rt.cdr_str_to_enum.values():
        idx = np.nonzero(cdr_def[:n_ab] == cdr)[0]
        if idx.size:
            anchor[max(0, idx[0] - 1)] = cdr
            anchor[min(idx[-1] + 1, n_ab - 1)] = cdr

    letters = np.array(list(rt.restypes))
    str_h = ''.join(letters[seq[:n_h]])
    str_l = ''.join(letters[seq[n_h:n_ab]])
    str_a = ''.join(letters[seq[n_ab:]])

    def rep(x):
        t = torch.from_numpy(np.ascontiguousarray(x))
        return t[None].repeat(batch_size, *([1] * t.ndim)).contiguous()

    batch = dict(
        name=tuple([name] * batch_size),
        seq=rep(seq), mask=rep(np.ones(n, dtype=bool)),
        str_heavy_seq=tuple([str_h] * batch_size), str_light_seq=tuple([str_l] * batch_size),
        atom14_gt_positions=rep(atoms), atom14_gt_exists=rep(exists),
        cdr_def=rep(cdr_def), chain_id=rep(chain_id), residx=rep(residx), anchor_flag=rep(anchor),
        antigen_origin_str_seq=tuple([str_a] * batch_size),
        antigen_origin_atom14_gt_positions=[atoms[n_ab:].copy() for _ in range(batch_size)],
        antigen_origin_atom14_gt_exists=[exists[n_ab:].copy() for _ in range(batch_size)],
        antigen_origin_chain_ids=[chain_id[n_ab:].copy() for _ in range(batch_size)],
        antigen_origin_residx=[residx[n_ab:].copy() for _ in range(batch_size)],
    )
    return batch


def small_complex(n_antigen=9, seed=0, batch_size=1):
    """Test-sized complex: N = 29 + 26 + n_antigen (64 by default)."""
    return synthetic_complex(n_antigen=n_antigen, heavy=SMALL_HEAVY, light=SMALL_LIGHT, seed=seed,
                             batch_size=batch_size, name='tiny_H_L_A')
