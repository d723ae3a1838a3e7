"""`.npz` structure records -> collated batch dicts for the sampler (reference: abx/data/dataset.py:91-589).

Only what the sampling path needs: one record -> `structure_item` (antibody-centred coordinates, chain /
region ids) -> `patch_around_anchor` (antigen residues within 16 A of the CDR anchors, dataset.py:497-552)
-> `collate` (padded [B,N] tensors, dataset.py:384-466) -> FeatureBuilder.  The reference's SAbDab
preprocessing (ANARCI numbering, mmCIF parsing) is out of scope: records are read in the schema that
preprocessing writes (preprocess/make_ab_data_from_mmcif.py:107-141).
"""
import os

import numpy as np
import torch

from abx_b200.data import residue_tables as rt
from abx_b200.data.pdb_io import str_seq_to_index

UNK = 20


def _patch_idx(a, b, mask_a, mask_b, distance_threshold=16.0):
    """dataset.py:32-42: antigen residues with any atom within the threshold of any anchor atom, +-5 in sequence."""
    diff = a[:, None, :, None, :] - b[None, :, None, :, :]
    mask = mask_a[:, None, :, None] & mask_b[None, :, None, :]
    dist = torch.where(mask, torch.norm(diff, dim=-1), torch.tensor(1e10))
    dist = dist.reshape(a.shape[0], b.shape[0], -1).min(dim=2)[0].min(dim=1)[0]
    hit = torch.nonzero(dist < distance_threshold).reshape(-1).tolist()
    return sorted({i for j in hit for i in range(j - 5, j + 5)})


def patch_around_anchor(data, distance_threshold=16.0):
    """dataset.py:497-552 (inference branch): anchor flags at the residues flanking each CDR and the antigen
    patch around them.  Returns None when no antigen residue is near the paratope, like the reference."""
    cdr = data['antibody_cdr_def']
    n_ab = cdr.shape[0]
    anchor = torch.zeros_like(cdr)
    idx = []
    for enum in rt.cdr_str_to_enum.values():
        flag = cdr == enum
        if not bool(flag.any()):
            continue
        pos = torch.nonzero(flag).reshape(-1)
        left, right = max(0, int(pos.min()) - 1), min(int(pos.max()) + 1, n_ab - 1)
        anchor[left] = enum
        anchor[right] = enum
        sel = [left, right]
        idx.extend(_patch_idx(data['antigen_atom14_gt_positions'], data['antibody_atom14_gt_positions'][sel],
                              data['antigen_atom14_gt_exists'].bool(), data['antibody_atom14_gt_exists'][sel].bool(),
                              distance_threshold))
    n_ag = data['antigen_atom14_gt_positions'].shape[0]
    ca = data['antigen_atom14_gt_positions'][:, rt.atom_order['CA']]
    valid = set(torch.nonzero(ca).reshape(-1, ca.shape[-1] if ca.dim() > 1 else 1)[:, 0].tolist()) if n_ag else set()
    keep = sorted(i for i in set(idx) & valid if 0 <= i < n_ag)
    out = dict(data)
    out['anchor_flag'] = anchor
    for k in ('antigen_atom14_gt_positions', 'antigen_atom14_gt_exists', 'antigen_residx', 'antigen_chain_ids', 'antigen_seq',
              'antigen_cdr_def', 'antigen_mask'):
        out[k] = data[k][keep]
    out['antigen_str_seq'] = ''.join(data['antigen_str_seq'][i] for i in keep)
    for k in ('atom14_gt_positions', 'atom14_gt_exists', 'str_seq', 'residx', 'chain_ids'):
        out[f'antigen_origin_{k}'] = out[f'antigen_{k}']
    return out if keep else None


def structure_item(struc, name, max_antigen_seq_len=32):
    """dataset.py:317-383: tensors of one record, coordinates centred on the antibody CA centroid, antigen cropped
    to the paratope patch and then to `max_antigen_seq_len` residues (dataset.py:129-133, first window)."""
    def arr(key, default):
        return torch.from_numpy(np.asarray(struc[key])) if key in struc else default
    ab_xyz = arr('antibody_coords', torch.zeros(0, 14, 3)).float()
    ab_msk = arr('antibody_coord_mask', torch.zeros(0, 14)).bool()
    ab_chain = arr('antibody_chain_ids', torch.zeros(0, dtype=torch.int64)).long()
    ab_str = str(struc['antibody_str_seq']) if 'antibody_str_seq' in struc else ''
    n_h = int((ab_chain == 0).sum())
    ag_xyz = arr('antigen_coords', torch.zeros(0, 14, 3)).float()
    ag_msk = arr('antigen_coord_mask', torch.zeros(0, 14)).bool()
    ag_str = str(struc['antigen_str_seq']) if 'antigen_str_seq' in struc else ''
    ca = rt.atom_order['CA']
    centre = ab_xyz[:, ca].sum(0) / (ab_msk[:, ca].sum(0, keepdim=True) + 1e-5)
    ab_xyz = (ab_xyz - centre) * ab_msk[..., None]
    ag_xyz = (ag_xyz - centre) * ag_msk[..., None]
    data = dict(
        name=name, str_heavy_seq=ab_str[:n_h], str_light_seq=ab_str[n_h:],
        antibody_seq=torch.tensor(str_seq_to_index(ab_str), dtype=torch.int64),
        antibody_residx=arr('antibody_residx', torch.zeros(0, dtype=torch.int64)).long(),
        antibody_mask=torch.ones(len(ab_str), dtype=torch.bool),
        antibody_atom14_gt_positions=ab_xyz, antibody_atom14_gt_exists=ab_msk,
        antibody_cdr_def=arr('antibody_cdr_def', torch.zeros(0, dtype=torch.int64)).long(), antibody_chain_ids=ab_chain,
        antigen_atom14_gt_positions=ag_xyz, antigen_atom14_gt_exists=ag_msk, antigen_str_seq=ag_str,
        antigen_seq=torch.tensor(str_seq_to_index(ag_str), dtype=torch.int64),
        antigen_mask=torch.ones(len(ag_str), dtype=torch.bool),
        antigen_chain_ids=arr('antigen_chain_ids', torch.zeros(0, dtype=torch.int64)).long(),
        antigen_residx=arr('antigen_residx', torch.zeros(0, dtype=torch.int64)).long(),
        antigen_cdr_def=arr('antigen_cdr_def', torch.zeros(0, dtype=torch.int64)).long())
    data = patch_around_anchor(data)
    if data is None:
        return None
    if len(data['antigen_str_seq']) > max_antigen_seq_len:
        for k, v in list(data.items()):
            if k.startswith('antigen') and 'origin' not in k:
                data[k] = v[:max_antigen_seq_len]
    return data


def _pad(items, length, value=0):
    out = []
    for t in items:
        pad = torch.full((length - t.shape[0],) + tuple(t.shape[1:]), value, dtype=t.dtype)
        out.append(torch.cat([t, pad], dim=0))
    return torch.stack(out, dim=0)


def collate(items):
    """dataset.py:384-466: antibody block padded to the longest antibody, antigen block appended."""
    n_ab = max(len(b['str_heavy_seq']) + len(b['str_light_seq']) for b in items)
    n_ag = max(len(b['antigen_str_seq']) for b in items)

    def both(ab_key, ag_key, value=0):
        return torch.cat([_pad([b[ab_key] for b in items], n_ab, value), _pad([b[ag_key] for b in items], n_ag, value)], dim=1)

    ret = dict(
        name=tuple(b['name'] for b in items),
        seq=both('antibody_seq', 'antigen_seq', UNK), mask=both('antibody_mask', 'antigen_mask'),
        str_heavy_seq=tuple(b['str_heavy_seq'] for b in items), str_light_seq=tuple(b['str_light_seq'] for b in items),
        atom14_gt_positions=both('antibody_atom14_gt_positions', 'antigen_atom14_gt_positions'),
        atom14_gt_exists=both('antibody_atom14_gt_exists', 'antigen_atom14_gt_exists'),
        cdr_def=both('antibody_cdr_def', 'antigen_cdr_def'), chain_id=both('antibody_chain_ids', 'antigen_chain_ids'),
        residx=both('antibody_residx', 'antigen_residx'), anchor_flag=_pad([b['anchor_flag'] for b in items], n_ab))
    ret.update(
        antigen_origin_str_seq=tuple(b['antigen_origin_str_seq'] for b in items),
        antigen_origin_atom14_gt_positions=[b['antigen_origin_atom14_gt_positions'].numpy() for b in items],
        antigen_origin_atom14_gt_exists=[b['antigen_origin_atom14_gt_exists'].numpy() for b in items],
        antigen_origin_chain_ids=[b['antigen_origin_chain_ids'].numpy() for b in items],
        antigen_origin_residx=[b['antigen_origin_residx'].numpy() for b in items])
    return ret


def load(data_dir, name_idx, feats=None, batch_size=1, max_antigen_seq_len=32, rank=0, world_size=1, **_):
    """Generator over collated (and, with `feats`, featurised) batches — the iterable `dataset.load` returns
    (dataset.py:554-571).  With world_size > 1 each rank takes every world_size-th name."""
    from abx_b200.model.features import FeatureBuilder
    builder = FeatureBuilder(feats) if feats else None
    names = list(name_idx)[rank::world_size]
    items = []
    for name in names:
        path = os.path.join(data_dir, name + '.npz')
        if not os.path.exists(path):
            continue
        with np.load(path, allow_pickle=True) as z:
            item = structure_item({k: z[k] for k in z.files}, name, max_antigen_seq_len)
        if item is None:
            continue
        items.append(item)
        if len(items) == batch_size:
            batch = collate(items)
            items = []
            yield builder.build(batch) if builder else batch
    if items:
        batch = collate(items)
        yield builder.build(batch) if builder else batch
