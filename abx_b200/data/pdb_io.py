"""PDB output / input for the sampler (reference: abx/data/utils.py:187-263 `make_chain` / `save_pdb`,
inference.py:126-164 `postprocess_one` / `postprocess_trajectory`), written without Biopython: the
reference builds Bio.PDB objects only to print fixed-column ATOM records, which is done here directly.
"""
import os

import numpy as np
import torch

from abx_b200.data import residue_tables as rt

RESTYPES_WITH_X = list(rt.restypes) + ['X']
_ATOM14_NAMES = None


def _atom14_names():
    """restype (1-letter, incl. 'X' -> UNK) -> 14 atom names ('' = unused slot), residue_constants.py:330-377."""
    global _ATOM14_NAMES
    if _ATOM14_NAMES is None:
        names = rt.table('restype_atom14_names')                  # [21 (resnames order, UNK last), 14] array of str
        _ATOM14_NAMES = {aa: [str(n) for n in names[i]] for i, aa in enumerate(RESTYPES_WITH_X)}
    return _ATOM14_NAMES


def restype_1to3():
    return {aa: str(r) for aa, r in zip(RESTYPES_WITH_X, rt.table('resnames'))}


def index_to_str_seq(seq):
    """abx/common/utils.py:15-17."""
    return ''.join(RESTYPES_WITH_X[int(i)] for i in seq)


def str_seq_to_index(str_seq):
    """abx/common/utils.py:5-13 (unknown letters map to X = 20)."""
    order = {aa: i for i, aa in enumerate(RESTYPES_WITH_X)}
    return [order.get(aa, 20) for aa in str_seq]


def _atom_line(serial, name, resname, chain, resseq, xyz, bfactor):
    # Bio.PDB's fullname is f'{name:<4s}' in the reference (data/utils.py:205), i.e. left-justified in columns 13-16
    return ('ATOM  %5d %-4s %3s %1s%4d    %8.3f%8.3f%8.3f%6.2f%6.2f          %2s  \n'
            % (serial % 100000, name, resname, chain[:1], resseq, xyz[0], xyz[1], xyz[2], 1.0, bfactor, name[:1].rjust(2)))


def chain_lines(aa_types, coords, chain_id, bfactors, mask=None, serial0=1):
    """data/utils.py:187-233: one chain, residues numbered from 1, atom14 slots with a name are written."""
    names = _atom14_names()
    one_to_three = restype_1to3()
    lines, serial = [], serial0
    for i, (aa, xyz) in enumerate(zip(aa_types, coords)):
        if mask is not None and not bool(mask[i]):
            continue
        resname = one_to_three.get(aa, 'UNK')
        for j, atom_name in enumerate(names.get(aa, names['X'])):
            if atom_name == '':
                continue
            lines.append(_atom_line(serial, atom_name, resname, chain_id, i + 1, xyz[j], float(bfactors[i][j])))
            serial += 1
    if lines:
        lines.append('TER\n')
    return lines, serial


def save_pdb(str_heavy_seq, heavy_chain, str_light_seq, light_chain, coord, pdb_path, pLDDT, antigen_data):
    """data/utils.py:235-263."""
    coord = np.asarray(coord)
    pLDDT = np.asarray(pLDDT, dtype=np.float64)
    n_h, n_l = len(str_heavy_seq), len(str_light_seq)
    assert n_h + n_l == coord.shape[0]
    b = np.repeat(pLDDT[..., None], 37, axis=-1)
    lines, serial = chain_lines(str_heavy_seq, coord[:n_h], heavy_chain, b[:n_h])
    more, serial = chain_lines(str_light_seq, coord[n_h:], light_chain, b[n_h:], serial0=serial)
    lines += more
    start = 0
    chain_ids = np.asarray(antigen_data['antigen_chain_ids'])
    for i, chain in enumerate(antigen_data['antigen_chains']):
        chain_len = int(np.sum(chain_ids == i + 2))
        bb = np.full((chain_len, 37), pLDDT[0] if pLDDT.size else 0.0)
        seq = antigen_data['antigen_str_seq'][start:start + chain_len]
        xyz = np.asarray(antigen_data['antigen_coords'])[start:start + chain_len]
        msk = np.asarray(antigen_data['antigen_coord_mask'])[start:start + chain_len, 1]          # CA present
        start += chain_len
        more, serial = chain_lines(seq, xyz, chain, bb, mask=msk, serial0=serial)
        lines += more
    lines.append('END\n')
    with open(pdb_path, 'w') as f:
        f.writelines(lines)


def postprocess_one(name, str_heavy_seq, str_light_seq, coord, args, pLDDT, antigen_data, time=None):
    """inference.py:126-134."""
    pdb_file = f'{args.output_dir}/{name}@{time:.4f}.pdb' if time else f'{args.output_dir}/{name}.pdb'
    parts = name.split('_')
    save_pdb(str_heavy_seq, parts[1], str_light_seq, parts[2], coord, pdb_file, pLDDT, antigen_data)
    return pdb_file


def _np(x):
    return x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)


def postprocess_trajectory(batch, traj, args):
    """inference.py:136-164."""
    fields = ('name', 'str_heavy_seq', 'str_light_seq', 'antigen_origin_str_seq', 'antigen_origin_atom14_gt_positions',
              'antigen_origin_atom14_gt_exists', 'antigen_origin_chain_ids')
    names, heavy, light, ag_seq, ag_xyz, ag_mask, ag_chain = map(batch.get, fields)
    written = []
    for data in traj:
        pLDDT, seq, coords = _np(data['pLDDT']), _np(data['seq']), _np(data['atom14_results'])
        time = data['time'] if len(traj) > 1 else None
        for i, name in enumerate(names):
            n_h, n_l = len(heavy[i]), len(light[i])
            antigen_data = {'antigen_str_seq': ag_seq[i], 'antigen_coords': ag_xyz[i], 'antigen_coord_mask': ag_mask[i],
                            'antigen_chain_ids': ag_chain[i], 'antigen_chains': list(name.split('_'))[-1]}
            written.append(postprocess_one(name, index_to_str_seq(seq[i, :n_h]), index_to_str_seq(seq[i, n_h:n_h + n_l]),
                                           coords[i, :n_h + n_l], args, pLDDT[i], antigen_data, time))
    return written


# ---- minimal reader (ATOM records -> atom14 arrays per chain) ------------------------------------------
def read_pdb_chains(path):
    """{chain id: dict(str_seq, coords [L,14,3] f32, coord_mask [L,14] bool, resseq [L] int, icode [L])} from the
    ATOM records of the first model; residues with unknown names become 'X'.  Enough for the files the
    reference ships (test_data/*.pdb) and for the files `save_pdb` writes."""
    names = _atom14_names()
    three_to_one = {v: k for k, v in restype_1to3().items()}
    chains, order = {}, []
    with open(path) as f:
        for line in f:
            if line.startswith('ENDMDL'):
                break
            if not line.startswith('ATOM'):
                continue
            atom, alt, resname, chain = line[12:16].strip(), line[16], line[17:20].strip(), line[21]
            if alt not in (' ', 'A'):
                continue
            key = (int(line[22:26]), line[26])
            xyz = (float(line[30:38]), float(line[38:46]), float(line[46:54]))
            c = chains.setdefault(chain, {'res': {}, 'order': []})
            if chain not in order:
                order.append(chain)
            if key not in c['res']:
                c['res'][key] = (three_to_one.get(resname, 'X'), {})
                c['order'].append(key)
            c['res'][key][1][atom] = xyz
    out = {}
    for chain in order:
        c = chains[chain]
        L = len(c['order'])
        coords, mask = np.zeros((L, 14, 3), np.float32), np.zeros((L, 14), bool)
        seq = []
        for i, key in enumerate(c['order']):
            aa, atoms = c['res'][key]
            seq.append(aa)
            for j, an in enumerate(names.get(aa, names['X'])):
                if an and an in atoms:
                    coords[i, j], mask[i, j] = atoms[an], True
        out[chain] = dict(str_seq=''.join(seq), coords=coords, coord_mask=mask,
                          resseq=np.array([k[0] for k in c['order']], np.int64), icode=[k[1] for k in c['order']])
    return out


def ensure_dir(path):
    os.makedirs(path, exist_ok=True)
    return path
