"""Drop-in for the reference's `design.py`: the same driver for ONE complex given as a PDB file named
`<code>_<heavy>_<light>_<antigen chains>.pdb` (design.py:152,318,349,407).

The reference numbers the antibody chains with ANARCI (IMGT) to find the CDRs; ANARCI is an un-vendored
dependency that is not available here, so the CDR definition must come with the input: either the PDB is
already IMGT-numbered (residue numbers 27-38 / 56-65 / 105-117 are CDR1/2/3) or a sidecar
`<pdb_file>.cdr.json` gives {"H": [[start,end],...3], "L": [[start,end],...3]} as 0-based index ranges.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from abx_b200 import cli  # noqa: E402

IMGT_CDRS = ((27, 38), (56, 65), (105, 117))


def _cdr_def(chain, first_region, ranges=None):
    """Region ids 0..6 (heavy) / 7..13 (light): FR1, CDR1, FR2, CDR2, FR3, CDR3, FR4 (residue_constants.py:14)."""
    L = len(chain['str_seq'])
    out = np.zeros(L, np.int64)
    if ranges is None:
        num = chain['resseq']
        if num.max() < 105:
            raise SystemExit('design.py: the antibody chains are not IMGT-numbered and no .cdr.json sidecar was given '
                             '(ANARCI is not available offline)')
        bounds = [int(np.searchsorted(num, v)) for lo, hi in IMGT_CDRS for v in (lo, hi + 1)]
    else:
        bounds = [v for lo, hi in ranges for v in (lo, hi + 1)]
    edges = [0] + bounds + [L]
    for r in range(7):
        out[edges[r]:edges[r + 1]] = first_region + r
    return out


def load_batches(args):
    from abx_b200.data import dataset
    from abx_b200.data.pdb_io import read_pdb_chains
    name = os.path.basename(args.pdb_file).rsplit('.', 1)[0]
    _, h_id, l_id, ag_ids = (name.split('_') + ['', '', ''])[:4]
    chains = read_pdb_chains(args.pdb_file)
    side = args.pdb_file + '.cdr.json'
    ranges = json.load(open(side)) if os.path.exists(side) else {}
    H, Lc = chains[h_id], chains[l_id]
    ags = [chains[c] for c in ag_ids if c in chains]
    rec = dict(
        antibody_str_seq=H['str_seq'] + Lc['str_seq'],
        antibody_coords=np.concatenate([H['coords'], Lc['coords']]), antibody_coord_mask=np.concatenate([H['coord_mask'], Lc['coord_mask']]),
        antibody_chain_ids=np.concatenate([np.zeros(len(H['str_seq']), np.int64), np.ones(len(Lc['str_seq']), np.int64)]),
        antibody_residx=np.concatenate([np.arange(len(H['str_seq'])), np.arange(len(Lc['str_seq'])) + 512]),
        antibody_cdr_def=np.concatenate([_cdr_def(H, 0, ranges.get('H')), _cdr_def(Lc, 7, ranges.get('L'))]),
        antigen_str_seq=''.join(a['str_seq'] for a in ags),
        antigen_coords=np.concatenate([a['coords'] for a in ags]) if ags else np.zeros((0, 14, 3), np.float32),
        antigen_coord_mask=np.concatenate([a['coord_mask'] for a in ags]) if ags else np.zeros((0, 14), bool),
        antigen_chain_ids=np.concatenate([np.full(len(a['str_seq']), i + 2, np.int64) for i, a in enumerate(ags)]) if ags else np.zeros(0, np.int64),
        antigen_residx=np.concatenate([np.arange(len(a['str_seq'])) for a in ags]) if ags else np.zeros(0, np.int64),
        antigen_cdr_def=np.full(sum(len(a['str_seq']) for a in ags), 14, np.int64))
    item = dataset.structure_item(rec, name)
    if item is None:
        raise SystemExit('design.py: no antigen residue within 16 A of the CDR anchors')
    yield dataset.collate([item])


if __name__ == '__main__':
    cli.main(cli.build_parser(single_pdb=True).parse_args(), load_batches)
