"""Drop-in for the reference's `design.py`: the same driver for ONE complex given as a PDB file named
`<code>_<heavy>_<light>_<antigen chains>.pdb` (design.py:152,318,349,407).

The reference numbers the antibody chains with ANARCI (IMGT), keeps the variable domains and labels the CDRs by IMGT
position.  ANARCI is not available offline: `abx_b200/data/numbering.py` locates the domain and the regions from the
conserved IMGT anchors (Cys23, Trp41, Cys104, the [WF]G.G J motif) or from the residue numbers when the file is already
IMGT-numbered, and refuses chains where that fails.  A sidecar `<pdb_file>.cdr.json` = {"H": [[start,end] x 3],
"L": [[start,end] x 3]} (0-based inclusive index ranges into the chain as read) overrides it.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from abx_b200 import cli  # noqa: E402
from abx_b200.data.numbering import NumberingError, assign_regions  # noqa: E402


def _domain(chain, kind, ranges=None):
    """-> (slice of the chain that is kept, region ids 0..6 (heavy) / 7..13 (light) of the kept residues)."""
    L = len(chain['str_seq'])
    if ranges is not None:
        bounds = [v for lo, hi in ranges for v in (lo, hi + 1)]
        edges = [0] + bounds + [L]
        out = np.zeros(L, np.int64)
        for r in range(7):
            out[edges[r]:edges[r + 1]] = (0 if kind == 'H' else 7) + r
        return slice(0, L), out
    try:
        start, end, reg = assign_regions(chain['str_seq'], chain['resseq'], kind)
    except NumberingError as e:
        raise SystemExit(f'design.py: cannot locate the IMGT regions of the {"heavy" if kind == "H" else "light"} chain ({e}); '
                         'ANARCI is not available offline — give the CDRs in a <pdb_file>.cdr.json sidecar')
    return slice(start, end), reg


def load_batches(args):
    from abx_b200.data import dataset
    from abx_b200.data.pdb_io import read_pdb_chains
    name = os.path.basename(args.pdb_file).rsplit('.', 1)[0]
    _, h_id, l_id, ag_ids = (name.split('_') + ['', '', ''])[:4]
    chains = read_pdb_chains(args.pdb_file)
    side = args.pdb_file + '.cdr.json'
    ranges = json.load(open(side)) if os.path.exists(side) else {}
    H, Lc = chains[h_id], chains[l_id]
    (hs, h_reg), (ls, l_reg) = _domain(H, 'H', ranges.get('H')), _domain(Lc, 'L', ranges.get('L'))
    H = {k: v[hs] for k, v in H.items()}                       # variable domains only (make_ab_data_from_mmcif.py:152)
    Lc = {k: v[ls] for k, v in Lc.items()}
    ags = [chains[c] for c in ag_ids if c in chains]
    rec = dict(
        antibody_str_seq=H['str_seq'] + Lc['str_seq'],
        antibody_coords=np.concatenate([H['coords'], Lc['coords']]), antibody_coord_mask=np.concatenate([H['coord_mask'], Lc['coord_mask']]),
        antibody_chain_ids=np.concatenate([np.zeros(len(H['str_seq']), np.int64), np.ones(len(Lc['str_seq']), np.int64)]),
        antibody_residx=np.concatenate([np.arange(len(H['str_seq'])), np.arange(len(Lc['str_seq'])) + 512]),
        antibody_cdr_def=np.concatenate([h_reg, l_reg]),
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
