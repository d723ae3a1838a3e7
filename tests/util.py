"""Shared helpers for the parity tests: golden loaders and oracle construction."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def golden(name):
    with np.load(os.path.join(GOLDEN, name + '.npz')) as z:
        return {k: torch.from_numpy(z[k]) if z[k].dtype.kind in 'fiub' else z[k] for k in z.files}


def reference_score_table():
    """[1000,1000] IGSO(3) score-norm table holding the REFERENCE's rows (tests/golden/igso3.npz) at
    every sigma index any golden uses; NaN elsewhere so an unexpected row is caught."""
    g = golden('igso3')
    tab = torch.full((1000, 1000), float('nan'))
    tab[g['rows'].long()] = g['score_norms']
    cdf = torch.full((1000, 1000), float('nan'))
    cdf[g['rows'].long()] = g['cdf']
    pdf = torch.full((1000, 1000), float('nan'))
    pdf[g['rows'].long()] = g['pdf']
    return tab, cdf, pdf


def oracle_diffuser():
    from oracle.diffusers import OracleDiffuser
    tab, cdf, pdf = reference_score_table()
    return OracleDiffuser(tab, cdf=cdf, pdf=pdf)


def maxabs(a, b):
    return float((torch.as_tensor(a).double() - torch.as_tensor(b).double()).abs().max())


def batch_from_golden(g, prefix='batch_'):
    """Rebuild the model-boundary batch dict from a golden file."""
    b = {k[len(prefix):]: v.clone() for k, v in g.items() if k.startswith(prefix) and not k.startswith(prefix + 'gt_frame')}
    b['rigidgroups_gt_frames'] = (g[prefix + 'gt_frame_rots'], g[prefix + 'gt_frame_trans'])
    b['is_recycling'] = False
    return b


_REF_SHAPES = None


def reference_param_shapes():
    """name -> shape of the reference ScoreNetwork state_dict (190 tensors, ESM disabled); stored as a
    small JSON next to the goldens so no test needs the reference to know the checkpoint layout."""
    global _REF_SHAPES
    if _REF_SHAPES is None:
        import json
        with open(os.path.join(GOLDEN, 'state_dict_shapes.json')) as f:
            _REF_SHAPES = {k: tuple(v) for k, v in json.load(f).items()}
    return _REF_SHAPES


def seeded_params(seed=0):
    from abx_b200.utils.weights import seeded_state_dict
    return seeded_state_dict(reference_param_shapes(), seed)
