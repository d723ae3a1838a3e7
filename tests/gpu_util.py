"""Helpers for the GPU parity tests (tests marked `gpu`): product objects wired to reference tables."""
import copy
import tempfile

import torch

from tests.util import reference_score_table

DIFF_CONF = {
    'inference_step': 100,
    'diffuse': {'diffuse_trans': True, 'diffuse_rot': True, 'diffuse_seq': True},
    'r3': {'min_b': 0.1, 'max_b': 20.0, 'coordinate_scaling': 0.1},
    'so3': {'num_omega': 1000, 'num_sigma': 1000, 'min_sigma': 0.1, 'max_sigma': 1.5, 'schedule': 'logarithmic',
            'cache_dir': None, 'use_cached_score': True},
    'seq': {'rate_const': 0.3},
}

_built = None


def built_diffuser():
    """FullDiffuser whose IGSO(3) tables were built by the CUDA table kernel (fresh cache dir)."""
    global _built
    if _built is None:
        from abx_b200.diffuser.full_diffuser import FullDiffuser
        conf = copy.deepcopy(DIFF_CONF)
        conf['so3']['cache_dir'] = tempfile.mkdtemp(prefix='abx_igso3_')
        _built = FullDiffuser(conf)
    return _built


def reference_table_diffuser():
    """Same object with the REFERENCE's table rows (tests/golden/igso3.npz; NaN where no golden uses a row)
    so that lookups can be compared bit for bit with the oracle."""
    d = copy.copy(built_diffuser())
    so3 = copy.copy(d._so3_diffuser)
    tab, cdf, pdf = reference_score_table()
    so3._score_norms, so3._cdf, so3._pdf = tab, cdf, pdf
    so3._dev = {}
    d._so3_diffuser = so3
    return d


def cuda(x):
    return x.cuda() if torch.is_tensor(x) else x
