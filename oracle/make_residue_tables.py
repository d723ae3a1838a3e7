"""Dump the numeric residue/atom tables of the reference (abx/common/residue_constants.py:213-377,
AF2 constants + default_rigids.json) into abx_b200/data/residue_tables.npz.

Run in the build container only (needs /root/reference).  The output is constant DATA
(SURVEY.md §2 row 12), committed so that neither the product nor the tests need the reference.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_harness  # noqa: E402

ref_harness.install()
from abx.common import residue_constants as rc  # noqa: E402

names = [
    'restype_atom14_to_atom37', 'restype_atom37_to_atom14', 'restype_atom14_mask', 'restype_atom37_mask',
    'restype_atom37_to_rigid_group', 'restype_atom37_rigid_group_positions', 'restype_atom14_to_rigid_group',
    'restype_atom14_rigid_group_positions', 'restype_rigid_group_default_frame', 'restype_rigidgroup_mask',
    'restype_rigidgroup_base_atom37_idx', 'restype_rigidgroup_base_atom14_idx', 'restype_rigidgroup_is_ambiguous',
    'restype_rigidgroup_rots', 'restype_ambiguous_atoms_swap_index', 'restype_atom14_is_ambiguous',
    'chi_angles_atom_indices',
]
out = {n: np.asarray(getattr(rc, n)) for n in names}
out['chi_angles_mask'] = np.asarray(rc.chi_angles_mask, dtype=np.float32)
out['chi_pi_periodic'] = np.asarray(rc.chi_pi_periodic, dtype=np.float32)
out['restypes'] = np.array(rc.restypes)
out['atom_types'] = np.array(rc.atom_types)
out['resnames'] = np.array(rc.resnames)
atom14_names = np.array([[a for a in rc.restype_name_to_atom14_names[r]] for r in rc.resnames])
out['restype_atom14_names'] = atom14_names
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'abx_b200', 'data', 'residue_tables.npz')
np.savez_compressed(dst, **out)
for k, v in out.items():
    print(k, v.shape, v.dtype)
print('wrote', dst, os.path.getsize(dst))
