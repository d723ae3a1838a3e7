"""Residue / atom constant tables (data only).

The numeric tables are the AF2 residue constants the reference builds at import time
(reference abx/common/residue_constants.py:213-377); they are dumped once into
`residue_tables.npz` by oracle/make_residue_tables.py and loaded here.
"""
import functools
import os

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'residue_tables.npz')

restype_num = 20
num_ab_regions = 14                      # residue_constants.py:11
residue_chain_index_offset = 512         # residue_constants.py:12
cdr_str_to_enum = {'H1': 1, 'H2': 3, 'H3': 5, 'L1': 8, 'L2': 10, 'L3': 12}   # residue_constants.py:14
restypes = 'ARNDCQEGHILKMFPSTWYV'
atom_order = {'N': 0, 'CA': 1, 'C': 2, 'CB': 3, 'O': 4}


@functools.lru_cache(maxsize=None)
def tables():
    with np.load(_PATH) as z:
        return {k: z[k] for k in z.files}


def table(name):
    return tables()[name]
