"""IMGT region assignment and variable-domain (Fv) extraction for antibody chains WITHOUT ANARCI.

The reference numbers every antibody chain with ANARCI (IMGT scheme), keeps the aligned variable domain
(abx/preprocess/numbering.py:99-131, make_ab_data_from_mmcif.py:143-157) and labels residues by IMGT position
(numbering.py:45-97): FR1 1-26, CDR1 27-38, FR2 39-55, CDR2 56-65, FR3 66-104, CDR3 105-117, FR4 118-128.
ANARCI (HMMER + germline HMMs) is an un-vendored dependency that cannot be installed offline, so the domain is
located here from the anchors IMGT numbering itself is built on:

  1st-CYS 23, CONSERVED-TRP 41, 2nd-CYS 104, J-PHE / J-TRP 118 of the [WF]G.G motif,

and from the fixed framework lengths between them: CDR1 = Cys23+4 .. Trp41-3; FR2 has 17 residues in every germline
(CDR2 starts at Trp41+15); FR3 (66-104) has 38 residues in heavy chains (IMGT gap at 73) and 36 in kappa / lambda chains
(gaps at 73, 81, 82), counted back from Cys104; CDR3 = Cys104+1 .. J-anchor-1; FR4 = 11 (heavy) / 10 (light) residues.
Chains that are already IMGT-numbered (residue 23 = C, 41 = W, 104 = C by residue number) use their numbers directly.
Anything else — no anchor set with plausible CDR lengths — raises NumberingError: a wrong CDR definition silently designs
the wrong residues, so there is no best-effort fallback.
"""
import re

import numpy as np

IMGT_REGIONS = ((1, 26), (27, 38), (39, 55), (56, 65), (66, 104), (105, 117), (118, 128))   # numbering.py:45-65
_FR2_LEN = 17
_FR3_LEN = {'H': 38, 'L': 36}
_FR4_LEN = {'H': 11, 'L': 10}
_TRP41 = re.compile(r'W[A-Z][RQKHL][QKRHEL]')
_J118 = re.compile(r'[WF]G[A-Z]G')


class NumberingError(ValueError):
    pass


def _regions_from_bounds(n, bounds, first_region):
    """bounds = start indices of CDR1, FR2, CDR2, FR3, CDR3, FR4 inside a domain of n residues."""
    edges = [0] + list(bounds) + [n]
    out = np.zeros(n, np.int64)
    for r in range(7):
        out[edges[r]:edges[r + 1]] = first_region + r
    return out


def _is_imgt_numbered(str_seq, resseq):
    at = {int(n): aa for n, aa in zip(resseq, str_seq)}
    return at.get(23) == 'C' and at.get(104) == 'C' and at.get(41) == 'W' and at.get(118) in ('W', 'F')


def _by_imgt_numbers(resseq, first_region):
    num = np.asarray(resseq)
    keep = np.nonzero((num >= 1) & (num <= 128))[0]
    start, end = int(keep[0]), int(keep[-1]) + 1
    if not np.array_equal(keep, np.arange(start, end)):
        raise NumberingError('IMGT-numbered chain is not contiguous within positions 1-128')
    reg = np.zeros(end - start, np.int64)
    for r, (lo, hi) in enumerate(IMGT_REGIONS):
        reg[(num[start:end] >= lo) & (num[start:end] <= hi)] = first_region + r
    return start, end, reg


def _by_anchors(seq, kind):
    best = None
    for c23 in (m.start() for m in re.finditer('C', seq)):
        for m in _TRP41.finditer(seq, c23 + 4 + 5 + 2, c23 + 4 + 12 + 2 + 4):     # CDR1 of 5-12 residues, then 2 FR2 residues
            w41 = m.start()
            cdr1 = (c23 + 4, w41 - 2)                                           # [start, end)
            cdr2_start = w41 + _FR2_LEN - 2
            for fr3_len in (_FR3_LEN[kind], _FR3_LEN['L' if kind == 'H' else 'H']):
                for c104 in (k.start() for k in re.finditer('C', seq[cdr2_start:])):
                    c104 += cdr2_start
                    fr3_start = c104 - fr3_len + 1
                    if not 0 <= fr3_start - cdr2_start <= 12:
                        continue
                    j = _J118.search(seq, c104 + 1 + 2, c104 + 1 + 36 + 4)
                    if j is None:
                        continue
                    cand = dict(c23=c23, cdr1=cdr1, cdr2=(cdr2_start, fr3_start), cdr3=(c104 + 1, j.start()),
                                preferred=fr3_len == _FR3_LEN[kind])
                    if best is None or (cand['preferred'] and not best['preferred']):
                        best = cand
                if best is not None and best['preferred']:
                    break
            if best is not None:
                break
        if best is not None:
            break
    if best is None:
        raise NumberingError('no Cys23 / Trp41 / Cys104 / [WF]G.G anchor set with plausible CDR lengths')
    return best


def assign_regions(str_seq, resseq, kind):
    """-> (start, end, regions): the variable domain is str_seq[start:end]; regions[k] is the reference's region id
    (residue_constants.py:14: heavy 0-6 = FR1, CDR1, FR2, CDR2, FR3, CDR3, FR4; light 7-13) of domain residue k."""
    assert kind in ('H', 'L')
    first = 0 if kind == 'H' else 7
    if _is_imgt_numbered(str_seq, resseq):
        return _by_imgt_numbers(resseq, first)
    a = _by_anchors(str_seq, kind)
    end = min(len(str_seq), a['cdr3'][1] + _FR4_LEN[kind])
    # IMGT positions 1-22 in front of Cys23: heavy and lambda germlines have no residue at position 10 (21 residues), kappa
    # ones do (22); lambda J regions end in TVL, kappa ones in EIK / LEIK
    lam = kind == 'L' and 'TVL' in str_seq[a['cdr3'][1] + 4:end]
    start = max(0, a['c23'] - (21 if kind == 'H' or lam else 22))
    b = [a['cdr1'][0], a['cdr1'][1], a['cdr2'][0], a['cdr2'][1], a['cdr3'][0], a['cdr3'][1]]
    return start, end, _regions_from_bounds(end - start, [v - start for v in b], first)
