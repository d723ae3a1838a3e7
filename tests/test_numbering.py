"""IMGT region assignment without ANARCI (abx_b200/data/numbering.py; reference: abx/preprocess/numbering.py:45-131) on the
reference's own test complexes, and the design.py loader on BASELINE config 1's input (test_data/6ct7_H_L_S.pdb; the
fixture tests/data/6ct7_H_L_S.pdb keeps its backbone + CB atoms)."""
import os
import types

import numpy as np
import pytest

from abx_b200.data.numbering import NumberingError, assign_regions

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# chain sequences as read from test_data/6ct7_H_L_S.pdb (Fab: variable + constant domain, numbered 1..n) and
# test_data/6qd7_X_Z_F|E.pdb (Fv only); expected regions = the IMGT definition (FR1, CDR1, FR2, CDR2, FR3, CDR3, FR4)
CASES = [
    ('H', 'EVQLVESGGGLVEPGGSLRLSCAVSGFDFEKAWMSWVRQAPGQGLQWVARIKSTADGGTTSYAAPVEGRFIISRDDSRNMLYLQMNSLKTEDTAVYYCTSAHWGQGTLVTVSSASTKGPSV'
          'FPLAPSSKSTSGGTAALGCLVKDYFPEPVTVSWNSGALTSGVHTFPAVLQSSGLYSLSSVVTVPSSSLGTQTYICNVNHKPSNTKVDKRVEPK',
     ['EVQLVESGGGLVEPGGSLRLSCAVS', 'GFDFEKAW', 'MSWVRQAPGQGLQWVAR', 'IKSTADGGTT', 'SYAAPVEGRFIISRDDSRNMLYLQMNSLKTEDTAVYYC', 'TSAH',
      'WGQGTLVTVSS']),
    ('L', 'SYELTQPPSVSVSPGQTARITCSGEALPMQFAHWYQQRPGKAPVIVVYKDSERPSGVPERFSGSSSGTTATLTITGVQAEDEADYYCQSPDSTNTYEVFGGGTKLTVLSQPKAAPSVTLF'
          'PPSSEELQANKATLVCLISDFYPGAVTVAWKADSSPVKAGVETTTPSKQSNNKYAASSYLSLTPEQWKSHRSYSCQVTHEGSTVEKTVAPTE',
     ['SYELTQPPSVSVSPGQTARITCSGE', 'ALPMQF', 'AHWYQQRPGKAPVIVVY', 'KDS', 'ERPSGVPERFSGSSSGTTATLTITGVQAEDEADYYC', 'QSPDSTNTYEV',
      'FGGGTKLTVL']),
    ('H', 'VQLLESGGGLVQPGGSLRLSCEASGFPLRDYAMSWVRQAPGRGLQWVSTIGGNDNAANYADSVKGRFTVSRDNSKSTIYLQMNSLRAEDTALYFCAKSVRLSRPSPFDLWGQGSLVTVSS',
     ['VQLLESGGGLVQPGGSLRLSCEAS', 'GFPLRDYA', 'MSWVRQAPGRGLQWVST', 'IGGNDNAA', 'NYADSVKGRFTVSRDNSKSTIYLQMNSLRAEDTALYFC', 'AKSVRLSRPSPFDL',
      'WGQGSLVTVSS']),
    ('L', 'EIVLTQSPATLSLSPGERATLSCRASQSVSTYLAWYQHQPGQAPRLLIYEASNRATGIPARFSGSGSGTEFTLTISSLEPEDVAVYYCQQRASWPLTFGGGTKVEIKR',
     ['EIVLTQSPATLSLSPGERATLSCRAS', 'QSVSTY', 'LAWYQHQPGQAPRLLIY', 'EAS', 'NRATGIPARFSGSGSGTEFTLTISSLEPEDVAVYYC', 'QQRASWPLT',
      'FGGGTKVEIK']),
]


@pytest.mark.parametrize('kind,seq,expected', CASES)
def test_regions_from_conserved_anchors(kind, seq, expected):
    complete = len(expected[0]) >= 25                            # the 6qd7 heavy chain lacks its first residue
    for lead in ('', 'MKHLWFFLLLVAAPRWVLS') if complete else ('',):     # a signal peptide in front must be trimmed as well
        s = lead + seq
        start, end, reg = assign_regions(s, np.arange(1, len(s) + 1), kind)
        dom, base = s[start:end], 0 if kind == 'H' else 7
        got = [''.join(c for c, r in zip(dom, reg) if r == base + i) for i in range(7)]
        assert got == expected
        assert ''.join(got) == dom and start == len(lead)


def test_imgt_numbered_chain_uses_its_numbers():
    # the 6ct7 heavy variable domain with IMGT residue numbers: gaps at 10, 31-34 (CDR1 of 8), 60-61 (CDR2... of 10 has none),
    # 73, and 109-116 (CDR3 of 4)
    kind, seq, expected = CASES[0]
    dom = ''.join(expected)
    lens = [len(e) for e in expected]
    num = ([n for n in range(1, 27) if n != 10] + [27, 28, 29, 30, 35, 36, 37, 38] + list(range(39, 56)) + list(range(56, 66)) +
           [n for n in range(66, 105) if n != 73] + [105, 106, 116, 117] + list(range(118, 129)))
    assert len(num) == len(dom) == sum(lens)
    tail = 'ASTKGPSV'
    start, end, reg = assign_regions(dom + tail, np.array(num + list(range(129, 129 + len(tail)))), kind)
    assert (start, end) == (0, len(dom))
    assert [int((reg == i).sum()) for i in range(7)] == lens


def test_unrecognised_chain_is_refused():
    with pytest.raises(NumberingError):
        assign_regions('MDVFMKGLSKAKEGVVAAAEKTKQGVAEAAGKTKEGVLYVGSKTKEGVVHGVATVAEKTKEQVTNVGGAVVTGVTAVAQKTVEGAGSIAAATGFVKKDQ',
                       np.arange(1, 101), 'H')


def test_design_loader_on_baseline_config_1():
    """design.py on the reference's test complex: Fv trimmed (113 + 108 residues instead of the 214 + 212 of the Fab),
    CDR-H3 = TSAH, ending right before the WGQG motif, the decapeptide antigen kept."""
    import design
    batch = next(design.load_batches(types.SimpleNamespace(pdb_file=os.path.join(ROOT, 'tests', 'data', '6ct7_H_L_S.pdb'))))
    h, l = batch['str_heavy_seq'][0], batch['str_light_seq'][0]
    assert (len(h), len(l)) == (113, 108)
    cdr = batch['cdr_def'][0].numpy()
    assert [int((cdr == r).sum()) for r in range(15)] == [25, 8, 17, 10, 38, 4, 11, 25, 6, 17, 3, 36, 11, 10, 10]
    h3 = np.nonzero(cdr == 5)[0]
    assert h[h3[0]:h3[-1] + 1] == 'TSAH' and h[h3[-1] + 1:h3[-1] + 5] == 'WGQG'
    assert batch['seq'].shape[1] == 113 + 108 + 10
