"""Host-side IO of the drivers (CPU): PDB writer/reader round trip (reference data/utils.py:187-263,
inference.py:126-164), `.npz` record -> collated batch (dataset.py:317-466,497-552), CLI surface."""
import os
import types

import numpy as np
import torch

from abx_b200.data import dataset, pdb_io
from abx_b200.data.synthetic import small_complex, synthetic_complex


def test_pdb_writer_reader_round_trip(tmp_path):
    b = small_complex(batch_size=1)
    n_ab = b['anchor_flag'].shape[1]
    traj = [{'seq': b['seq'][:, :n_ab], 'atom14_results': b['atom14_gt_positions'][:, :n_ab],
             'pLDDT': np.full((1, n_ab), 87.5), 'time': 0.01}]
    files = pdb_io.postprocess_trajectory(b, traj, types.SimpleNamespace(output_dir=str(tmp_path)))
    assert files == [os.path.join(str(tmp_path), 'tiny_H_L_A.pdb')]
    chains = pdb_io.read_pdb_chains(files[0])
    assert list(chains) == ['H', 'L', 'A']
    assert chains['H']['str_seq'] == b['str_heavy_seq'][0] and chains['L']['str_seq'] == b['str_light_seq'][0]
    n_h = len(b['str_heavy_seq'][0])
    xyz = b['atom14_gt_positions'][0].numpy()
    exists = b['atom14_gt_exists'][0].numpy()
    got = np.concatenate([chains['H']['coords'], chains['L']['coords']])
    assert np.array_equal(np.concatenate([chains['H']['coord_mask'], chains['L']['coord_mask']]), exists[:n_ab])
    assert np.abs(got - xyz[:n_ab] * exists[:n_ab, :, None]).max() < 1e-3            # %8.3f columns
    line = open(files[0]).readline()
    assert line.startswith('ATOM      1 N  ') and line[21] == 'H' and float(line[60:66]) == 87.5
    assert chains['H']['coords'].shape == (n_h, 14, 3)


def test_trajectory_file_names(tmp_path):
    b = small_complex(batch_size=1)
    n_ab = b['anchor_flag'].shape[1]
    frame = {'seq': b['seq'][:, :n_ab], 'atom14_results': b['atom14_gt_positions'][:, :n_ab], 'pLDDT': np.zeros((1, n_ab))}
    traj = [dict(frame, time=1.0), dict(frame, time=0.505)]
    files = pdb_io.postprocess_trajectory(b, traj, types.SimpleNamespace(output_dir=str(tmp_path)))
    assert [os.path.basename(f) for f in files] == ['tiny_H_L_A@1.0000.pdb', 'tiny_H_L_A@0.5050.pdb']   # inference.py:129


def _record(b):
    n_ab = b['anchor_flag'].shape[1]
    g = lambda k, sl: b[k][0, sl].numpy()     # noqa: E731
    ab, ag = slice(0, n_ab), slice(n_ab, None)
    return dict(antibody_str_seq=b['str_heavy_seq'][0] + b['str_light_seq'][0], antibody_coords=g('atom14_gt_positions', ab),
                antibody_coord_mask=g('atom14_gt_exists', ab), antibody_chain_ids=g('chain_id', ab), antibody_residx=g('residx', ab),
                antibody_cdr_def=g('cdr_def', ab), antigen_str_seq=b['antigen_origin_str_seq'][0],
                antigen_coords=g('atom14_gt_positions', ag), antigen_coord_mask=g('atom14_gt_exists', ag),
                antigen_chain_ids=g('chain_id', ag), antigen_residx=g('residx', ag), antigen_cdr_def=g('cdr_def', ag))


def test_npz_records_collate_like_the_reference_dataset(tmp_path):
    b = synthetic_complex(n_antigen=60, batch_size=1)
    np.savez(tmp_path / 'synt_H_L_A.npz', **_record(b))
    small = small_complex(batch_size=1)
    np.savez(tmp_path / 'tiny_H_L_A.npz', **_record(small))
    batches = list(dataset.load(str(tmp_path), ['synt_H_L_A', 'missing_X_Y_Z', 'tiny_H_L_A'], batch_size=2))
    assert len(batches) == 1
    out = batches[0]
    n_ab = b['anchor_flag'].shape[1]
    assert out['name'] == ('synt_H_L_A', 'tiny_H_L_A')
    assert out['anchor_flag'].shape == (2, n_ab) and torch.equal(out['anchor_flag'][0], b['anchor_flag'][0])
    assert out['seq'].shape[1] == n_ab + 32                                    # antigen cropped to 32 residues
    assert torch.equal(out['seq'][0, :n_ab], b['seq'][0, :n_ab])
    n_small = small['anchor_flag'].shape[1]
    assert bool((out['seq'][1, n_small:n_ab] == 20).all()) and not bool(out['mask'][1, n_small:n_ab].any())   # padding
    # antibody CA centroid at the origin (dataset.py:354-366)
    ca = out['atom14_gt_positions'][0, :n_ab, 1]
    assert float(ca.mean(0).abs().max()) < 1e-3
    # every kept antigen residue lies within the +-5 window of a residue within 16 A of an anchor
    assert out['antigen_origin_str_seq'][0] and len(out['antigen_origin_str_seq'][0]) <= 60


def test_cli_keeps_the_reference_flags():
    from abx_b200 import cli
    ref = ['--gpu_list', '--device', '--model', '--model_features', '--model_config', '--output_dir', '--mode', '--batch_size',
           '--num_samples', '--verbose']
    for single, extra in ((False, ['--name_idx', '--data_dir']), (True, ['--pdb_file'])):
        opts = {s for a in cli.build_parser(single)._actions for s in a.option_strings}
        assert set(ref + extra) <= opts
    a = cli.build_parser(False).parse_args(['--model', 'm', '--model_features', 'f', '--model_config', 'c', '--name_idx', 'i',
                                            '--data_dir', 'd', '--output_dir', 'o'])
    assert (a.mode, a.batch_size, a.num_samples, a.gpu_list, a.device) == ('design', 1, 100, [0], 'gpu')
