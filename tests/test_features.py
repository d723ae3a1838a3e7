"""CPU checks of the host-side feature pipeline against the batch the REFERENCE's FeatureBuilder produced
for the same synthetic complex (stored inside tests/golden/model.npz)."""
import torch

from abx_b200.data.synthetic import small_complex
from abx_b200.model import features as F_, r3
from tests.util import golden, maxabs


def test_static_features_match_reference_pipeline():
    g = golden('model')
    batch = small_complex(n_antigen=9, seed=0, batch_size=2)
    for name in ('make_restype_atom_constants', 'make_gt_frames', 'make_torsion_angles'):
        batch = F_._feats_fn[name](is_training=False)(batch)
    assert torch.equal(batch['residx_atom37_to_atom14'].int(), g['batch_residx_atom37_to_atom14'].int())
    assert maxabs(batch['rigidgroups_gt_frames'][0], g['batch_gt_frame_rots']) < 1e-5
    assert maxabs(batch['rigidgroups_gt_frames'][1], g['batch_gt_frame_trans']) < 1e-5
    assert maxabs(batch['torsion_angles_sin_cos'], g['batch_torsion_angles_sin_cos']) < 1e-5
    rig0 = r3.rigids_to_tensor7((batch['rigidgroups_gt_frames'][0][:, :, 0], batch['rigidgroups_gt_frames'][1][:, :, 0]))
    assert maxabs(rig0, g['batch_rigids_0']) < 1e-5
    diffused, _ = F_.design_mask(batch, 'H3')
    assert torch.equal(1 - diffused, g['batch_fixed_mask'].int())
    assert int(diffused.sum()) == 2 * 5            # 6-residue H3 of the small complex, last residue stays fixed (sic)
    every, _ = F_.design_mask(batch, 'cdrs')
    assert int(every[0].sum()) > int(diffused[0].sum())


def test_feature_builder_registry_reads_reference_config():
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    feats = json.load(open(os.path.join(root, 'abx_b200', 'config', 'config_data_feature.json')))
    assert all(name in F_._feats_fn for name, _ in feats)
