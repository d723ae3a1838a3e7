"""Torsion-angle head of the structure module (reference: abx/model/sidechain.py)."""
import torch
from torch import nn
from torch.nn import functional as F

from abx_b200.model import atom
from abx_b200.model.common_modules import Linear, mlp
from abx_b200.model.quat_affine import l2_normalize


class ResNetBlock(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.Sequential(nn.ReLU(), Linear(dim, dim, init='relu'), nn.ReLU(), Linear(dim, dim, init='final'))

    def forward(self, act):
        return mlp(self.net, act, residual=act)


class TorsionModule(nn.Module):
    def __init__(self, config, num_in_channel, num_in_initial_channel):
        super().__init__()
        c = config
        self.proj_act = nn.Sequential(nn.ReLU(), Linear(num_in_channel, c.num_channel, init='linear'))
        self.proj_init_act = nn.Sequential(nn.ReLU(), Linear(num_in_initial_channel, c.num_channel, init='linear'))
        self.blocks = nn.Sequential(*[ResNetBlock(c.num_channel) for _ in range(c.num_residual_block)])
        self.projection = Linear(c.num_channel, 7 * 2, init='linear')

    def forward(self, act, init_act):
        act = self.blocks(self.proj_act(act) + self.proj_init_act(init_act))
        out = self.projection(F.relu(act))
        return out.reshape(out.shape[:-1] + (7, 2))


class MultiRigidSidechain(nn.Module):
    def __init__(self, config, num_in_seq_channel):
        super().__init__()
        self.torsion_module = TorsionModule(config.torsion, config.num_channel, config.num_channel)
        self.config = config

    def forward(self, seq, backb_to_global, representations_list, batch, compute_atom_pos=False):
        """sidechain.py:64-91: fixed residues keep their ground-truth torsions."""
        assert len(representations_list) == 2
        raw = self.torsion_module(*representations_list)
        fixed = batch['fixed_mask'][..., None, None].bool()
        gt = batch['torsion_angles_sin_cos']
        outputs = {'angles_sin_cos': torch.where(fixed, gt, l2_normalize(raw, dim=-1)),
                   'unnormalized_angles_sin_cos': torch.where(fixed, gt, raw)}
        if compute_atom_pos:
            frames = atom.torsion_angles_to_frames(seq, backb_to_global, outputs['angles_sin_cos'])
            outputs.update(atom_pos=atom.frames_and_literature_positions_to_atom14_pos(seq, frames), frames=frames)
        return outputs
