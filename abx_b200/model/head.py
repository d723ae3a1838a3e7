"""Heads of the score network (reference: abx/model/head.py).  Distogram / metric / tmscore heads only feed
the training losses; at inference their outputs are never read, so they keep their parameters (checkpoint
compatibility) but are not evaluated."""
import functools
from collections import OrderedDict

import torch
from torch import nn

from abx_b200.data import residue_tables as rt
from abx_b200.model import atom, quat_affine
from abx_b200.model.common_modules import LayerNorm, Linear, as_config, mlp
from abx_b200.model.score_network import IpaScore


def batched_select(params, indices, batch_dims=0):
    """abx/model/utils.py batched_select for the two call patterns of the heads."""
    if batch_dims == 0:
        return params[indices]
    assert batch_dims == 2
    idx = indices.long()
    trailing = params.shape[3:]
    idx = idx.reshape(idx.shape + (1,) * len(trailing)).expand(idx.shape + trailing)
    return torch.gather(params, 2, idx)


def plddt(logits):
    """abx/model/utils.py:157-171: expected lDDT bin centre x 100."""
    nb = logits.shape[-1]
    centers = torch.arange(start=0.5 / nb, end=1.0, step=1.0 / nb, device=logits.device)
    return torch.sum(torch.softmax(logits, dim=-1) * centers, dim=-1) * 100


class DistogramHead(nn.Module):
    def __init__(self, config, num_in_channel):
        super().__init__()
        c = config
        self.proj = Linear(num_in_channel + 2 * c.index_embed_size, c.num_bins, init='final')
        self.config = config

    def forward(self, headers, representations, batch):
        return None          # training-loss head, unused by the sampler


class DiffusionHead(nn.Module):
    def __init__(self, config, num_in_seq_channel, num_in_pair_channel, diffuser):
        super().__init__()
        self.ScoreNetwork = IpaScore(config, num_in_seq_channel, num_in_pair_channel, diffuser)
        self.config = config

    def forward(self, headers, representations, batch):
        return self.ScoreNetwork(representations, batch)


def _mlp(dim, hidden, out):
    return nn.Sequential(LayerNorm(dim), Linear(dim, hidden, init='relu'), nn.ReLU(), Linear(hidden, hidden, init='relu'),
                         nn.ReLU(), Linear(hidden, out, init='relu'))


class SequenceHead(nn.Module):
    def __init__(self, config, num_res=20):
        super().__init__()
        self.net = _mlp(config.num_channel, config.num_hidden_channel, num_res)
        self.config = config

    def forward(self, headers, representations, batch):
        """head.py:162-201."""
        fold = headers['folding']
        logits = mlp(self.net, fold['representations']['structure_module'])
        seq_0 = torch.argmax(logits, dim=-1)            # == argmax of the softmax (head.py:165-166)
        fixed_mask = batch['fixed_mask']
        seq_0 = seq_0 * (1 - fixed_mask) + batch['seq_t'] * fixed_mask
        rigids = fold['rigids']
        frames = atom.torsion_angles_to_frames(seq_0, (quat_affine.quat_to_rot(rigids[..., :4]), rigids[..., 4:]),
                                               fold['sidechains'][-1]['angles_sin_cos'])
        pos14 = atom.frames_and_literature_positions_to_atom14_pos(seq_0, frames)
        fold.update(final_atom14_positions=pos14,
                    final_atom_positions=batched_select(pos14, batch['residx_atom37_to_atom14'], batch_dims=2),
                    atom14_atom_exists=atom._table('restype_atom14_mask', seq_0.device)[seq_0],
                    atom37_atom_exists=atom._table('restype_atom37_mask', seq_0.device)[seq_0])
        fold['sidechains'][-1].update(atom_pos=pos14, frames=frames)
        return dict(logits=logits, seq_0=seq_0)


class PredictedLDDTHead(nn.Module):
    def __init__(self, config, bins=50):
        super().__init__()
        self.net = _mlp(config.num_channel, config.num_hidden_channel, bins)
        self.config = config

    def forward(self, headers, representations, batch):
        logits = mlp(self.net, headers['folding']['representations']['structure_module'])
        return dict(logits=logits, pLDDT=plddt(logits))


class HeaderBuilder:
    @staticmethod
    def build(config, seq_channel, pair_channel, parent, diffuser=None):
        """head.py:230-256: registers the heads on `parent` under their config names; the diffusion head is
        exposed as 'folding'.  'metric' and 'tmscore' have no parameters and only produce training metrics."""
        config = as_config(config)
        factory = OrderedDict(
            diffusion_module=functools.partial(DiffusionHead, num_in_seq_channel=seq_channel,
                                               num_in_pair_channel=pair_channel, diffuser=diffuser),
            sequence_module=SequenceHead,
            distogram=functools.partial(DistogramHead, num_in_channel=pair_channel),
            predicted_lddt=PredictedLDDTHead)
        out = []
        for name, h in factory.items():
            if name not in config:
                continue
            head = h(config=config[name])
            if isinstance(parent, nn.Module):
                parent.add_module(name, head)
            out.append(('folding' if name == 'diffusion_module' else name, head, config[name]))
        return out
