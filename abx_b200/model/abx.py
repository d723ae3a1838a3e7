"""`ScoreNetwork` — recycling wrapper of the score network (reference: abx/model/abx.py)."""
import torch
from torch import nn

from abx_b200.model.common_modules import as_config, dgram_from_positions, pseudo_beta_fn_v2
from abx_b200.model.head import HeaderBuilder
from abx_b200.model.seqformer import EmbeddingAndSeqformer


def get_prev(batch, value, config):
    """abx.py:17-26: self-conditioning features for the next call."""
    config = as_config(config)
    cb = pseudo_beta_fn_v2(batch['seq'], value['heads']['folding']['final_atom_positions'])
    return {'prev_pos': dgram_from_positions(cb, **config.embeddings_and_seqformer.prev_pos).detach(),
            'prev_seq': value['representations']['seq'].detach(),
            'prev_pair': value['representations']['pair'].detach()}


class ScoreNetworkIteration(nn.Module):
    def __init__(self, model_conf, diffuser):
        super().__init__()
        self._model_conf = as_config(model_conf)
        es = self._model_conf.embeddings_and_seqformer
        self.seqformer = EmbeddingAndSeqformer(es)
        self.diffuser = diffuser
        self.heads = HeaderBuilder.build(self._model_conf.heads, seq_channel=es.seq_channel, pair_channel=es.pair_channel,
                                         parent=self, diffuser=diffuser)

    def forward(self, batch, compute_loss=False):
        """abx.py:42-63.  `compute_loss` selects the final pass, which also evaluates the pLDDT head."""
        seq_act, pair_act = self.seqformer(batch)
        representations = {'pair': pair_act, 'seq': seq_act}
        ret = {'representations': representations, 'heads': {}}
        for name, module, _ in self.heads:
            if compute_loss or name in ('folding', 'sequence_module'):
                value = module(ret['heads'], representations, batch)
                if value is not None:
                    ret['heads'][name] = value
        return ret


class ScoreNetwork(nn.Module):
    def __init__(self, model_conf, diffuser):
        super().__init__()
        self._model_conf = as_config(model_conf)
        es = self._model_conf.embeddings_and_seqformer
        self.num_in_seq_channel, self.num_in_pair_channel, self.index_embed_size = es.seq_channel, es.pair_channel, es.index_embed_size
        self.impl = ScoreNetworkIteration(self._model_conf, diffuser)

    def forward(self, input_feats, compute_loss=True):
        """abx.py:75-104.  Mutates `input_feats` exactly as the reference does: prev_* and — the
        parity-critical quirk — seq_t <- the recycle's predicted seq_0 (:97-98)."""
        B, N = input_feats['seq'].shape[:2]
        device = input_feats['seq'].device
        if 'prev_seq' not in input_feats:
            input_feats.update(
                prev_pos=torch.zeros([B, N, N], device=device, dtype=torch.int64),
                prev_seq=torch.zeros([B, N, self.num_in_seq_channel + self.index_embed_size], device=device),
                prev_pair=torch.zeros([B, N, N, self.num_in_pair_channel + 2 * self.index_embed_size], device=device))
        with torch.no_grad():
            input_feats.update(is_recycling=True)
            for _ in range(self._model_conf.num_recycle):
                ret = self.impl(input_feats, compute_loss=False)
                prev = get_prev(input_feats, ret, self._model_conf)
                if 'sequence_module' in ret['heads']:
                    input_feats.update(seq_t=ret['heads']['sequence_module']['seq_0'])
                input_feats.update(prev)
            input_feats.update(is_recycling=False)
            return self.impl(input_feats, compute_loss=compute_loss)
