"""Structure-conditioned input encoders of the trunk (reference: abx/model/encoder.py:123-269).

Both encoders only see the FIXED residues (everything is masked by mask & fixed_mask), whose sequence,
coordinates and torsions never change during sampling, so their outputs are invariant across the reverse
steps and across the samples of one complex: the trunk evaluates them once per complex
(EmbeddingAndSeqformer.static_embeddings) instead of 303 times per sample.  ESM2 conditioning
(encoder.py:21-121) needs the un-vendored fair-esm package and 3B-parameter weights: not supported.
"""
import torch
from torch import nn
from torch.nn import functional as F

from abx_b200.model.common_modules import Linear, dgram_from_positions, pseudo_beta_fn_v2

RESTYPE_NUM, NUM_AB_REGIONS = 20, 14


class ResidueEmbedding(nn.Module):

    def __init__(self, config):
        super().__init__()
        d = config.seq_channel
        self.max_aa_types = RESTYPE_NUM
        self.aatype_embed = nn.Embedding(self.max_aa_types + 3, d)
        self.cdr_embed = nn.Embedding(NUM_AB_REGIONS + 1, d)
        self.coordinate_embed = nn.Sequential(Linear(14 * 3 + 7 * 2, d), nn.ReLU(), Linear(d, d))
        self.mlp = nn.Sequential(Linear(d * 3 + 2, d * 2), nn.ReLU(), Linear(d * 2, d), nn.ReLU(), Linear(d, d),
                                 nn.ReLU(), Linear(d, d))

    def forward(self, batch):
        """encoder.py:149-175."""
        mask = torch.logical_and(batch['mask'], batch['fixed_mask'])
        B, L = mask.shape
        aa_feat = self.aatype_embed(batch['seq_t'].long()) * mask[:, :, None]
        cdr_feat = self.cdr_embed(batch['cdr_def'])
        geo = torch.cat([batch['atom14_gt_positions'].reshape(B, L, -1), batch['torsion_angles_sin_cos'].reshape(B, L, -1)], -1)
        x = torch.cat([aa_feat, batch['chain_id'][..., None], batch['residx'][..., None], cdr_feat,
                       self.coordinate_embed(geo)], dim=-1)
        return self.mlp(x) * mask[:, :, None]


class PairEmbedding(nn.Module):

    def __init__(self, config):
        super().__init__()
        d = config.pair_channel
        self.max_aa_types = RESTYPE_NUM + 3
        self.max_relpos = 32
        self.aa_pair_embed = nn.Embedding(self.max_aa_types ** 2, d)
        self.relpos_embed = nn.Embedding(2 * self.max_relpos + 1, d)
        self.aapair_to_distcoef = nn.Embedding(self.max_aa_types ** 2, 14 * 14)
        self.distance_embed = nn.Sequential(Linear(14 * 14, d), nn.ReLU(), Linear(d, d), nn.ReLU())
        self.dgram_embed = nn.Embedding(config.prev_pos.num_bins, d)
        self.prev_pos = config.prev_pos
        self.out_mlp = nn.Sequential(Linear(d * 4, d), nn.ReLU(), Linear(d, d), nn.ReLU(), Linear(d, d))

    def forward(self, batch):
        """encoder.py:211-269."""
        mask = torch.logical_and(batch['mask'], batch['fixed_mask'])
        mask_pair = mask[:, :, None] * mask[:, None, :]
        B, L = mask.shape
        aa, chain, residx = batch['seq_t'], batch['chain_id'], batch['residx']
        coords, ca_mask = batch['atom14_gt_positions'], batch['atom14_gt_exists'][..., 1]
        aa_pair = (aa[:, :, None] * self.max_aa_types + aa[:, None, :]).long()
        f_aa = self.aa_pair_embed(aa_pair)
        same = chain[:, :, None] == chain[:, None, :]
        rel = torch.clamp(residx[:, :, None] - residx[:, None, :], min=-self.max_relpos, max=self.max_relpos)
        f_rel = self.relpos_embed(rel + self.max_relpos) * same[..., None]
        dist = (torch.linalg.norm(coords[:, :, None, :, None] - coords[:, None, :, None, :], dim=-1, ord=2) / 10
                ).reshape(B, L, L, -1)
        coef = F.softplus(self.aapair_to_distcoef(aa_pair))
        gauss = torch.exp(-1 * coef * dist ** 2)
        m_atom = ca_mask[:, :, None, None] * ca_mask[:, None, :, None]
        f_dist = self.distance_embed(gauss * m_atom)
        cb = pseudo_beta_fn_v2(aa, coords)
        f_dg = self.dgram_embed(dgram_from_positions(cb, **self.prev_pos))
        return self.out_mlp(torch.cat([f_aa, f_rel, f_dist, f_dg], dim=-1)) * mask_pair[..., None]
