"""`IpaScore` — the structure module of the score network (reference: abx/model/score_network.py:30-196):
eight weight-shared Invariant-Point-Attention iterations over the residue frames, frame updates, torsion
angles, and the conversion of the predicted frames into SO(3)/R^3 scores.

Same constructor, parameter names and output dict as the reference.  The IPA layer runs on the sm_100a
kernels (abx_b200/model/folding.py); because the pair activations and the IPA weights are identical in all
eight iterations, the pair-bias projection is evaluated once per call instead of eight times.  Both scores
come from one launch of `abx_se3_scores` (FullDiffuser.calc_scores).
"""
import torch
from torch import nn
from torch.nn import functional as F

from abx_b200.model import quat_affine, r3
from abx_b200.model.common_modules import LayerNorm, Linear, as_config, mlp
from abx_b200.model.folding import InvariantPointAttention as IPA
from abx_b200.model.sidechain import MultiRigidSidechain


class IpaScore(nn.Module):

    def __init__(self, config, num_in_seq_channel, num_in_pair_channel, diffuser):
        super().__init__()
        config = as_config(config)
        c = config.IPA
        self.score_network_conf = config
        self._embed_conf = config.embed
        self.config = c
        self.diffuser = diffuser
        num_pair_channel = num_in_pair_channel
        self.num_in_seq_channel = self._embed_conf.index_embed_size + num_in_seq_channel
        self.num_in_pair_channel = 2 * self._embed_conf.index_embed_size + num_in_pair_channel

        self.proj_init_seq_act = Linear(self.num_in_seq_channel, c.num_channel, init='linear')
        self.proj_init_pair_act = Linear(self.num_in_pair_channel, num_pair_channel, init='linear')
        self.init_seq_layer_norm = LayerNorm(c.num_channel)
        self.init_pair_layer_norm = LayerNorm(num_pair_channel)
        self.proj_seq = Linear(c.num_channel, c.num_channel, init='linear')
        self.attention_module = IPA(c, num_pair_channel)
        self.attention_layer_norm = LayerNorm(c.num_channel)
        layers = []
        for k in range(c.num_layer_in_transition):
            last = k == c.num_layer_in_transition - 1
            layers.append(Linear(c.num_channel, c.num_channel, init='linear' if last else 'final'))
            if not last:
                layers.append(nn.ReLU())
        self.transition_module = nn.Sequential(*layers)
        self.transition_layer_norm = LayerNorm(c.num_channel)
        self.affine_update = Linear(c.num_channel, 6, init='final')
        self.sidechain_module = MultiRigidSidechain(c, num_in_seq_channel)

    def _apply_mask(self, aatype_diff, aatype_0, diff_mask):
        return diff_mask * aatype_diff + (1 - diff_mask) * aatype_0

    def forward(self, representations, batch):
        c = self.config
        seq_act, static_pair_act = representations['seq'], representations['pair']
        seq = batch['seq_t']
        node_mask = batch['mask'].type(torch.float32)
        keep = (1 - batch['fixed_mask'])[..., None]                       # 1 = residue is being designed
        init_rigids = batch['rigids_t'].type(torch.float32)              # score_network.py:90
        init_quats, init_trans = init_rigids[..., :4], init_rigids[..., 4:]
        b, n = seq.shape

        from abx_b200 import lib
        L_ = lib.load()
        dev = seq_act.device
        delta_quat, _ = quat_affine.make_identity(out_shape=(b, n), device=dev)
        init_quats = init_quats.contiguous()
        init_trans_s = (init_trans / c.position_scale).contiguous()
        fixed = batch['fixed_mask'].to(torch.int32).contiguous()
        # frame state, updated in place by abx_ipa_frame_update each iteration
        curr_quats = init_quats.clone()
        curr_trans = init_trans_s.clone()
        curr_rots = quat_affine.quat_to_rot(curr_quats).contiguous()

        seq_act = self.init_seq_layer_norm(self.proj_init_seq_act(seq_act))              # :117-120
        static_pair_act = self.init_pair_layer_norm(self.proj_init_pair_act(static_pair_act))
        initial_seq_act = seq_act
        seq_act = self.proj_seq(seq_act)
        outputs = dict(traj=[], sidechains=[])

        pair_bias = self.attention_module.pair_bias(static_pair_act)     # shared by the 8 iterations
        for fold_it in range(c.num_layer):                               # :126-163
            is_last = fold_it == c.num_layer - 1
            seq_act = self.attention_module(inputs_1d=seq_act, inputs_2d=static_pair_act, mask=node_mask,
                                            in_rigids=(curr_rots, curr_trans), pair_bias=pair_bias, residual=seq_act)
            seq_act = self.attention_layer_norm(seq_act)
            seq_act = self.transition_layer_norm(mlp(self.transition_module, seq_act, residual=seq_act))

            upd = self.affine_update(seq_act)                            # [B,N,6] = (quaternion, translation) update
            with lib.device_guard(upd):                                  # :137-149 in one launch
                lib.check(L_.abx_ipa_frame_update(lib.stream(), b, n, lib.ptr(upd), lib.ptr(init_quats), lib.ptr(init_trans_s),
                                                  lib.ptr(fixed), lib.ptr(delta_quat), lib.ptr(curr_quats), lib.ptr(curr_trans),
                                                  lib.ptr(curr_rots)))
            # the frame buffers are updated in place: the per-layer trajectory keeps copies
            outputs['traj'].append((curr_rots if is_last else curr_rots.clone(), curr_trans * c.position_scale))
            if is_last:
                outputs['sidechains'].append(self.sidechain_module(
                    seq, (curr_rots, curr_trans * c.position_scale), [seq_act, initial_seq_act], batch,
                    compute_atom_pos=False))     # atom positions are rebuilt by SequenceHead from seq_0 (head.py:181-186)

        curr_quats_ = self._apply_mask(quat_affine.quat_multiply(init_quats, delta_quat), init_quats, keep)   # :166-169
        final_trans = curr_trans * c.position_scale
        rot_score, trans_score = self.diffuser.calc_scores(init_quats, curr_quats_, init_trans, final_trans, batch['t'])
        outputs['rot_score'] = rot_score
        outputs['trans_score'] = trans_score
        outputs['representations'] = {'structure_module': seq_act}
        outputs['rigids'] = torch.cat([curr_quats_, final_trans], dim=-1)
        return outputs
