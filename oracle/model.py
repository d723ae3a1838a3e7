"""ORACLE (test infrastructure, not product code) — functional torch-CPU restatement of the
reference score network: Invariant Point Attention, the IpaScore structure module, the sequence /
pLDDT heads, the embedding + Seqformer trunk, and the recycling wrapper.

Parameters come in as a flat dict `P` with the reference's `state_dict` names, so
the same seeded weights drive the reference (when goldens are generated), this oracle and the B200
modules.  Pinned by tests/golden/{ipa,ipascore,model,sampler}.npz (outputs of the reference itself).
ESM is disabled (weights unavailable offline); heads whose outputs the sampler never reads
(distogram / metric / tmscore, head.py:26-141) are not evaluated.
"""
import math

import torch
import torch.nn.functional as F

from abx_b200.data import residue_tables as rt
from oracle import quat as Q

IPA_CONF = dict(num_layer=8, num_head=12, num_channel=256, num_scalar_qk=16, num_scalar_v=16, num_point_qk=4,
                num_point_v=8, position_scale=10.0)
MODEL_CONF = dict(num_recycle=2, seq_channel=512, pair_channel=128, max_relative_feature=32, index_embed_size=32,
                  prev_pos=dict(min_bin=3.375, max_bin=21.375, num_bins=15))

SN = 'impl.diffusion_module.ScoreNetwork.'
TR = 'impl.seqformer.'

# Working dtype of the activations.  float32 restates the reference; float64 (set_compute_dtype) gives the
# "true" value of the same arithmetic, against which the float32 noise of the reference, of this oracle and of
# the CUDA path is measured (oracle/make_golden.py gen_big, tests/test_gpu_model.py error budgets).
DTYPE = torch.float32


def set_compute_dtype(dtype):
    """torch.float32 (reference semantics) or torch.float64 (error-budget runs: pass float64 params and inputs)."""
    global DTYPE
    assert dtype in (torch.float32, torch.float64)
    DTYPE = dtype


def linear(P, name, x):
    b = P.get(name + '.bias')
    return F.linear(x, P[name + '.weight'], b)


def layer_norm(P, name, x):
    return F.layer_norm(x, (x.shape[-1],), P[name + '.weight'], P[name + '.bias'], 1e-5)


# ---------------------------------------------------------------------------------------------------
# Invariant Point Attention  (abx/model/folding.py:47-132)
# ---------------------------------------------------------------------------------------------------

def ipa_forward(P, x, z, mask, rots, trans, prefix=SN + 'attention_module.', c=IPA_CONF, return_parts=False):
    """x [B,N,256], z [B,N,N,128], mask [B,N] float, rots [B,N,3,3], trans [B,N,3] (nm) -> [B,N,256]."""
    B, N, _ = x.shape
    H, Cqk, Cv, Pqk, Pv = c['num_head'], c['num_scalar_qk'], c['num_scalar_v'], c['num_point_qk'], c['num_point_v']
    w_scalar = math.sqrt(1.0 / (3 * max(Cqk, 1)))                    # :59-66
    w_point = math.sqrt(1.0 / (3 * max(Pqk, 1) * 9.0 / 2))
    w_pair = math.sqrt(1.0 / 3)

    q_s = linear(P, prefix + 'proj_q_scalar', x).reshape(B, N, H, Cqk)                      # :69-70
    kv_s = linear(P, prefix + 'proj_kv_scalar', x).reshape(B, N, H, Cqk + Cv)               # :72-77
    k_s, v_s = kv_s[..., :Cqk], kv_s[..., Cqk:]
    logits = torch.einsum('bihc,bjhc->bhij', q_s * w_scalar, k_s)                           # :79

    # point projections: channel layout (r n) with r = xyz outermost (:81-86)
    q_p = linear(P, prefix + 'proj_q_point_local', x).reshape(B, N, 3, H * Pqk).transpose(-1, -2)
    kv_p = linear(P, prefix + 'proj_kv_point_local', x).reshape(B, N, 3, H * (Pqk + Pv)).transpose(-1, -2)
    q_g = Q.rigids_apply((rots, trans), q_p).reshape(B, N, H, Pqk, 3)                       # :89-92
    kv_g = Q.rigids_apply((rots, trans), kv_p).reshape(B, N, H, Pqk + Pv, 3)
    k_g, v_g = kv_g[..., :Pqk, :], kv_g[..., Pqk:, :]                                       # :93

    d2 = torch.sum((q_g[:, :, None] - k_g[:, None]) ** 2, dim=(-1, -2))                     # [B,i,j,H]  :95
    gamma = F.softplus(P[prefix + 'trainable_point_weights'])
    logits = logits + (-0.5 * w_point * gamma * d2).permute(0, 3, 1, 2)                     # :96-99
    logits = logits + w_pair * linear(P, prefix + 'proj_pair', z).permute(0, 3, 1, 2)       # :101-104

    m2 = (mask[:, :, None] * mask[:, None, :])[:, None]                                     # :106-109
    logits = logits.masked_fill(~m2.bool(), torch.finfo(logits.dtype).min)
    attn = torch.softmax(logits, dim=-1)                                                    # :111

    o_s = torch.einsum('bhij,bjhc->bihc', attn, v_s).reshape(B, N, H * Cv)                  # :114-116
    o_pg = torch.einsum('bhij,bjhnr->bihnr', attn, v_g).reshape(B, N, H * Pv, 3)            # :119-120
    o_pl = Q.rigids_apply(Q.invert_rigids((rots, trans)), o_pg)                             # :121
    o_pair = torch.einsum('bhij,bijc->bihc', attn, z).reshape(B, N, -1)                     # :126-127
    feats = torch.cat([o_s,
                       o_pl.transpose(-1, -2).reshape(B, N, 3 * H * Pv),                    # '(r n)'  :122
                       torch.sqrt(torch.sum(o_pl ** 2, dim=-1) + 1e-8),                     # :123
                       o_pair], dim=-1)
    out = linear(P, prefix + 'final_proj', feats)                                           # :130-132
    if return_parts:
        return out, dict(attn=attn, feats=feats, logits=logits)
    return out


# ---------------------------------------------------------------------------------------------------
# torsions -> atoms  (abx/model/sidechain.py, atom.py)
# ---------------------------------------------------------------------------------------------------

def torsion_module(P, prefix, act, init_act):
    """sidechain.py:28-53."""
    a = linear(P, prefix + 'proj_act.1', F.relu(act)) + linear(P, prefix + 'proj_init_act.1', F.relu(init_act))
    for k in range(2):
        h = linear(P, f'{prefix}blocks.{k}.net.1', F.relu(a))
        a = a + linear(P, f'{prefix}blocks.{k}.net.3', F.relu(h))
    out = linear(P, prefix + 'projection', F.relu(a))
    return out.reshape(out.shape[:-1] + (7, 2))


def torsion_angles_to_frames(aatype, rots, trans, sin_cos):
    """atom.py:9-58: 8 rigid groups per residue in the global frame."""
    m = torch.from_numpy(rt.table('restype_rigid_group_default_frame')).to(aatype.device)[aatype].to(sin_cos.dtype)           # [B,N,8,4,4]
    d_rot, d_trans = m[..., :3, :3], m[..., :3, 3]
    sin = F.pad(sin_cos[..., 0], (1, 0), value=0.)
    cos = F.pad(sin_cos[..., 1], (1, 0), value=1.)
    zeros, ones = torch.zeros_like(sin), torch.ones_like(sin)
    rx = torch.stack([ones, zeros, zeros, zeros, cos, -sin, zeros, sin, cos], dim=-1).reshape(sin.shape + (3, 3))
    f_rot, f_trans = torch.einsum('...rd,...dm->...rm', d_rot, rx), d_trans
    frames = [(f_rot[:, :, g], f_trans[:, :, g]) for g in range(8)]
    chi2 = Q.rigids_mul_rigids(frames[4], frames[5])
    chi3 = Q.rigids_mul_rigids(chi2, frames[6])
    chi4 = Q.rigids_mul_rigids(chi3, frames[7])
    to_bb = frames[:5] + [chi2, chi3, chi4]
    bb_rot = torch.stack([f[0] for f in to_bb], dim=2)
    bb_trans = torch.stack([f[1] for f in to_bb], dim=2)
    return Q.rigids_mul_rigids((rots[:, :, None].expand(-1, -1, 8, -1, -1), trans[:, :, None].expand(-1, -1, 8, -1)),
                               (bb_rot, bb_trans))


def frames_to_atom14(aatype, frames):
    """atom.py:60-76."""
    f_rot, f_trans = frames
    grp = torch.from_numpy(rt.table('restype_atom14_to_rigid_group')).to(aatype.device)[aatype].long()       # [B,N,14]
    a_rot = torch.gather(f_rot, 2, grp[..., None, None].expand(grp.shape + (3, 3)))
    a_trans = torch.gather(f_trans, 2, grp[..., None].expand(grp.shape + (3,)))
    lit = torch.from_numpy(rt.table('restype_atom14_rigid_group_positions')).to(aatype.device)[aatype].to(f_rot.dtype)
    return a_trans + torch.einsum('...rd,...d->...r', a_rot, lit)


# ---------------------------------------------------------------------------------------------------
# IpaScore  (abx/model/score_network.py:83-196)
# ---------------------------------------------------------------------------------------------------

def ipascore_forward(P, diffuser, rep_seq, rep_pair, batch, c=IPA_CONF):
    seq = batch['seq_t']
    node_mask = batch['mask'].to(DTYPE)
    fixed = batch['fixed_mask']
    init_rigids = batch['rigids_t'].to(DTYPE)                                               # :90
    init_q, init_t = init_rigids[..., :4], init_rigids[..., 4:]
    scale = c['position_scale']
    B, N = seq.shape

    delta_q = torch.cat([torch.ones(B, N, 1, dtype=DTYPE, device=seq.device), torch.zeros(B, N, 3, dtype=DTYPE, device=seq.device)], dim=-1)               # make_identity :107
    cur_q, cur_t = init_q, init_t / scale
    cur_R = Q.quat_to_rot(cur_q)

    s = layer_norm(P, SN + 'init_seq_layer_norm', linear(P, SN + 'proj_init_seq_act', rep_seq))      # :117-120
    z = layer_norm(P, SN + 'init_pair_layer_norm', linear(P, SN + 'proj_init_pair_act', rep_pair))
    s_init = s
    s = linear(P, SN + 'proj_seq', s)
    keep = (1 - fixed[..., None])
    traj = []
    for it in range(c['num_layer']):                                                        # :126-163
        s = s + ipa_forward(P, s, z, node_mask, cur_R, cur_t)
        s = layer_norm(P, SN + 'attention_layer_norm', s)
        h = F.relu(linear(P, SN + 'transition_module.0', s))
        h = F.relu(linear(P, SN + 'transition_module.2', h))
        s = s + linear(P, SN + 'transition_module.4', h)
        s = layer_norm(P, SN + 'transition_layer_norm', s)
        upd = linear(P, SN + 'affine_update', s)
        dq, dx = upd[..., :3], upd[..., 3:]
        delta_q = Q.quat_precompose_vec(delta_q, dq)
        cur_q = Q.quat_precompose_vec(cur_q, dq)
        cur_t = Q.rigids_apply((cur_R, cur_t), dx)                                          # rigids_mul_vecs :140
        cur_q = keep * cur_q + (1 - keep) * init_q                                          # :142-147
        cur_t = keep * cur_t + (1 - keep) * (init_t / scale)
        cur_R = Q.quat_to_rot(cur_q)
        traj.append((cur_R, cur_t * scale))

    # torsion angles on the last iteration (sidechain.py:64-80)
    raw = torsion_module(P, SN + 'sidechain_module.torsion_module.', s, s_init)
    ang = Q.l2_normalize(raw)
    ang = torch.where(fixed[..., None, None].bool(), batch['torsion_angles_sin_cos'], ang)

    q_hat = Q.quat_multiply(init_q, delta_q)                                                # :166-169
    q_hat = keep * q_hat + (1 - keep) * init_q
    rot_score = diffuser.calc_quat_score(init_q, q_hat, batch['t'])                         # :173-177
    trans_score = diffuser.calc_trans_score(init_t, cur_t * scale, batch['t'])              # :180-184
    rigids = torch.cat([q_hat, cur_t * scale], dim=-1)
    return dict(rot_score=rot_score, trans_score=trans_score, rigids=rigids, structure_module=s,
                angles_sin_cos=ang, traj=traj)


# ---------------------------------------------------------------------------------------------------
# heads  (abx/model/head.py:143-227)
# ---------------------------------------------------------------------------------------------------

def _mlp_head(P, prefix, act):
    h = layer_norm(P, prefix + 'net.0', act)
    h = F.relu(linear(P, prefix + 'net.1', h))
    h = F.relu(linear(P, prefix + 'net.3', h))
    return linear(P, prefix + 'net.5', h)


def sequence_head(P, fold, batch):
    """head.py:162-201: logits, argmax sequence (fixed positions keep seq_t), atom14 / atom37."""
    logits = _mlp_head(P, 'impl.sequence_module.', fold['structure_module'])
    seq_0 = torch.max(torch.softmax(logits, dim=-1), dim=-1)[1]
    fixed = batch['fixed_mask']
    seq_0 = seq_0 * (1 - fixed) + batch['seq_t'] * fixed
    rig = fold['rigids']
    frames = torsion_angles_to_frames(seq_0, Q.quat_to_rot(rig[..., :4]), rig[..., 4:], fold['angles_sin_cos'])
    atom14 = frames_to_atom14(seq_0, frames)
    idx = batch['residx_atom37_to_atom14'].long()
    atom37 = torch.gather(atom14, 2, idx[..., None].expand(idx.shape + (3,)))               # batched_select
    return dict(logits=logits, seq_0=seq_0, atom14=atom14, atom37=atom37)


def plddt_head(P, fold):
    """head.py:221-226 + utils.py:157-171."""
    logits = _mlp_head(P, 'impl.predicted_lddt.', fold['structure_module'])
    nb = logits.shape[-1]
    centers = torch.arange(start=0.5 / nb, end=1.0, step=1.0 / nb)
    return torch.sum(torch.softmax(logits, dim=-1) * centers, dim=-1) * 100


# ---------------------------------------------------------------------------------------------------
# trunk: embeddings  (abx/model/encoder.py, seqformer.py:123-226)
# ---------------------------------------------------------------------------------------------------

def pseudo_beta(atoms):
    """common_modules.py:61-83 (v2): ideal C-beta from N, CA, C (indices 0,1,2 in atom14 and atom37)."""
    n, ca, cc = atoms[..., 0, :], atoms[..., 1, :], atoms[..., 2, :]
    b, c = ca - n, cc - ca
    a = torch.cross(b, c, dim=-1)
    return -0.58273431 * a + 0.56802827 * b - 0.54067466 * c + ca


def dgram_bins(pos, num_bins, min_bin, max_bin):
    """common_modules.py:107-120."""
    sq = torch.linspace(min_bin, max_bin, steps=num_bins - 1) ** 2
    d2 = torch.sum((pos[:, :, None] - pos[:, None]) ** 2, dim=-1, keepdim=True)
    return torch.sum(d2 > sq, dim=-1).long()


def _seq(P, prefix, x, layers):
    """nn.Sequential of Linear / ReLU: `layers` = indices of the Linear members."""
    for n, k in enumerate(layers):
        x = linear(P, f'{prefix}{k}', x)
        if n + 1 < len(layers):
            x = F.relu(x)
    return x


def residue_embedding(P, batch):
    """encoder.py:149-175."""
    p = TR + 'encode_residue_emb.'
    mask = torch.logical_and(batch['mask'], batch['fixed_mask'])
    B, L = mask.shape
    aa = P[p + 'aatype_embed.weight'][batch['seq_t'].long()] * mask[:, :, None]
    cdr = P[p + 'cdr_embed.weight'][batch['cdr_def']]
    geo = torch.cat([batch['atom14_gt_positions'].reshape(B, L, -1), batch['torsion_angles_sin_cos'].reshape(B, L, -1)], -1)
    coord = _seq(P, p + 'coordinate_embed.', geo, [0, 2])
    x = torch.cat([aa, batch['chain_id'][..., None], batch['residx'][..., None], cdr, coord], dim=-1)
    return _seq(P, p + 'mlp.', x, [0, 2, 4, 6]) * mask[:, :, None]


def pair_embedding(P, batch, prev_pos_conf):
    """encoder.py:211-269."""
    p = TR + 'encode_pair_emb.'
    mask = torch.logical_and(batch['mask'], batch['fixed_mask'])
    mask_pair = mask[:, :, None] * mask[:, None, :]
    B, L = mask.shape
    aa, chain, residx = batch['seq_t'], batch['chain_id'], batch['residx']
    coords, ca_mask = batch['atom14_gt_positions'], batch['atom14_gt_exists'][..., 1]
    aa_pair = (aa[:, :, None] * 23 + aa[:, None, :]).long()
    f_aa = P[p + 'aa_pair_embed.weight'][aa_pair]
    same = chain[:, :, None] == chain[:, None, :]
    rel = torch.clamp(residx[:, :, None] - residx[:, None, :], min=-32, max=32)
    f_rel = P[p + 'relpos_embed.weight'][rel + 32] * same[..., None]
    dist = (torch.linalg.norm(coords[:, :, None, :, None] - coords[:, None, :, None, :], dim=-1, ord=2) / 10).reshape(B, L, L, -1)
    coef = F.softplus(P[p + 'aapair_to_distcoef.weight'][aa_pair])
    gauss = torch.exp(-1 * coef * dist ** 2)
    m_atom = ca_mask[:, :, None, None] * ca_mask[:, None, :, None]
    f_dist = F.relu(_seq(P, p + 'distance_embed.', gauss * m_atom, [0, 2]))
    f_dg = P[p + 'dgram_embed.weight'][dgram_bins(pseudo_beta(coords), **prev_pos_conf)]
    x = torch.cat([f_aa, f_rel, f_dist, f_dg], dim=-1)
    return _seq(P, p + 'out_mlp.', x, [0, 2, 4]) * mask_pair[..., None]


def timestep_embedding(t, dim, max_positions=10000):
    """seqformer.py:49-66."""
    t = t * max_positions
    half = dim // 2
    e = math.log(max_positions) / (half - 1)
    e = torch.exp(torch.arange(half, dtype=DTYPE) * -e)
    e = t.to(DTYPE)[:, None] * e[None, :]
    return torch.cat([torch.sin(e), torch.cos(e)], dim=1)


def embed_inputs(P, batch, mc=MODEL_CONF):
    """seqformer.py:170-224 (ESM branch disabled): returns seq [B,N,544], pair [B,N,N,192]."""
    seq_t, residx = batch['seq_t'], batch['residx']
    n_ab = batch['anchor_flag'].shape[1]
    mrf = mc['max_relative_feature']
    B, N = seq_t.shape

    def relpos(pos):
        off = pos[:, None, :] - pos[:, :, None]
        return torch.clip(off + mrf, min=0, max=2 * mrf) + 1

    ab_seq = P[TR + 'proj_aa_type.weight'][seq_t[:, :n_ab].long()]
    ab_pair = P[TR + 'proj_rel_pos.weight'][relpos(residx[:, :n_ab])]
    ag_embed = P[TR + 'proj_aa_type.weight'][batch['seq'][:, n_ab:]]
    ag_seq = layer_norm(P, TR + 'aa_proj.0', ag_embed)
    ag_seq = linear(P, TR + 'aa_proj.3', F.relu(linear(P, TR + 'aa_proj.1', ag_seq)))
    ag_pair = P[TR + 'proj_rel_pos.weight'][relpos(batch['residx'][:, n_ab:])]

    seq_act = torch.cat([ab_seq, ag_seq], dim=1)
    pair_act = torch.zeros(B, N, N, ab_pair.shape[-1], dtype=ab_pair.dtype)                                    # pair_concat :24-45
    pair_act[:, :n_ab, :n_ab] = ab_pair
    pair_act[:, n_ab:, n_ab:] = ag_pair
    seq_act = seq_act + residue_embedding(P, batch)
    pair_act = pair_act + pair_embedding(P, batch, mc['prev_pos'])

    te = timestep_embedding(batch['t'], mc['index_embed_size'])[:, None, :].expand(B, N, -1)   # Embedder :93-119
    seq_act = torch.cat([seq_act, te], dim=-1).to(DTYPE)
    pair_act = torch.cat([pair_act, te[:, :, None, :].expand(B, N, N, -1), te[:, None, :, :].expand(B, N, N, -1)], dim=-1).to(DTYPE)

    seq_act = seq_act + layer_norm(P, TR + 'prev_seq_norm', batch['prev_seq'])              # :213-217
    pair_act = pair_act + layer_norm(P, TR + 'prev_pair_norm', batch['prev_pair'])
    pair_act = pair_act + P[TR + 'proj_prev_pos.weight'][batch['prev_pos']]                 # :219-220
    return seq_act, pair_act


# ---------------------------------------------------------------------------------------------------
# trunk: one Seqformer block  (abx/model/seqformer.py:228-606)
# ---------------------------------------------------------------------------------------------------

def _attention(P, prefix, q_data, k_data, bias, k_mask, num_head, split_first):
    """seqformer.py:228-301 (gating on, no inception kernels)."""
    if split_first:
        q, k, v = (linear(P, prefix + n, d) for n, d in (('proj_q', q_data), ('proj_k', k_data), ('proj_v', k_data)))
        q, k, v = (x.reshape(x.shape[:-1] + (num_head, -1)).transpose(-2, -3) for x in (q, k, v))      # b s h l d
    else:
        t = linear(P, prefix + 'proj_in', q_data)
        t = t.reshape(t.shape[:-1] + (num_head, -1)).transpose(-2, -3)
        q, k, v = torch.chunk(t, 3, dim=-1)
    key_dim = q.shape[-1]
    q = q * key_dim ** (-0.5)
    logits = torch.einsum('...hqd,...hkd->...hqk', q, k)
    logits = logits + bias[:, None]
    logits = logits.masked_fill(~k_mask[:, :, None, None, :].bool(), torch.finfo(logits.dtype).min)
    w = torch.softmax(logits, dim=-1)
    o = torch.einsum('bshqk,bshkd->bshqd', w, v)
    o = o.transpose(-2, -3).reshape(q_data.shape[:-1] + (-1,))
    o = o * torch.sigmoid(linear(P, prefix + 'gate', q_data))
    return linear(P, prefix + 'proj_out', o)


def _transition(P, prefix, x):
    """seqformer.py:358-376."""
    h = layer_norm(P, prefix + 'transition.0', x)
    return linear(P, prefix + 'transition.3', F.relu(linear(P, prefix + 'transition.1', h)))


def _triangle_mult(P, prefix, act, mask, outgoing):
    """seqformer.py:413-504."""
    pm = mask[:, :, None, None] * mask[:, None, :, None]
    act = layer_norm(P, prefix + 'norm', act)
    left = pm * linear(P, prefix + 'left_proj', act)
    right = pm * linear(P, prefix + 'right_proj', act)
    left = left * torch.sigmoid(linear(P, prefix + 'left_gate', act))
    right = right * torch.sigmoid(linear(P, prefix + 'right_gate', act))
    if outgoing:
        out = torch.einsum('bikc,bjkc->bijc', left, right)
    else:
        out = torch.einsum('bkic,bkjc->bijc', left, right)
    out = linear(P, prefix + 'proj_out', layer_norm(P, prefix + 'final_norm', out))
    return out * torch.sigmoid(linear(P, prefix + 'final_gate', act))


def _triangle_attn(P, prefix, pair, mask, per_column):
    """seqformer.py:506-550."""
    if per_column:
        pair = pair.transpose(1, 2)
    pair = layer_norm(P, prefix + 'norm', pair)
    bias = linear(P, prefix + 'proj_pair', pair).permute(0, 3, 1, 2)
    out = _attention(P, prefix + 'attn.', pair, pair, bias, mask[:, None, :], 4, True)
    return out.transpose(1, 2) if per_column else out


def seqformer_block(P, seq, pair, mask):
    """seqformer.py:569-606 (inference: no dropout)."""
    p = TR + 'seqformer.blocks.0.'
    # row attention with pair bias :303-356
    s = layer_norm(P, p + 'seq_attn.seq_norm', seq)
    bias = linear(P, p + 'seq_attn.proj_pair', layer_norm(P, p + 'seq_attn.pair_norm', pair)).permute(0, 3, 1, 2)
    seq = seq + _attention(P, p + 'seq_attn.attn.', s[:, None], None, bias, mask[:, None, :], 32, False)[:, 0]
    seq = seq + _transition(P, p + 'seq_transition.', seq)
    # outer product mean :378-411
    m = mask[:, :, None].to(seq.dtype)
    a = layer_norm(P, p + 'outer_product_mean.norm', seq)
    left = m * linear(P, p + 'outer_product_mean.left_proj', a)
    right = m * linear(P, p + 'outer_product_mean.right_proj', a)
    opm = torch.cat([left[:, None, :, :] * right[:, :, None, :], left[:, None, :, :] - right[:, :, None, :]], dim=-1)
    pair = pair + linear(P, p + 'outer_product_mean.out_proj', opm)
    mf = mask.to(pair.dtype)
    pair = pair + _triangle_mult(P, p + 'triangle_multiplication_outgoing.', pair, mf, True)
    pair = pair + _triangle_mult(P, p + 'triangle_multiplication_incoming.', pair, mf, False)
    pair = pair + _triangle_attn(P, p + 'triangle_attention_starting_node.', pair, mask, False)
    pair = pair + _triangle_attn(P, p + 'triangle_attention_ending_node.', pair, mask, True)
    pair = pair + _transition(P, p + 'pair_transition.', pair)
    return seq, pair


# ---------------------------------------------------------------------------------------------------
# ScoreNetwork  (abx/model/abx.py)
# ---------------------------------------------------------------------------------------------------

def iteration(P, diffuser, batch, with_plddt):
    """abx.py:42-63."""
    seq, pair = embed_inputs(P, batch)
    seq, pair = seqformer_block(P, seq, pair, batch['mask'])
    fold = ipascore_forward(P, diffuser, seq, pair, batch)
    sh = sequence_head(P, fold, batch)
    out = dict(rep_seq=seq, rep_pair=pair, **fold, **sh)
    if with_plddt:
        out['pLDDT'] = plddt_head(P, fold)
    return out


def get_prev(batch, out, mc=MODEL_CONF):
    """abx.py:17-26."""
    return dict(prev_pos=dgram_bins(pseudo_beta(out['atom37']), **mc['prev_pos']),
                prev_seq=out['rep_seq'], prev_pair=out['rep_pair'])


def score_network(P, diffuser, batch, mc=MODEL_CONF):
    """abx.py:75-104.  Mutates `batch` exactly as the reference does: prev_* and — the parity-critical
    quirk — seq_t <- the recycle's predicted seq_0 (:97-98)."""
    B, N = batch['seq'].shape
    if 'prev_seq' not in batch:
        dev = batch['seq_t'].device
        batch.update(prev_pos=torch.zeros(B, N, N, dtype=torch.int64, device=dev), prev_seq=torch.zeros(B, N, 544, dtype=DTYPE, device=dev),
                     prev_pair=torch.zeros(B, N, N, 192, dtype=DTYPE, device=dev))
    with torch.no_grad():
        for _ in range(mc['num_recycle']):
            out = iteration(P, diffuser, batch, with_plddt=False)
            prev = get_prev(batch, out, mc)
            batch['seq_t'] = out['seq_0']
            batch.update(prev)
        return iteration(P, diffuser, batch, with_plddt=True)
