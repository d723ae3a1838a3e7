"""Trunk of the score network: input embeddings + one Seqformer block (reference: abx/model/seqformer.py).

Same module tree / parameter names as the reference (checkpoints load with strict=True).  The block runs on
the sm_100a kernels behind the C ABI (abx_b200/ops.py): every dense layer on the tcgen05 3xTF32 GEMM with gates,
masks, residuals and output rearranges in its epilogue, LayerNorm and the triangle-attention core (logits never
materialised) on their own kernels, the triangle multiplication as a GLU-GEMM plus a batched product.  The
step-invariant part of the embeddings is computed once per complex (`static_embeddings` / `cache_static`).
"""
import math

import torch
from torch import nn
from torch.nn import functional as F

from abx_b200.model.common_modules import LayerNorm, Linear, as_config, mlp
from abx_b200.model.encoder import PairEmbedding, ResidueEmbedding

RESTYPE_NUM, NUM_AB_REGIONS = 20, 14


def get_timestep_embedding(timesteps, embedding_dim, max_positions=10000):
    """seqformer.py:49-66."""
    assert timesteps.dim() == 1 and embedding_dim % 2 == 0
    timesteps = timesteps * max_positions
    half = embedding_dim // 2
    freq = torch.exp(torch.arange(half, dtype=torch.float32, device=timesteps.device) * -(math.log(max_positions) / (half - 1)))
    arg = timesteps.float()[:, None] * freq[None, :]
    return torch.cat([torch.sin(arg), torch.cos(arg)], dim=1)


class Attention(nn.Module):
    """seqformer.py:228-301 (gating on, no inception kernels)."""

    def __init__(self, input_dim, key_dim, value_dim, gating=True, num_head=4, split_first=True):
        super().__init__()
        assert key_dim % num_head == 0 and value_dim % num_head == 0 and gating
        self.key_dim, self.value_dim, self.num_head, self.split_first = key_dim, value_dim, num_head, split_first
        if split_first:
            self.proj_q = Linear(input_dim, key_dim, init='attn', bias=False)
            self.proj_k = Linear(input_dim, key_dim, init='attn', bias=False)
            self.proj_v = Linear(input_dim, value_dim, init='attn', bias=False)
        else:
            assert key_dim == value_dim
            self.proj_in = Linear(input_dim, key_dim * 3, init='attn', bias=False)
        self.gate = Linear(input_dim, value_dim, init='gate')
        self.proj_out = Linear(value_dim, input_dim, init='final')

    def _qkv_weight(self):
        """[Wq; Wk; Wv] as one [3*dim, C] operand so self-attention projects q, k, v in a single GEMM."""
        ws = (self.proj_q.weight, self.proj_k.weight, self.proj_v.weight)
        key = tuple((w.data_ptr(), w._version) for w in ws)
        if getattr(self, '_qkv_cache', (None,))[0] != key:
            self._qkv_cache = (key, torch.cat([w.detach() for w in ws], dim=0).contiguous())
        return self._qkv_cache[1]

    def _qkvg_weight(self):
        """[Wq; Wk; Wv; Wgate] and the matching bias (zeros for q, k, v) for the fused pair-attention path."""
        ws = (self.proj_q.weight, self.proj_k.weight, self.proj_v.weight, self.gate.weight, self.gate.bias)
        key = tuple((w.data_ptr(), w._version) for w in ws)
        if getattr(self, '_qkvg_cache', (None,))[0] != key:
            with torch.no_grad():
                w = torch.cat([t.detach() for t in ws[:4]], dim=0).contiguous()
                b = torch.cat([torch.zeros(w.shape[0] - ws[4].shape[0], device=w.device, dtype=w.dtype), ws[4].detach()])
            self._qkvg_cache = (key, w, b.contiguous())
        return self._qkvg_cache[1], self._qkvg_cache[2]

    def forward(self, q_data, k_data=None, bias=None, k_mask=None, residual=None):
        """q_data [B,S,L,C]; bias [B,H,L,L] (shared over S); k_mask [B,S|1,L] bool.
        `residual` [B,S,L,C] is added in the output projection's epilogue."""
        from abx_b200 import ops
        H = self.num_head
        if self.split_first:
            if k_data is None or k_data is q_data:
                t = ops.linear(q_data, self._qkv_weight())
                t = t.reshape(t.shape[:-1] + (3, H, -1))
                q, k, v = (t[..., i, :, :].transpose(-2, -3) for i in range(3))                       # b s h l d
            else:
                q, k, v = self.proj_q(q_data), self.proj_k(k_data), self.proj_v(k_data)
                q, k, v = (x.reshape(x.shape[:-1] + (H, -1)).transpose(-2, -3) for x in (q, k, v))
        else:
            t = self.proj_in(q_data)
            t = t.reshape(t.shape[:-1] + (H, -1)).transpose(-2, -3)
            q, k, v = torch.chunk(t, 3, dim=-1)
        # logits = (q / sqrt(d)) k^T + bias, keys masked to finfo.min, softmax, weights @ v  (seqformer.py:283-301).
        # The key mask is folded into the (S-times smaller) bias once: min + q.k == min in fp32, exactly what
        # masked_fill produces, and it keeps the big [B,S,H,L,L] tensor down to matmul / add_ / softmax_ / matmul.
        add = bias[:, None]                                                                           # b 1 h q k
        if k_mask is not None:
            add = add.masked_fill(~k_mask[:, :, None, None, :].bool(), torch.finfo(q.dtype).min)
        logits = torch.matmul(q * (q.shape[-1] ** -0.5), k.transpose(-1, -2))
        logits += add
        o = torch.matmul(torch.softmax(logits, dim=-1), v)
        o = o.transpose(-2, -3).reshape(q_data.shape[:-1] + (-1,))
        gated = self.gate(q_data, act='sigmoid_mul', gate=o)                  # sigmoid(gate(q)) * o in the epilogue
        return self.proj_out(gated, residual=residual)


    def forward_pair(self, x, bias, key_mask, residual=None, transpose_n=0):
        """Self-attention over the rows of a [B,S,L,C] pair tensor on the fused kernels: one q|k|v GEMM, the
        attention core without materialised logits (abx_pair_attention), gate and residual in GEMM epilogues.
        `transpose_n`: x is the 'b j i c' view of the pair tensor; output / residual are 'b i j c'."""
        from abx_b200 import ops
        w, b = self._qkvg_weight()
        qkvg = ops.linear(x, w, b)                                            # q | k | v | gate pre-activation in one GEMM
        gated = ops.pair_attention(qkvg, bias, key_mask, self.num_head, gated=True)   # sigmoid(gate) * attention
        return self.proj_out(gated, residual=residual, transpose_n=transpose_n)


class SeqAttentionWithPairBias(nn.Module):
    def __init__(self, config, num_in_seq_channel, num_in_pair_channel):
        super().__init__()
        self.seq_norm = LayerNorm(num_in_seq_channel)
        self.pair_norm = LayerNorm(num_in_pair_channel)
        self.proj_pair = Linear(num_in_pair_channel, config.num_head, init='linear', bias=False)
        self.attn = Attention(num_in_seq_channel, num_in_seq_channel, num_in_seq_channel, num_head=config.num_head,
                              split_first=False)
        self.config = config

    def forward(self, seq_act, pair_act, mask, residual=None):
        s = self.seq_norm(seq_act)
        bias = self.proj_pair(self.pair_norm(pair_act)).permute(0, 3, 1, 2)
        res = residual[:, None] if residual is not None else None
        return self.attn(s[:, None], bias=bias, k_mask=mask[:, None, :], residual=res)[:, 0]


class Transition(nn.Module):
    def __init__(self, config, num_in_channel):
        super().__init__()
        inter = num_in_channel * config.num_intermediate_factor
        self.transition = nn.Sequential(LayerNorm(num_in_channel), Linear(num_in_channel, inter, init='linear'), nn.ReLU(),
                                        Linear(inter, num_in_channel, init='final'))

    def forward(self, act, mask=None, residual=None):
        """LN -> Linear+ReLU -> Linear (+ residual), the ReLU and the residual in the GEMM epilogues."""
        return mlp(self.transition, act, residual=residual)


class OuterProductMean(nn.Module):
    def __init__(self, config, num_in_channel, num_out_channel):
        super().__init__()
        c = config
        self.norm = LayerNorm(num_in_channel)
        self.left_proj = Linear(num_in_channel, c.num_outer_channel, init='linear')
        self.right_proj = Linear(num_in_channel, c.num_outer_channel, init='linear')
        self.out_proj = Linear(2 * c.num_outer_channel, num_out_channel, init='final')

    def forward(self, act, mask, residual=None):
        """seqformer.py:378-411: concat(left_j * right_i, left_j - right_i) -> out_proj (+ residual)."""
        m = mask.to(act.dtype)
        a = self.norm(act)
        left, right = self.left_proj(a, row_scale=m), self.right_proj(a, row_scale=m)
        from abx_b200 import ops
        return self.out_proj(ops.outer_product(left, right), residual=residual)


class TriangleMultiplication(nn.Module):
    def __init__(self, config, num_in_channel):
        super().__init__()
        c = config
        inter = c.num_intermediate_channel
        self.norm = LayerNorm(num_in_channel)
        self.left_proj = Linear(num_in_channel, inter, init='linear')
        self.right_proj = Linear(num_in_channel, inter, init='linear')
        self.final_norm = LayerNorm(inter)
        self.left_gate = Linear(num_in_channel, inter, init='gate')
        self.right_gate = Linear(num_in_channel, inter, init='gate')
        self.final_gate = Linear(num_in_channel, num_in_channel, init='gate')
        self.proj_out = Linear(inter, num_in_channel, init='final')
        self.outgoing = c.orientation == 'per_row'
        self.config = c

    def _glu_weight(self):
        """Rows of left_proj / left_gate / right_proj / right_gate interleaved in blocks of 64 so that every
        128-column GEMM tile holds 64 projection columns followed by their 64 gate columns (GLU epilogue)."""
        ps = (self.left_proj, self.left_gate, self.right_proj, self.right_gate)
        key = tuple((p.weight.data_ptr(), p.weight._version, p.bias._version) for p in ps)
        if getattr(self, '_glu_cache', (None,))[0] != key:
            with torch.no_grad():
                ws, bs = [], []
                for proj, gate in ((ps[0], ps[1]), (ps[2], ps[3])):
                    for c in range(0, proj.weight.shape[0], 64):
                        ws += [proj.weight[c:c + 64], gate.weight[c:c + 64]]
                        bs += [proj.bias[c:c + 64], gate.bias[c:c + 64]]
                self._glu_cache = (key, torch.cat(ws, 0).contiguous(), torch.cat(bs, 0).contiguous())
        return self._glu_cache[1], self._glu_cache[2]

    def forward(self, act, mask, residual=None):
        """seqformer.py:413-504.  Gates, pair mask and the residual ride in GEMM epilogues."""
        from abx_b200 import ops
        pm = (mask[:, :, None] * mask[:, None, :]).to(act.dtype)
        n = act.shape[1]
        x = self.norm(act)
        # outgoing: sum_k l[i,k] r[j,k]; incoming: sum_k l[k,i] r[k,j] = the same product on the transposed LN output
        xt = x if self.outgoing else self.norm(act, transpose_n=n)
        w, b = self._glu_weight()
        prod = ops.triangle_product(xt, w, b, pm, self.final_norm.weight, self.final_norm.bias, self.final_norm.eps)
        return self.proj_out(prod, act='gate', gate=self.final_gate(x), residual=residual)


class TriangleAttention(nn.Module):
    def __init__(self, config, num_in_pair_channel):
        super().__init__()
        c = config
        self.norm = LayerNorm(num_in_pair_channel)
        self.proj_pair = Linear(num_in_pair_channel, c.num_head, init='linear', bias=False)
        self.attn = Attention(num_in_pair_channel, num_in_pair_channel, num_in_pair_channel, gating=c.gating,
                              num_head=c.num_head, split_first=True)
        self.per_column = c.orientation == 'per_column'
        self.config = c

    def forward(self, pair_act, seq_mask, residual=None):
        """seqformer.py:506-550 (+ residual).  Per-column orientation: the LayerNorm kernel writes its output
        transposed ('b j i c') and the output projection stores back transposed, so no rearrange copies."""
        n = pair_act.shape[1] if self.per_column else 0
        x = self.norm(pair_act, transpose_n=n)
        bias = self.proj_pair(x).permute(0, 3, 1, 2)
        return self.attn.forward_pair(x, bias, seq_mask, residual=residual, transpose_n=n)


class SeqformerIteration(nn.Module):
    def __init__(self, config, seq_channel, pair_channel):
        super().__init__()
        c = config
        self.seq_attn = SeqAttentionWithPairBias(c.seq_attention_with_pair_bias, seq_channel, pair_channel)
        self.seq_transition = Transition(c.seq_transition, seq_channel)
        self.outer_product_mean = OuterProductMean(c.outer_product_mean, seq_channel, pair_channel)
        self.triangle_multiplication_outgoing = TriangleMultiplication(c.triangle_multiplication_outgoing, pair_channel)
        self.triangle_multiplication_incoming = TriangleMultiplication(c.triangle_multiplication_incoming, pair_channel)
        self.triangle_attention_starting_node = TriangleAttention(c.triangle_attention_starting_node, pair_channel)
        self.triangle_attention_ending_node = TriangleAttention(c.triangle_attention_ending_node, pair_channel)
        self.pair_transition = Transition(c.pair_transition, pair_channel)
        self.config = config

    def forward(self, seq_act, pair_act, seq_mask):
        """seqformer.py:569-606 (inference: dropout is the identity)."""
        seq_act = self.seq_attn(seq_act, pair_act, seq_mask, residual=seq_act)
        seq_act = self.seq_transition(seq_act, residual=seq_act)
        pair_act = self.outer_product_mean(seq_act, seq_mask, residual=pair_act)
        mf = seq_mask.to(pair_act.dtype)
        pair_act = self.triangle_multiplication_outgoing(pair_act, mf, residual=pair_act)
        pair_act = self.triangle_multiplication_incoming(pair_act, mf, residual=pair_act)
        pair_act = self.triangle_attention_starting_node(pair_act, seq_mask, residual=pair_act)
        pair_act = self.triangle_attention_ending_node(pair_act, seq_mask, residual=pair_act)
        pair_act = self.pair_transition(pair_act, residual=pair_act)
        return seq_act, pair_act


class Seqformer(nn.Module):
    def __init__(self, config):
        super().__init__()
        c = config
        self.blocks = nn.ModuleList([
            SeqformerIteration(c.seqformer, c.seq_channel + c.index_embed_size, c.pair_channel + 2 * c.index_embed_size)
            for _ in range(c.seqformer_num_block)])

    def forward(self, seq_act, pair_act, mask, is_recycling=True):
        for block in self.blocks:
            seq_act, pair_act = block(seq_act, pair_act, mask)
        return seq_act, pair_act


class EmbeddingAndSeqformer(nn.Module):

    def __init__(self, config):
        super().__init__()
        c = as_config(config)
        if c.esm.enabled:
            raise NotImplementedError('ESM2 conditioning needs the fair-esm package and esm2_t36_3B weights, which are '
                                      'not available offline; set embeddings_and_seqformer.esm.enabled=false')
        self.num_token = RESTYPE_NUM + 3
        self.num_region = NUM_AB_REGIONS + 1
        self.proj_aa_type = nn.Embedding(self.num_token, c.seq_channel, padding_idx=20)
        self.encode_residue_emb = ResidueEmbedding(c)
        self.encode_pair_emb = PairEmbedding(c)
        self.aa_proj = nn.Sequential(LayerNorm(c.seq_channel), Linear(c.seq_channel, c.seq_channel), nn.ReLU(),
                                     Linear(c.seq_channel, c.seq_channel))
        self.proj_rel_pos = nn.Embedding(c.max_relative_feature * 2 + 2, c.pair_channel)
        if c.recycle_features:
            self.prev_seq_norm = LayerNorm(c.seq_channel + c.index_embed_size)
            self.prev_pair_norm = LayerNorm(c.pair_channel + 2 * c.index_embed_size)
        if c.recycle_pos:
            self.proj_prev_pos = nn.Embedding(c.prev_pos.num_bins, c.pair_channel + 2 * c.index_embed_size)
        self.seqformer = Seqformer(c)
        self.config = c
        self._static = None          # (seq_static, pair_static) of the current complex, see static_embeddings

    # ---- step-invariant part of the embeddings ---------------------------------------------------------
    def static_embeddings(self, batch):
        """Everything in seqformer.py:176-207 that does not depend on the diffused residues or on t:
        antigen sequence embedding, relative-position pair embedding, ResidueEmbedding and PairEmbedding
        (both see fixed residues only).  Returns (seq_static [B,N,C] with zeros in the antibody sequence
        embedding slot, pair_static [B,N,N,Cz])."""
        c = self.config
        n_ab = batch['anchor_flag'].shape[1]
        residx = batch['residx']

        def relpos(pos):
            off = pos[:, None, :] - pos[:, :, None]
            return torch.clip(off + c.max_relative_feature, min=0, max=2 * c.max_relative_feature) + 1

        ag_seq = self.aa_proj(self.proj_aa_type(batch['seq'][:, n_ab:]))
        B, N = batch['seq'].shape
        seq_static = self.encode_residue_emb(batch).clone()
        seq_static[:, n_ab:] += ag_seq
        pair_static = self.encode_pair_emb(batch).clone()
        pair_static[:, :n_ab, :n_ab] += self.proj_rel_pos(relpos(residx[:, :n_ab]))                   # pair_concat :24-45
        pair_static[:, n_ab:, n_ab:] += self.proj_rel_pos(relpos(residx[:, n_ab:]))
        return seq_static, pair_static

    # inputs the static embeddings read (encoder.py:149-175,211-269; seqformer.py:176-207)
    _STATIC_KEYS = ('mask', 'fixed_mask', 'cdr_def', 'chain_id', 'residx', 'seq', 'atom14_gt_positions',
                    'atom14_gt_exists', 'torsion_angles_sin_cos')

    def _single_complex(self, batch):
        """True when every batch element carries the same complex (independent samples of one complex: the
        sampler's batching), i.e. one static embedding serves the whole batch."""
        for k in self._STATIC_KEYS:
            v = batch[k]
            if v.shape[0] > 1 and not torch.equal(v, v[:1].expand_as(v)):
                return False
        keep = torch.logical_and(batch['mask'], batch['fixed_mask'])
        st = batch['seq_t'] * keep                      # the encoders only see seq_t on fixed residues
        return bool(st.shape[0] == 1 or torch.equal(st, st[:1].expand_as(st)))

    def cache_static(self, batch):
        """Evaluate the static embeddings once and reuse them in every later forward until `clear_static()`:
        a single [1,...] copy when all batch elements are samples of the same complex, else one per element
        (a collated batch of different complexes, `--batch_size > 1`)."""
        if self._single_complex(batch):
            batch = {k: (v[:1] if torch.is_tensor(v) and v.dim() > 0 else v) for k, v in batch.items()}
        self._static = self.static_embeddings(batch)

    def clear_static(self):
        self._static = None

    def forward(self, batch):
        c = self.config
        seq_t = batch['seq_t']
        n_ab = batch['anchor_flag'].shape[1]
        B, N = seq_t.shape
        seq_static, pair_static = self._static if self._static is not None else self.static_embeddings(batch)

        te = get_timestep_embedding(batch['t'], c.index_embed_size)                                   # Embedder :93-119
        ab_seq = self.proj_aa_type(seq_t[:, :n_ab].long())
        seq_act = torch.cat([F.pad(ab_seq, (0, 0, 0, N - n_ab)) + seq_static, te[:, None, :].expand(B, N, -1)], dim=-1).float()
        # pair input in one kernel: concat(static, te_i, te_j) [_cross_concat: both blocks equal te[b]]
        #   + LayerNorm(prev_pair) + proj_prev_pos[prev_pos]                                        (:193-222)
        from abx_b200 import ops
        use_prev = c.recycle_features and 'prev_pair' in batch
        use_pos = c.recycle_pos and 'prev_pos' in batch
        if pair_static.shape[0] == 1:
            pair_act = ops.pair_input(pair_static, te, batch['prev_pair'] if use_prev else None,
                                      self.prev_pair_norm if use_prev else None,
                                      batch['prev_pos'] if use_pos else None, self.proj_prev_pos.weight if use_pos else None)
        else:                    # per-element static embeddings (no complex-level cache): assemble with torch ops
            pair_act = torch.cat([pair_static.expand(B, -1, -1, -1), te[:, None, None, :].expand(B, N, N, -1),
                                  te[:, None, None, :].expand(B, N, N, -1)], dim=-1).float()
            if use_prev:
                pair_act = pair_act + self.prev_pair_norm(batch['prev_pair'])
            if use_pos:
                pair_act = pair_act + self.proj_prev_pos(batch['prev_pos'])
        if c.recycle_features and 'prev_seq' in batch:
            seq_act = seq_act + self.prev_seq_norm(batch['prev_seq'])
        return self.seqformer(seq_act, pair_act, mask=batch['mask'], is_recycling=batch.get('is_recycling', True))
