"""Small building blocks shared by the model modules (reference: abx/model/common_modules.py)."""
import numpy as np
import torch
from torch import nn


class ConfigView(dict):
    """Attribute access over a (nested) dict; accepts ml_collections.ConfigDict-like objects too."""

    def __getitem__(self, k):
        v = dict.__getitem__(self, k)
        return ConfigView(v) if isinstance(v, dict) and not isinstance(v, ConfigView) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def get(self, k, default=None):
        v = dict.get(self, k, default)
        return ConfigView(v) if isinstance(v, dict) else v


def as_config(c):
    if isinstance(c, ConfigView):
        return c
    if isinstance(c, dict):
        return ConfigView(c)
    if hasattr(c, 'to_dict'):
        return ConfigView(c.to_dict())
    return c


class LayerNorm(nn.LayerNorm):
    """nn.LayerNorm (same parameters / state_dict keys) on the streaming LayerNorm kernel (abx_layernorm).
    `transpose_n=n` returns the result of a [B,n,n,C] input as 'b j i c'.  CUDA tensors only."""

    def forward(self, x, transpose_n=0):
        from abx_b200 import ops
        C = x.shape[-1]
        if C % 4 != 0 or C > 1024 or len(self.normalized_shape) != 1:
            if not x.is_cuda:
                raise ops.lib.AbxError('abx_b200 layers take CUDA tensors only (no CPU fallback)')
            assert transpose_n == 0
            return super().forward(x)
        return ops.layer_norm(x, self.weight, self.bias, self.eps, transpose_n=transpose_n)


class AbxLinear(nn.Linear):
    """nn.Linear (same parameters / state_dict keys) whose forward runs on the tcgen05 3xTF32 GEMM
    (abx_gemm_tf32x3) with the activation / gate / mask / residual of the surrounding reference code fused
    into the epilogue.  CUDA tensors only; layers whose input width is not a multiple of 4 (two small
    once-per-complex encoder layers) use torch's CUDA matmul."""

    def forward(self, x, act=None, residual=None, gate=None, row_scale=None, transpose_n=0):
        from abx_b200 import ops
        if self.in_features % 4 != 0:
            if not x.is_cuda:
                raise ops.lib.AbxError('abx_b200 layers take CUDA tensors only (no CPU fallback)')
            assert act in (None, 'relu') and gate is None and row_scale is None and transpose_n == 0
            y = torch.nn.functional.linear(x, self.weight, self.bias)
            y = torch.relu(y) if act == 'relu' else y
            return y if residual is None else y + residual
        return ops.linear(x, self.weight, self.bias, act=act, residual=residual, gate=gate, row_scale=row_scale,
                          transpose_n=transpose_n)


def mlp(seq, x, residual=None):
    """Run an nn.Sequential of (LayerNorm | AbxLinear | ReLU) with every ReLU fused into the preceding
    GEMM's epilogue and `residual` into the last one.  Module indices (state_dict keys) are untouched."""
    mods = list(seq)
    i = 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, AbxLinear):
            relu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
            last = i + (2 if relu else 1) >= len(mods)
            x = m(x, act='relu' if relu else None, residual=residual if last else None)
            i += 2 if relu else 1
        else:
            x = m(x)
            i += 1
    return x


def Linear(input_dim, output_dim, init='linear', bias=True, config=None):
    """AbxLinear with the AF2-style initialisers of common_modules.py:11-38."""
    assert init in ('gate', 'final', 'attn', 'relu', 'linear')
    layer = AbxLinear(input_dim, output_dim, bias=bias)
    with torch.no_grad():
        if init in ('gate', 'final'):
            layer.weight.zero_()
        elif init == 'attn':
            nn.init.xavier_uniform_(layer.weight)
        else:
            std = np.sqrt((2.0 if init == 'relu' else 1.0) / input_dim) / 0.87962566103423978
            nn.init.trunc_normal_(layer.weight, mean=0.0, std=std)
        if bias:
            layer.bias.fill_(1.0 if init == 'gate' else 0.0)
    return layer


def pseudo_beta_fn_v2(aatype, all_atom_positions, all_atom_masks=None):
    """common_modules.py:61-83: ideal C-beta from N, CA, C (atom indices 0, 1, 2)."""
    n, ca, c = all_atom_positions[..., 0, :], all_atom_positions[..., 1, :], all_atom_positions[..., 2, :]
    b, cc = ca - n, c - ca
    a = torch.cross(b, cc, dim=-1)
    cb = -0.58273431 * a + 0.56802827 * b - 0.54067466 * cc + ca
    if all_atom_masks is not None:
        return cb, all_atom_masks[..., 1]
    return cb


def dgram_from_positions(positions, num_bins, min_bin, max_bin):
    """common_modules.py:107-120: index of the squared-distance bin, [B,N,N] int64."""
    breaks = torch.linspace(min_bin, max_bin, steps=num_bins - 1, device=positions.device) ** 2
    d2 = torch.sum((positions[:, :, None] - positions[:, None]) ** 2, dim=-1)
    return torch.bucketize(d2, breaks, right=False)      # == sum(d2 > breaks)
