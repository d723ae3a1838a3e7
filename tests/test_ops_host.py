"""Host-side helpers of abx_b200.ops that prepare kernel operands (CPU tensors: no GPU, no library call)."""
import torch

from abx_b200 import ops


def test_bias_tiles_layout_scaling_mask_and_padding():
    """Layout contract of abx_pair_attention_tc5 (include/abx_b200.h): bias_tiles[b][h][kt][it][j][i] = log2(e) bias[b,h,32 it+i,64 kt+j],
    finfo.min at masked keys, -inf at padding keys."""
    g = torch.Generator().manual_seed(5)
    B, H, L = 2, 3, 70
    bias = torch.randn(B, H, L, L, generator=g)
    mask = torch.ones(B, L, dtype=torch.bool)
    mask[1, 5] = False
    mask[0, 69] = False
    t = ops._bias_tiles(bias, mask)
    assert t.shape == (B, H, 2, 3, 64, 32) and t.is_contiguous()
    log2e = 1.4426950408889634
    fmin = torch.finfo(torch.float32).min
    for b, h, i, j in [(0, 0, 0, 0), (1, 2, 33, 64), (0, 1, 69, 68), (1, 0, 31, 63), (0, 2, 64, 1)]:
        assert float(t[b, h, j // 64, i // 32, j % 64, i % 32]) == float(bias[b, h, i, j] * log2e)
    assert float(t[1, 0, 0, 0, 5, 7]) == fmin and float(t[0, 2, 1, 2, 5, 3]) == fmin          # masked keys (1,5) and (0,69)
    assert torch.isinf(t[:, :, 1, :, 6:, :]).all() and (t[:, :, 1, :, 6:, :] < 0).all()      # keys 70..127 are padding
    assert torch.isfinite(t[:, :, 0]).all()                                                  # rows beyond L stay finite
    assert torch.equal(ops._bias_tiles(bias, None)[0, 0, 0, 0, 5, 7], bias[0, 0, 7, 5] * log2e)


def test_weight_lo_is_the_exact_tf32_remainder():
    w = torch.randn(64, 48, generator=torch.Generator().manual_seed(3))
    lo = ops.weight_lo(w)
    hi = w - lo
    assert torch.equal(hi.view(torch.int32) & 8191, torch.zeros_like(hi, dtype=torch.int32))   # 13 low mantissa bits clear
    assert torch.equal(hi + lo, w)
    assert ops.weight_lo(w) is lo                                                              # cached per tensor version
    w.add_(1.0)
    assert ops.weight_lo(w) is not lo
