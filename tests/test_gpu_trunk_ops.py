"""GPU parity of the trunk kernels behind the C ABI: streaming LayerNorm, fused pair-bias attention and the
transposing GEMM epilogue, each against a plain float64/float32 torch restatement of the reference op
(seqformer.py:283-301, :506-550)."""
import pytest
import torch

from abx_b200.utils.weights import np_randn
from tests.util import maxabs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('shape', [(3, 7, 192), (2, 50, 50, 128), (1, 33, 544), (5, 256), (2, 9, 9, 1024), (4, 384), (3, 5, 512), (2, 768)])
def test_layernorm_matches_torch(cuda_device, shape):
    from abx_b200 import ops
    x = (np_randn(1, *shape) * 3 + 0.5).cuda()
    g, b = np_randn(2, shape[-1]).cuda(), np_randn(3, shape[-1]).cuda()
    y = ops.layer_norm(x, g, b, 1e-5)
    ref = torch.nn.functional.layer_norm(x.double(), (shape[-1],), g.double(), b.double(), 1e-5)
    assert maxabs(y.cpu(), ref.cpu()) < 2e-6 * float(ref.abs().max())


def test_layernorm_transposed_output(cuda_device):
    from abx_b200 import ops
    x = np_randn(4, 2, 37, 37, 192).cuda()
    g, b = np_randn(5, 192).cuda(), np_randn(6, 192).cuda()
    y = ops.layer_norm(x, g, b, 1e-5, transpose_n=37)
    ref = torch.nn.functional.layer_norm(x, (192,), g, b, 1e-5).transpose(1, 2)
    assert maxabs(y.cpu(), ref.cpu()) < 1e-5


def _attention_reference(qkv, bias, key_mask, H):
    """seqformer.py:283-301 in float64."""
    B, S, L, C3 = qkv.shape
    D = C3 // 3 // H
    q, k, v = (t.reshape(B, S, L, H, D).permute(0, 1, 3, 2, 4).double() for t in qkv.chunk(3, dim=-1))
    logits = torch.einsum('bshqd,bshkd->bshqk', q * D ** -0.5, k) + bias.double()[:, None]
    if key_mask is not None:
        logits = logits.masked_fill(~key_mask.bool()[:, None, None, None, :], torch.finfo(torch.float32).min)
    w = torch.softmax(logits, dim=-1)
    return torch.einsum('bshqk,bshkd->bsqhd', w, v).reshape(B, S, L, H * D)


@pytest.mark.parametrize('impl', ['tc5', 'mma', 'simt'])
@pytest.mark.parametrize('B,S,L,H,D', [(1, 3, 37, 4, 48), (2, 5, 350, 4, 48), (1, 2, 400, 2, 32), (1, 1, 1, 1, 16), (1, 2, 65, 3, 64),
                                       (1, 2, 351, 4, 48)])
def test_pair_attention_matches_reference(cuda_device, B, S, L, H, D, impl):
    from abx_b200 import ops
    qkv = np_randn(10, B, S, L, 3 * H * D).cuda()
    bias = (np_randn(11, B, H, L, L) * 2).cuda()
    mask = torch.ones(B, L, dtype=torch.bool)
    if L > 8:
        mask[0, -5:] = False
        mask[-1, 3] = False
    out = ops.pair_attention(qkv, bias, mask.cuda(), H, impl=impl)
    ref = _attention_reference(qkv, bias, mask.cuda(), H)
    assert maxabs(out.cpu(), ref.cpu()) < 3e-6 * max(1.0, float(ref.abs().max()))
    out2 = ops.pair_attention(qkv, bias, None, H, impl=impl)
    ref2 = _attention_reference(qkv, bias, None, H)
    assert maxabs(out2.cpu(), ref2.cpu()) < 3e-6 * max(1.0, float(ref2.abs().max()))


def test_pair_attention_all_keys_masked_is_uniform(cuda_device):
    """masked_fill(finfo.min) semantics: a fully masked row softmaxes to uniform weights, not NaN."""
    from abx_b200 import ops
    qkv = np_randn(12, 1, 2, 40, 3 * 4 * 48).cuda()
    bias = np_randn(13, 1, 4, 40, 40).cuda()
    mask = torch.zeros(1, 40, dtype=torch.bool).cuda()
    ref = _attention_reference(qkv, bias, mask, 4)
    for impl in ('tc5', 'mma', 'simt'):
        out = ops.pair_attention(qkv, bias, mask, 4, impl=impl)
        assert torch.isfinite(out).all()
        assert maxabs(out.cpu(), ref.cpu()) < 3e-6


def test_gemm_transposed_store(cuda_device):
    from abx_b200 import ops
    n = 21
    x, w, b = np_randn(20, 2, n, n, 64).cuda(), np_randn(21, 96, 64).cuda(), np_randn(22, 96).cuda()
    res, gate = np_randn(23, 2, n, n, 96).cuda(), np_randn(24, 2, n, n, 96).cuda()
    lin = torch.nn.functional.linear(x.double(), w.double(), b.double())
    y = ops.linear(x, w, b, residual=res, transpose_n=n)
    assert maxabs(y.cpu(), (lin.transpose(1, 2) + res.double()).cpu()) < 3e-6 * float(lin.abs().max())
    y = ops.linear(x, w, b, act='sigmoid_mul', gate=gate, transpose_n=n)
    assert maxabs(y.cpu(), (torch.sigmoid(lin) * gate.double()).transpose(1, 2).cpu()) < 3e-6 * float(gate.abs().max())


def test_pair_attention_fused_output_gate(cuda_device):
    from abx_b200 import ops
    B, S, L, H, D = 1, 3, 70, 4, 48
    qkvg = np_randn(30, B, S, L, 4 * H * D).cuda()
    bias = np_randn(31, B, H, L, L).cuda()
    mask = torch.ones(B, L, dtype=torch.bool).cuda()
    out = ops.pair_attention(qkvg, bias, mask, H, gated=True)
    qkv = qkvg[..., :3 * H * D].contiguous()
    ref = _attention_reference(qkv, bias, mask, H) * torch.sigmoid(qkvg[..., 3 * H * D:].double())
    assert maxabs(out.cpu(), ref.cpu()) < 3e-6 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize('n', [21, 350])
def test_triangle_product_matches_einsum(cuda_device, n):
    """GLU GEMM (channel-major store) + batched NT product + channel-major LayerNorm == the reference chain
    proj*sigmoid(gate)*mask -> einsum('bikc,bjkc->bijc') -> LayerNorm  (seqformer.py:452-502, outgoing)."""
    from abx_b200 import ops
    B, K, C = 2 if n < 100 else 1, 192, 128
    x = np_randn(50, B, n, n, K).cuda()
    wl, wlg, wr, wrg = (np_randn(51 + i, C, K).cuda() / K ** 0.5 for i in range(4))
    bl, blg, br, brg = (np_randn(55 + i, C).cuda() * 0.1 for i in range(4))
    g, be = np_randn(60, C).cuda(), np_randn(61, C).cuda()
    mask = torch.ones(B, n, device='cuda')
    mask[0, -3:] = 0
    pm = mask[:, :, None] * mask[:, None, :]
    ws, bs = [], []
    for (pw, pb), (gw, gb) in (((wl, bl), (wlg, blg)), ((wr, br), (wrg, brg))):
        for c in range(0, C, 64):
            ws += [pw[c:c + 64], gw[c:c + 64]]
            bs += [pb[c:c + 64], gb[c:c + 64]]
    out = ops.triangle_product(x, torch.cat(ws).contiguous(), torch.cat(bs).contiguous(), pm, g, be)
    xd = x.double()
    F = torch.nn.functional
    left = F.linear(xd, wl.double(), bl.double()) * torch.sigmoid(F.linear(xd, wlg.double(), blg.double())) * pm.double()[..., None]
    right = F.linear(xd, wr.double(), br.double()) * torch.sigmoid(F.linear(xd, wrg.double(), brg.double())) * pm.double()[..., None]
    ref = F.layer_norm(torch.einsum('bikc,bjkc->bijc', left, right), (C,), g.double(), be.double(), 1e-5)
    assert out.shape == (B, n, n, C)
    assert maxabs(out.cpu(), ref.cpu()) < 2e-5 * max(1.0, float(ref.abs().max()))
    # second call reuses the cached channel-major buffers (pad columns must still be zero)
    out2 = ops.triangle_product(x, torch.cat(ws).contiguous(), torch.cat(bs).contiguous(), pm, g, be)
    assert torch.equal(out, out2)


def test_pair_input_matches_torch(cuda_device):
    from abx_b200 import ops
    B, N, Cs, Ct = 3, 19, 128, 32
    C = Cs + 2 * Ct
    stat, te, prev = np_randn(70, 1, N, N, Cs).cuda(), np_randn(71, B, Ct).cuda(), (np_randn(72, B, N, N, C) * 2 + 1).cuda()
    norm = torch.nn.LayerNorm(C).cuda()
    with torch.no_grad():
        norm.weight.copy_(np_randn(73, C)); norm.bias.copy_(np_randn(74, C))
    emb = np_randn(75, 15, C).cuda()
    pos = torch.randint(0, 15, (B, N, N), generator=torch.Generator().manual_seed(0)).cuda()
    base = torch.cat([stat.expand(B, -1, -1, -1), te[:, None, None, :].expand(B, N, N, -1), te[:, None, None, :].expand(B, N, N, -1)], -1)
    with torch.no_grad():
        ref = base + norm(prev) + emb[pos]
        got = ops.pair_input(stat, te, prev, norm, pos, emb)
        assert maxabs(got.cpu(), ref.cpu()) < 1e-5
        assert torch.equal(ops.pair_input(stat, te), base)
        assert maxabs(ops.pair_input(stat, te, prev, norm).cpu(), (base + norm(prev)).cpu()) < 1e-5


def test_outer_product_matches_torch(cuda_device):
    from abx_b200 import ops
    left, right = np_randn(80, 2, 23, 32).cuda(), np_randn(81, 2, 23, 32).cuda()
    ref = torch.cat([left[:, None, :, :] * right[:, :, None, :], left[:, None, :, :] - right[:, :, None, :]], dim=-1)
    assert torch.equal(ops.outer_product(left, right), ref)
