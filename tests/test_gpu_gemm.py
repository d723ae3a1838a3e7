"""GPU parity of the tcgen05 3xTF32 GEMM (abx_gemm_tf32x3) against a float64 reference: the error must be
at the level of an fp32 GEMM (the sampler's 1e-4 A budget rules out single-pass TF32)."""
import pytest
import torch

from abx_b200.utils.weights import np_randn
from tests.util import maxabs

pytestmark = pytest.mark.gpu

SHAPES = [(128, 128, 32), (128, 128, 256), (1400, 1152, 256), (1400, 256, 2112), (37, 6, 256), (129, 20, 128),
          (700, 768, 192), (1000, 192, 768), (33, 40, 36), (4096, 544, 544)]


@pytest.mark.parametrize('tile_n', [0, 32, 64, 128])
@pytest.mark.parametrize('m,n,k', SHAPES)
def test_gemm_matches_float64(cuda_device, m, n, k, tile_n):
    from abx_b200 import ops
    x, w, b = np_randn(1, m, k).cuda(), np_randn(2, n, k).cuda(), np_randn(3, n).cuda()
    y = ops.linear(x, w, b, tile_n=tile_n)
    ref = torch.nn.functional.linear(x.double(), w.double(), b.double())
    fp32 = torch.nn.functional.linear(x, w, b)            # torch's own fp32 GEMM, for scale
    err, err32 = maxabs(y.cpu(), ref.cpu()), maxabs(fp32.cpu(), ref.cpu())
    scale = float(ref.abs().max())
    assert err < 2e-6 * scale * max(1.0, (k / 256) ** 0.5), (err, err32, scale)
    assert err < 8 * err32 + 1e-6 * scale, (err, err32)


def test_gemm_epilogues(cuda_device):
    from abx_b200 import ops
    m, n, k = 300, 200, 192
    x, w, b, r, g = (np_randn(i, *s).cuda() for i, s in enumerate([(m, k), (n, k), (n,), (m, n), (m, n)]))
    lin = torch.nn.functional.linear(x.double(), w.double(), b.double())
    tol = 3e-6 * float(lin.abs().max())
    assert maxabs(ops.linear(x, w, b, act='relu').cpu(), lin.clamp(min=0).cpu()) < tol
    assert maxabs(ops.linear(x, w, b, act='relu', residual=r).cpu(), (lin.clamp(min=0) + r.double()).cpu()) < tol
    assert maxabs(ops.linear(x, w, None, residual=r).cpu(), (lin - b.double() + r.double()).cpu()) < tol
    assert maxabs(ops.linear(x, w, b, act='gate', gate=g).cpu(), (lin * torch.sigmoid(g.double())).cpu()) < tol
    assert maxabs(ops.linear(x, w, b, act='sigmoid').cpu(), torch.sigmoid(lin).cpu()) < 5e-6


def test_gemm_strided_rows_and_batch_dims(cuda_device):
    from abx_b200 import ops
    big = np_randn(5, 3, 50, 320).cuda()
    x = big[..., 64:256]                                  # row stride 320, offset 64 floats: still TMA-able
    w = np_randn(6, 96, 192).cuda()
    y = ops.linear(x, w)
    ref = torch.nn.functional.linear(x.double(), w.double())
    assert y.shape == (3, 50, 96)
    assert maxabs(y.cpu(), ref.cpu()) < 2e-6 * float(ref.abs().max())


def test_gemm_rejects_bad_arguments(cuda_device):
    from abx_b200 import lib, ops
    with pytest.raises(lib.AbxError):
        ops.linear(torch.zeros(4, 6, device='cuda'), torch.zeros(3, 6, device='cuda'))     # K % 4 != 0
    with pytest.raises(lib.AbxError):
        ops.linear(torch.zeros(4, 8), torch.zeros(3, 8))                                   # CPU tensors


def test_gemm_glu_epilogue(cuda_device):
    """act 'glu': 128-column tiles of [64 projections | 64 gates] -> proj * sigmoid(gate) * row_scale."""
    from abx_b200 import ops
    m, k = 333, 192
    x, w, b, rs = np_randn(40, m, k).cuda(), np_randn(41, 512, k).cuda(), np_randn(42, 512).cuda(), np_randn(43, m).cuda()
    y = ops.linear(x, w, b, act='glu', row_scale=rs)
    lin = torch.nn.functional.linear(x.double(), w.double(), b.double()).reshape(m, 4, 2, 64)
    ref = (lin[:, :, 0] * torch.sigmoid(lin[:, :, 1])).reshape(m, 256) * rs.double()[:, None]
    assert y.shape == (m, 256)
    assert maxabs(y.cpu(), ref.cpu()) < 3e-6 * float(ref.abs().max())


def test_gemm_precomputed_weight_lo_matches_in_kernel_split(cuda_device):
    """abx_gemm_tf32x3_wlo with the caller's w - tf32_trunc(w) must give the bits of the in-kernel split (w_lo = NULL), and the
    per-weight cache of ops.weight_lo must follow in-place updates of the weight (load_state_dict bumps `_version`)."""
    from abx_b200 import lib, ops
    L = lib.load()
    m, n, k = 1000, 192, 192
    x, w, b = np_randn(11, m, k).cuda(), np_randn(12, n, k).cuda(), np_randn(13, n).cuda()

    def run(w_lo):
        y = torch.empty(m, n, device='cuda')
        lib.check(L.abx_gemm_tf32x3_wlo(lib.stream(), m, n, k, lib.ptr(x), k, lib.ptr(w), lib.ptr(w_lo), k, lib.ptr(b), None, None, None,
                                        0, 0, lib.ptr(y), n, 0))
        return y

    lo = ops.weight_lo(w)
    assert lo is not None and torch.equal(lo, w - (w.view(torch.int32) & -8192).view(torch.float32))
    assert float((lo.abs() / w.abs().clamp(min=1e-30)).max()) < 2 ** -10
    assert torch.equal(run(lo), run(None))
    y0 = ops.linear(x, w, b)
    w.mul_(1.5)                                            # in place: same storage, new version
    y1 = ops.linear(x, w, b)
    ref = torch.nn.functional.linear(x.double(), w.double(), b.double())
    assert maxabs(y1.cpu(), ref.cpu()) < 3e-6 * float(ref.abs().max())
    assert maxabs(y0.cpu(), ref.cpu()) > 1e-2            # the stale result is far away: the cache did notice the update
    assert ops.weight_lo(torch.zeros(2048, 1024, device='cuda')) is None      # large weights: split inside the kernel
