"""GPU parity of the IPA kernels (through the C-ABI, via the reference-shaped `InvariantPointAttention`)
against the reference's golden output and the CPU oracle."""
import pytest
import torch

from abx_b200.utils.weights import np_randn
from oracle import model as M
from oracle import quat as Q
from tests.util import golden, maxabs, seeded_params

pytestmark = pytest.mark.gpu

PREFIX = M.SN + 'attention_module.'


def make_ipa():
    from abx_b200.model.folding import InvariantPointAttention
    ipa = InvariantPointAttention(M.IPA_CONF, 128)
    P = seeded_params()
    ipa.load_state_dict({k[len(PREFIX):]: v for k, v in P.items() if k.startswith(PREFIX)}, strict=True)
    return ipa.cuda().eval(), P


def unchunk_bias(bias, B, N):
    """Chunked key-major pair bias [B, ceil(N/8), N, 100] (include/abx_b200.h) -> head-major [B,12,N,N]."""
    nc = (N + 7) // 8
    t = bias.view(B, nc, N, 100)[..., :96].reshape(B, nc, N, 8, 12)
    return t.permute(0, 4, 2, 1, 3).reshape(B, 12, N, nc * 8)[..., :N]


def test_linear_matches_torch(cuda_device):
    import ctypes
    from abx_b200 import lib
    L = lib.load()
    for (m, n, k) in ((37, 6, 256), (700, 1152, 256), (129, 256, 2112), (64, 65, 20)):
        x, w, b, r = np_randn(1, m, k), np_randn(2, n, k), np_randn(3, n), np_randn(4, m, n)
        for relu in (0, 1):
            y = torch.empty(m, n, device='cuda')
            xc, wc, bc, rc = x.cuda(), w.cuda(), b.cuda(), r.cuda()
            lib.check(L.abx_linear_f32(lib.stream(), m, n, k, lib.ptr(xc), k, lib.ptr(wc), lib.ptr(bc), lib.ptr(rc), relu,
                                       lib.ptr(y), n))
            ref = torch.nn.functional.linear(x.double(), w.double(), b.double())
            ref = (ref.clamp(min=0) if relu else ref) + r.double()
            assert maxabs(y.cpu(), ref) < 1e-5 * float(ref.abs().max()) * (k ** 0.5)


def test_ipa_matches_reference_golden(cuda_device):
    g = golden('ipa')
    ipa, _ = make_ipa()
    x, z = np_randn(211, 2, 37, 256), np_randn(212, 2, 37, 37, 128)
    with torch.no_grad():
        out = ipa(x.cuda(), z.cuda(), g['mask'].cuda(), (g['rots'].cuda(), g['trans'].cuda()))
    assert maxabs(out.cpu(), g['out']) < 2e-5 * float(g['out'].abs().max())


@pytest.mark.parametrize('B,N', [(1, 1), (1, 7), (1, 33), (3, 100), (5, 131), (2, 350), (8, 350), (1, 700), (1, 1536)])
def test_ipa_matches_oracle(cuda_device, B, N):
    """Ragged masks, sizes off the key-chunk / row-tile grids (every tile height the host picks: R = 1..20), the
    north-star size at the benchmark batch (B=8: one 20-row tile per SM) and the supported maximum; features and output."""
    ipa, P = make_ipa()
    gen = torch.Generator().manual_seed(100 + N)
    x, z = np_randn(300 + N, B, N, 256), np_randn(400 + N, B, N, N, 128)
    q = torch.randn(B, N, 4, generator=gen); q = q / q.norm(dim=-1, keepdim=True)
    rots, trans = Q.quat_to_rot(q), torch.randn(B, N, 3, generator=gen) * 2.0
    mask = torch.ones(B, N)
    if N > 8:
        mask[-1, -(N // 5):] = 0                   # padded tail on the last batch element
        mask[0, 3] = 0                             # a hole
    ref, parts = M.ipa_forward(P, x, z, mask, rots, trans, return_parts=True)
    with torch.no_grad():
        args = (x.cuda(), z.cuda(), mask.cuda(), (rots.cuda(), trans.cuda()))
        feats = ipa.attention_features(*args)
        out = ipa(*args)
        bias = ipa.pair_bias(z.cuda())
        out_b = ipa(*args, pair_bias=bias, residual=x.cuda())
    assert maxabs(feats.cpu(), parts['feats']) < 3e-5 * float(parts['feats'].abs().max())
    assert maxabs(out.cpu(), ref) < 3e-5 * float(ref.abs().max())
    assert maxabs(out_b.cpu(), ref + x) < 3e-5 * float((ref + x).abs().max())
    ref_bias = (3 ** -0.5) * M.linear(P, PREFIX + 'proj_pair', z).permute(0, 3, 1, 2)
    assert maxabs(unchunk_bias(bias.cpu(), B, N), ref_bias) < 1e-5 * float(ref_bias.abs().max())


@pytest.mark.parametrize('B,N', [(2, 45), (1, 350)])
def test_ipa_reference_point_moves(cuda_device, B, N):
    """The fused kernel fixes each (row, head)'s softmax reference point in the first key chunk and moves it only when a
    later logit exceeds it by 2^64 (accumulators in tensor memory / registers are then rescaled).  Exercise both ways it
    moves: the first chunks fully masked (reference point -FLT_MAX -> first real logit), and far-away leading keys whose
    point-distance term puts them > 64 log-2 units below a later key."""
    ipa, P = make_ipa()
    gen = torch.Generator().manual_seed(900 + N)
    x, z = np_randn(910 + N, B, N, 256), np_randn(920 + N, B, N, N, 128)
    q = torch.randn(B, N, 4, generator=gen); q = q / q.norm(dim=-1, keepdim=True)
    rots, trans = Q.quat_to_rot(q), torch.randn(B, N, 3, generator=gen) * 2.0
    trans[:, :11] += 60.0                          # keys 0-10 sit ~100 units away from every other residue
    mask = torch.ones(B, N)
    mask[0, :19] = 0                               # leading masked keys on batch element 0 (two full chunks + 3)
    mask[0, 30] = 0
    ref, parts = M.ipa_forward(P, x, z, mask, rots, trans, return_parts=True)
    with torch.no_grad():
        args = (x.cuda(), z.cuda(), mask.cuda(), (rots.cuda(), trans.cuda()))
        feats = ipa.attention_features(*args)
        out = ipa(*args)
    assert maxabs(feats.cpu(), parts['feats']) < 3e-5 * float(parts['feats'].abs().max())
    assert maxabs(out.cpu(), ref) < 3e-5 * float(ref.abs().max())


@pytest.mark.parametrize('rows,sizes', [(20, '2,350'), (17, '1,350'), (7, '3,100')])
def test_ipa_rescale_path_at_every_tile_height(cuda_device, rows, sizes):
    """The rare path made the common one: ABX_IPA_RESCALE_GAP=1 moves the softmax reference points (and rescales the
    tensor-memory accumulators through the service warps) in almost every chunk, ABX_IPA_ROWS fixes the tile height (full
    20-row tiles use all three MMA issuers, both converter warpgroups, every A ring and accumulator columns >= 256).
    tools/ipa_debug.py compares features and layer output with the oracle (3e-5) and reads the kernel's watchdog."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = {**os.environ, 'ABX_IPA_RESCALE_GAP': '1.0', 'ABX_IPA_ROWS': str(rows)}
    r = subprocess.run([sys.executable, os.path.join(root, 'tools', 'ipa_debug.py'), sizes], env=env, capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]


def test_ipa_is_frame_invariant(cuda_device):
    """Size-independent property at full size: a global rigid motion of all frames leaves the output unchanged."""
    ipa, _ = make_ipa()
    B, N = 2, 350
    gen = torch.Generator().manual_seed(77)
    x, z = np_randn(501, B, N, 256).cuda(), np_randn(502, B, N, N, 128).cuda()
    q = torch.randn(B, N, 4, generator=gen); q = q / q.norm(dim=-1, keepdim=True)
    rots, trans = Q.quat_to_rot(q), torch.randn(B, N, 3, generator=gen) * 2.0
    g = torch.randn(4, generator=gen); g = g / g.norm()
    G, s = Q.quat_to_rot(g), torch.randn(3, generator=gen)
    rots2 = torch.einsum('rd,bndm->bnrm', G, rots)
    trans2 = torch.einsum('rd,bnd->bnr', G, trans) + s
    mask = torch.ones(B, N).cuda()
    with torch.no_grad():
        a = ipa(x, z, mask, (rots.cuda(), trans.cuda()))
        b = ipa(x, z, mask, (rots2.cuda(), trans2.cuda()))
    assert maxabs(a.cpu(), b.cpu()) < 1e-4 * float(a.abs().max())


def test_ipa_rejects_bad_arguments(cuda_device):
    from abx_b200.lib import AbxError
    ipa, _ = make_ipa()
    x, z = torch.zeros(1, 4, 256), torch.zeros(1, 4, 4, 128)
    with pytest.raises(AbxError):
        ipa(x, z, torch.ones(1, 4), (torch.eye(3).expand(1, 4, 3, 3), torch.zeros(1, 4, 3)))     # CPU tensors


def test_ipa_graph_replay_is_bit_identical(cuda_device):
    """The layer-call (PDL-chained kernels, bulk-copy rings) captured in a CUDA graph replays bit-identically,
    with a ragged batch (B=5) and a padded mask."""
    ipa, _ = make_ipa()
    B, N = 5, 70
    x, z = np_randn(31, B, N, 256).cuda(), np_randn(32, B, N, N, 128).cuda()
    q = np_randn(33, B, N, 4); q = q / q.norm(dim=-1, keepdim=True)
    rig = (Q.quat_to_rot(q).cuda(), np_randn(34, B, N, 3).cuda())
    mask = torch.ones(B, N); mask[1, -9:] = 0; mask[4, 2] = 0
    mask = mask.cuda()
    with torch.no_grad():
        base = ipa(x, z, mask, rig)
        for _ in range(3):
            assert torch.equal(ipa(x, z, mask, rig), base)
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                captured = ipa(x, z, mask, rig)
        captured.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(captured, base)
