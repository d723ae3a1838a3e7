"""Thin torch-tensor wrappers over the C-ABI dense-layer kernels (include/abx_b200.h).

`linear` is the drop-in for `torch.nn.functional.linear` on CUDA fp32 tensors used by the model modules:
y = act(x W^T + b) (+ residual) on the tcgen05 3xTF32 GEMM (fp32-level accuracy).  No CPU fallback.
"""
import torch

from abx_b200 import lib

ACT = {None: 0, 'none': 0, 'relu': 1, 'gate': 2, 'sigmoid': 3, 'sigmoid_mul': 4, 'glu': 5}


_W_LO = {}          # (data_ptr, shape) -> (version, weight (keeps the storage, hence the address, alive), lo part)
_W_LO_MAX = 1024


def weight_lo(w):
    """Low part of a static fp32 weight for the 3xTF32 GEMM: w - tf32_trunc(w) (13 low mantissa bits cleared), cached per
    weight tensor and recomputed when the tensor is modified in place (load_state_dict bumps `_version`).  None for large
    weights: fetching a second operand stream only pays while the weight stays L2-resident (measured: -3 % on the trunk's
    layers, +5 % on an 8192^3 product)."""
    if w.numel() > (1 << 20):
        return None
    key = (w.data_ptr(), tuple(w.shape))
    hit = _W_LO.get(key)
    if hit is not None and hit[0] == w._version:
        return hit[2]
    with torch.no_grad():
        lo = w - (w.view(torch.int32) & -8192).view(torch.float32)
    if len(_W_LO) >= _W_LO_MAX:
        _W_LO.pop(next(iter(_W_LO)))
    _W_LO[key] = (w._version, w, lo)
    return lo


def linear(x, weight, bias=None, act=None, residual=None, gate=None, row_scale=None, out=None, tile_n=0, transpose_n=0):
    """x [..., K] (last dim contiguous, uniform row stride), weight [Nout, K] -> [..., Nout].

    v = xW^T + b;  act: None | 'relu' | 'sigmoid' | 'gate' (v * sigmoid(gate)) | 'sigmoid_mul' (sigmoid(v) * gate)
    with gate [..., Nout];  then y = v * row_scale[...] + residual[..., Nout].
    transpose_n = n: x is [B,n,n,K] and the result (and residual) are indexed 'b j i c' (see the header)."""
    L = lib.load()
    K = x.shape[-1]
    Nout = weight.shape[0]
    lead = x.shape[:-1]
    x2 = x.reshape(-1, K)
    if x2.dtype != torch.float32:
        x2 = x2.float()
    if x2.stride(-1) != 1 or (x2.shape[0] > 1 and x2.stride(0) % 4 != 0) or x2.data_ptr() % 16 != 0:
        x2 = x2.contiguous()
    M = x2.shape[0]
    ldx = x2.stride(0) if M > 1 else K
    w = weight.detach()
    if not w.is_contiguous():
        w = w.contiguous()
    n_y = Nout // 2 if act == 'glu' else Nout           # 'glu': projection/gate column pairs collapse (see the header)
    y = out if out is not None else torch.empty(lead + (n_y,), device=x.device, dtype=torch.float32)
    res = residual.reshape(-1, Nout).contiguous() if residual is not None else None
    g = gate.reshape(-1, Nout).contiguous() if gate is not None else None
    rs = row_scale.reshape(-1).to(torch.float32).contiguous() if row_scale is not None else None
    with lib.device_guard(x2):
        lib.check(L.abx_gemm_tf32x3_wlo(lib.stream(), M, Nout, K, lib.ptr_any(x2), ldx, lib.ptr(w, torch.float32),
                                        lib.ptr(weight_lo(w)), K,
                                        lib.ptr(bias.detach() if bias is not None else None), lib.ptr(res), lib.ptr(g),
                                        lib.ptr(rs), ACT[act], transpose_n, lib.ptr(y), n_y, tile_n))
    return y


def set_gemm_backend(name):
    """'auto' | 'simt' | 'tcgen05' for the node GEMMs inside abx_ipa_forward."""
    lib.check(lib.load().abx_set_gemm_backend({'auto': 0, 'simt': 1, 'tcgen05': 2}[name]))


def layer_norm(x, weight, bias, eps=1e-5, transpose_n=0):
    """torch.nn.functional.layer_norm over the last dimension on the streaming LayerNorm kernel.
    transpose_n = n: x is [B,n,n,C] and the result is returned as 'b j i c' (contiguous)."""
    L = lib.load()
    C = x.shape[-1]
    xc = x if (x.is_contiguous() and x.dtype == torch.float32) else x.float().contiguous()
    y = torch.empty_like(xc)
    with lib.device_guard(xc):
        lib.check(L.abx_layernorm(lib.stream(), xc.numel() // C, C, lib.ptr(xc), lib.ptr(weight.detach(), torch.float32),
                                  lib.ptr(bias.detach(), torch.float32), float(eps), transpose_n, lib.ptr(y)))
    return y


def _bias_tiles(bias, key_mask):
    """Pair bias [B,H,L,L] (+ key mask [B,L]) in the layout of abx_pair_attention_tc5 (include/abx_b200.h): blocks of 32 query rows
    x 64 keys, key-major inside a block, scaled by log2(e); masked keys -> finfo.min, padding keys -> -inf."""
    B, H, L, _ = bias.shape
    Lq, Lk = (L + 31) // 32 * 32, (L + 63) // 64 * 64
    t = torch.zeros(B, H, Lq, Lk, device=bias.device, dtype=torch.float32)
    t[:, :, :, L:] = float('-inf')
    t[:, :, :L, :L] = bias * 1.4426950408889634
    if key_mask is not None:
        t[:, :, :, :L].masked_fill_(~key_mask.to(torch.bool)[:, None, None, :], torch.finfo(torch.float32).min)
    return t.view(B, H, Lq // 32, 32, Lk // 64, 64).permute(0, 1, 4, 2, 5, 3).contiguous()


def pair_attention(qkv, bias, key_mask, num_head, impl='tc5', gated=False):
    """Fused attention core for TriangleAttention.  qkv [B,S,L,3*H*D] (q | k | v slices of one projection;
    with `gated=True` [B,S,L,4*H*D] = q | k | v | gate pre-activation), bias [B,H,L,L], key_mask [B,L]
    (bool/float, None = keep all)  ->  [B,S,L,H*D] (times sigmoid(gate) when gated).
    impl: 'tc5' (tcgen05 3xTF32, default; head dim 64 falls back to 'mma'), 'mma' (mma.sync 3xTF32) or 'simt'."""
    L_ = lib.load()
    B, S, L, Cw = qkv.shape
    HD = Cw // (4 if gated else 3)
    D = HD // num_head
    assert qkv.is_contiguous() and qkv.dtype == torch.float32
    bias = bias.float().contiguous()
    km = key_mask.to(torch.float32).contiguous() if key_mask is not None else None
    out = torch.empty(B, S, L, HD, device=qkv.device, dtype=torch.float32)
    base, esz = qkv.data_ptr(), 4
    with lib.device_guard(qkv):
        if impl == 'tc5' and L_.abx_pair_attention_tc5_supported(L, D):
            lib.check(L_.abx_pair_attention_tc5(lib.stream(), B, S, L, num_head, D, base, base + HD * esz, base + 2 * HD * esz, Cw,
                                                lib.ptr(_bias_tiles(bias, key_mask)), (base + 3 * HD * esz) if gated else None,
                                                lib.ptr(out)))
        else:
            lib.check(L_.abx_pair_attention_impl(lib.stream(), {'tc5': 2, 'simt': 1, 'mma': 2}[impl], B, S, L, num_head, D, base,
                                                 base + HD * esz, base + 2 * HD * esz, Cw,
                                                 lib.ptr(bias), lib.ptr(km), (base + 3 * HD * esz) if gated else None, lib.ptr(out)))
    return out


_CM_BUFFERS = {}


def _cm_buffer(tag, shape, device):
    """Zero-padded channel-major scratch: ONE grow-only flat buffer per (tag, device).  The kernels never write the
    pad columns, which must read as zero, so the view is zero-filled whenever its shape differs from the previous
    call's (the old contents would otherwise land in the new pad positions); with a steady shape nothing is cleared."""
    key = (tag, str(device))
    numel = 1
    for d in shape:
        numel *= int(d)
    flat, last = _CM_BUFFERS.get(key, (None, None))
    if flat is None or flat.numel() < numel:
        flat, last = torch.zeros(numel, device=device, dtype=torch.float32), tuple(shape)
    view = flat[:numel].view(*shape)
    if last != tuple(shape):
        view.zero_()
    _CM_BUFFERS[key] = (flat, tuple(shape))
    return view


def release_scratch():
    """Drop the cached scratch buffers (long runs over complexes of many sizes can return the memory)."""
    _CM_BUFFERS.clear()


def triangle_product(x, w_glu, b_glu, pair_mask, norm_weight, norm_bias, eps=1e-5):
    """sum_k left[i,k] * right[j,k] per channel + final LayerNorm for TriangleMultiplication (seqformer.py:452-502).
    x [B,n,n,Cin] = LN(pair) (pass the 'b j i c' transposed LN output for the incoming orientation); w_glu / b_glu:
    the tile-interleaved [left | right] x [proj | gate] weights (see TriangleMultiplication._glu_weight);
    pair_mask [B,n,n].  Returns LN(product) as [B,n,n,C] row-major."""
    L_ = lib.load()
    B, n, _, K = x.shape
    Nout = w_glu.shape[0]
    C = Nout // 4                              # channels of left (= of right)
    npad = (n + 3) // 4 * 4
    xc = x if x.is_contiguous() else x.contiguous()
    lr = _cm_buffer('lr', (B, 2 * C, n, npad), x.device)
    prod = _cm_buffer('prod', (B, C, n, npad), x.device)
    rs = pair_mask.reshape(-1).to(torch.float32).contiguous()
    out = torch.empty(B, n, n, C, device=x.device, dtype=torch.float32)
    with lib.device_guard(xc):
        st = lib.stream()
        lib.check(L_.abx_gemm_tf32x3_glu_cm_wlo(st, B * n * n, Nout, K, lib.ptr(xc), K, lib.ptr(w_glu), lib.ptr(weight_lo(w_glu)), K,
                                                lib.ptr(b_glu), lib.ptr(rs), n, npad, lib.ptr(lr)))
        a_ptr = lr.data_ptr()
        b_ptr = a_ptr + C * n * npad * 4        # right channels follow the left ones inside each batch element
        lib.check(L_.abx_gemm_tf32x3_batched_nt(st, B * C, n, npad, C, 2 * C * n, B * 2 * C * n - C * n, a_ptr, b_ptr,
                                                lib.ptr(prod), npad))
        lib.check(L_.abx_layernorm_cm(st, B, C, n, npad, lib.ptr(prod), lib.ptr(norm_weight.detach()), lib.ptr(norm_bias.detach()),
                                      float(eps), lib.ptr(out)))
    return out


def pair_input(stat, te, prev_pair=None, norm=None, prev_pos=None, emb=None):
    """concat(stat, te, te) + LayerNorm(prev_pair) + emb[prev_pos] in one kernel (seqformer.py:193-222).
    stat [1|B?,N,N,Cs] (first element used), te [B,Ct], prev_pair [B,N,N,C], norm = nn.LayerNorm, prev_pos [B,N,N]."""
    L_ = lib.load()
    N, Cs = stat.shape[-2], stat.shape[-1]
    B, Ct = te.shape
    C = Cs + 2 * Ct
    st = stat.reshape(-1, N, N, Cs)[0].float().contiguous()
    tec = te.float().contiguous()
    pp = prev_pair.float().contiguous() if prev_pair is not None else None
    pos = prev_pos.long().contiguous() if prev_pos is not None else None
    y = torch.empty(B, N, N, C, device=st.device, dtype=torch.float32)
    with lib.device_guard(st):
        lib.check(L_.abx_pair_input(lib.stream(), B, N, C, Cs, Ct, lib.ptr(st), lib.ptr(tec), lib.ptr(pp),
                                    lib.ptr(norm.weight.detach()) if pp is not None else None,
                                    lib.ptr(norm.bias.detach()) if pp is not None else None,
                                    float(norm.eps) if pp is not None else 1e-5, lib.ptr(pos),
                                    lib.ptr(emb.detach().float().contiguous()) if pos is not None else None, lib.ptr(y)))
    return y


def outer_product(left, right):
    """[B,N,C] x2 -> [B,N,N,2C] = concat(left_j * right_i, left_j - right_i) (seqformer.py:392-407)."""
    L_ = lib.load()
    B, N, C = left.shape
    l, r = left.float().contiguous(), right.float().contiguous()
    out = torch.empty(B, N, N, 2 * C, device=l.device, dtype=torch.float32)
    with lib.device_guard(l):
        lib.check(L_.abx_outer_product(lib.stream(), B, N, C, lib.ptr(l), lib.ptr(r), lib.ptr(out)))
    return out
