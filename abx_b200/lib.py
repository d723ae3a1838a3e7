"""ctypes binding of the C-ABI library (include/abx_b200.h -> abx_b200/csrc/libabx_b200.so).

There is no fallback: if the library is missing, or a call is made without a CUDA device, the caller
gets an exception.  Build with `python -m abx_b200.build` (or `__graft_entry__.build()`).
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'csrc', 'libabx_b200.so')

_vp, _i, _d, _sz = C.c_void_p, C.c_int, C.c_double, C.c_size_t


class DiffuserConsts(C.Structure):
    """abx_diffuser_consts (include/abx_b200.h)."""
    _fields_ = [('so3_exp_max', _d), ('so3_exp_min', _d), ('so3_g2_coef', _d), ('r3_min_b', _d), ('r3_delta_b', _d),
                ('r3_coord_scale', _d), ('seq_rate', _d), ('num_sigma', _i), ('num_omega', _i)]


class IpaWeights(C.Structure):
    """abx_ipa_weights (include/abx_b200.h)."""
    _fields_ = [(n, _vp) for n in ('w_q_scalar', 'b_q_scalar', 'w_kv_scalar', 'b_kv_scalar', 'w_q_point', 'b_q_point',
                                   'w_kv_point', 'b_kv_point', 'w_pair', 'b_pair', 'point_weights', 'w_final', 'b_final',
                                   'w_proj_cat', 'b_proj_cat')]


# name -> (restype, argtypes); mirrors include/abx_b200.h one to one (tests/test_abi.py cross-checks the header)
SIGNATURES = {
    'abx_last_error': (C.c_char_p, []),
    'abx_version': (_i, []),
    'abx_launch_count': (C.c_uint64, []),
    'abx_reset_launch_count': (None, []),
    'abx_device_check': (_i, [_i, C.POINTER(_i)]),
    'abx_se3_scores': (_i, [_vp, _i, _i, C.POINTER(DiffuserConsts), _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    'abx_so3_score_rotvec': (_i, [_vp, _i, _i, C.POINTER(DiffuserConsts), _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    'abx_igso3_score_series': (_i, [_vp, _i, _i, C.POINTER(DiffuserConsts), _vp, _vp, _i, _vp, _i, _vp]),
    'abx_seq_reverse_rates': (_i, [_vp, _i, _i, C.POINTER(DiffuserConsts), _vp, _vp, _vp, _d, _vp]),
    'abx_se3_reverse_step': (_i, [_vp, _i, _i, C.POINTER(DiffuserConsts), _vp, _i, _vp, _vp, _vp, _vp, _vp, _d, _d, _d,
                                  _vp, _vp, _vp, _i, _vp, _vp]),
    'abx_igso3_build_tables': (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    'abx_linear_f32': (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _i, _vp, _i]),
    'abx_gemm_tf32x3': (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _vp, _i, _i]),
    'abx_gemm_tf32x3_wlo': (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _vp, _i, _i]),
    'abx_gemm_tf32x3_glu_cm': (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _i, _vp, _vp, _i, _i, _vp]),
    'abx_gemm_tf32x3_glu_cm_wlo': (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _vp, _i, _vp, _vp, _i, _i, _vp]),
    'abx_gemm_tf32x3_batched_nt': (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i]),
    'abx_layernorm_cm': (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, C.c_float, _vp]),
    'abx_pair_input': (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, C.c_float, _vp, _vp, _vp]),
    'abx_outer_product': (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    'abx_layernorm': (_i, [_vp, C.c_longlong, _i, _vp, _vp, _vp, C.c_float, _i, _vp]),
    'abx_pair_attention': (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    'abx_pair_attention_impl': (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    'abx_pair_attention_tc5': (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    'abx_pair_attention_tc5_supported': (_i, [_i, _i]),
    'abx_set_gemm_backend': (_i, [_i]),
    'abx_gemm_profile': (_i, [_vp]),
    'abx_attention_profile': (_i, [_vp]),
    'abx_ipa_frame_update': (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'abx_ipa_workspace_bytes': (_sz, [_i, _i]),
    'abx_ipa_pair_bias_floats': (_sz, [_i, _i]),
    'abx_ipa_watchdog_read': (_i, [_vp]),
    'abx_ipa_profile': (_i, [_i, _vp]),
    'abx_ipa_pair_bias': (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp]),
    'abx_ipa_forward': (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, C.POINTER(IpaWeights), _vp, _vp, _vp, _vp, _sz]),
    'abx_ipa_attention_features': (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, C.POINTER(IpaWeights), _vp, _vp, _vp, _sz]),
}

_lib = None


def load():
    """dlopen the library once and attach the signatures.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f'{LIB_PATH} not found: build the CUDA extension first (python -m abx_b200.build). '
                              'abx_b200 has no CPU or PyTorch fallback.')
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


class AbxError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise AbxError(f'abx_b200 C-ABI call failed (code {rc}): {load().abx_last_error().decode()}')


def ptr(t, dtype=None):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise AbxError('abx_b200 kernels take CUDA tensors only (no CPU fallback)')
    if not t.is_contiguous():
        raise AbxError('abx_b200 kernels take contiguous tensors')
    if dtype is not None and t.dtype != dtype:
        raise AbxError(f'expected {dtype}, got {t.dtype}')
    return t.data_ptr()


def ptr_any(t):
    """Device pointer of a CUDA tensor whose rows may be strided (last dim contiguous)."""
    if not t.is_cuda:
        raise AbxError('abx_b200 kernels take CUDA tensors only (no CPU fallback)')
    if t.stride(-1) != 1:
        raise AbxError('innermost dimension must be contiguous')
    return t.data_ptr()


def device_guard(t):
    """Context manager selecting the tensor's CUDA device; CPU tensors are refused (no fallback)."""
    if not t.is_cuda:
        raise AbxError('abx_b200 kernels take CUDA tensors only (no CPU fallback)')
    return torch.cuda.device(t.device)


def stream():
    return torch.cuda.current_stream().cuda_stream


def launch_count():
    return int(load().abx_launch_count())


def reset_launch_count():
    load().abx_reset_launch_count()
