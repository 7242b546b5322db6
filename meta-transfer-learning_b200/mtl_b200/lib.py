"""ctypes loader for libmtl_b200.so (the C ABI declared in include/mtl_b200.h)."""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmtl_b200.so")
_lock = threading.Lock()
_lib = None


class MtlError(RuntimeError):
    pass


def library_path() -> str:
    return _LIB_PATH


class ModelCfg(C.Structure):
    _fields_ = [(k, C.c_int) for k in
                "n_enc n_dec d_model n_heads d_k d_v d_inner rank vocab n_freq".split()]


class CBatch(C.Structure):
    _fields_ = [("x", C.c_void_p), ("lens", C.c_void_p), ("trg", C.c_void_p),
                ("B", C.c_int), ("T", C.c_int), ("L", C.c_int), ("n", C.c_int),
                ("hyp_out", C.c_void_p), ("gold_out", C.c_void_p), ("ce_out", C.c_void_p)]


class LmCfg(C.Structure):
    _fields_ = [(k, C.c_int) for k in "vocab ninp nhid nlayers".split()]


class MetaHParams(C.Structure):
    _fields_ = [("lr", C.c_float), ("val_scale", C.c_float), ("clip", C.c_int), ("max_norm", C.c_float),
                ("dropout", C.c_float), ("label_smoothing", C.c_float), ("seed", C.c_ulonglong)]


class CLane(C.Structure):
    _fields_ = [("theta", C.c_void_p), ("grad", C.c_void_p), ("workspace", C.c_void_p),
                ("workspace_bytes", C.c_longlong)]


class MetaStepArgs(C.Structure):
    _fields_ = [("theta", C.c_void_p), ("copy_grad", C.c_void_p), ("pe_enc", C.c_void_p), ("pe_dec", C.c_void_p),
                ("n_tasks", C.c_int), ("train", C.POINTER(CBatch)), ("val", C.POINTER(CBatch)),
                ("n_lanes", C.c_int), ("lanes", C.POINTER(CLane)), ("hp", MetaHParams),
                ("results", C.c_void_p), ("seed_slot", C.c_void_p), ("use_graph", C.c_int)]


_P, _I, _F, _LL, _ULL, _U = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_ulonglong, C.c_uint
_D = C.c_double

# name -> (restype, argtypes); every symbol include/mtl_b200.h declares
SIGNATURES = {
    "mtl_last_error": (C.c_char_p, []),
    "mtl_abi_version": (_I, []),
    "mtl_launch_count": (_ULL, []),
    "mtl_session_create": (_I, [C.POINTER(ModelCfg), C.POINTER(_P)]),
    "mtl_session_destroy": (None, [_P]),
    "mtl_session_set_gemm_mode": (_I, [_P, _I]),
    "mtl_session_set_op_mode": (_I, [_P, _I, _I]),
    "mtl_session_set_flag": (_I, [_P, C.c_char_p, _I]),
    "mtl_param_arena_floats": (_LL, [_P]),
    "mtl_param_count": (_I, [_P]),
    "mtl_param_info": (_I, [_P, _I, C.POINTER(_LL), C.POINTER(_LL)]),
    "mtl_workspace_bytes": (_LL, [_P, _I, _I, _I]),
    "mtl_asr_forward": (_I, [_P, _P, _P, _P, _P, _LL, C.POINTER(CBatch), _F, _ULL, _F, _P, C.POINTER(_P),
                             C.POINTER(_I)]),
    "mtl_asr_backward": (_I, [_P, _P, _P, _F, _P, _I, _P]),
    "mtl_spectrogram": (_I, [_P, _I, _I, _I, _P, _P, _I, _I, _P, _P]),
    "mtl_encode_workspace_bytes": (_LL, [_P, _I, _I]),
    "mtl_asr_encode": (_I, [_P, _P, _P, _P, _LL, C.POINTER(CBatch), _P, _P]),
    "mtl_greedy_workspace_bytes": (_LL, [_P, _I, _I, _I]),
    "mtl_asr_greedy": (_I, [_P, _P, _P, _P, _LL, _P, _I, _I, _I, _I, _P, _P]),
    "mtl_meta_task": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _LL, C.POINTER(CBatch), C.POINTER(CBatch),
                           C.POINTER(MetaHParams), _P, _P]),
    "mtl_meta_tasks": (_I, [_P, C.POINTER(MetaStepArgs), _P]),
    "mtl_region_a_floats": (_LL, [_P]),
    "mtl_stream_wait_region_a": (_I, [_P, _P]),
    "mtl_graph_stats": (_I, [_P, C.POINTER(_ULL), C.POINTER(_ULL)]),
    "mtl_meta_finish": (_I, [_P, _P, _P, _P, _P, _P, _D, _D, _D, _D, _I, _F, _P, _LL, _P]),
    "mtl_arena_zero": (_I, [_P, _LL, _P]),
    "mtl_arena_copy": (_I, [_P, _P, _LL, _P]),
    "mtl_arena_axpy": (_I, [_P, _P, _F, _LL, _P]),
    "mtl_arena_sgd": (_I, [_P, _P, _F, _LL, _P]),
    "mtl_arena_clip": (_I, [_P, _LL, _F, _P, _P]),
    "mtl_arena_adam": (_I, [_P, _P, _P, _P, _P, _D, _D, _D, _D, _LL, _P]),
    "mtl_lm_param_floats": (_LL, [C.POINTER(LmCfg)]),
    "mtl_lm_param_count": (_I, [C.POINTER(LmCfg)]),
    "mtl_lm_param_info": (_I, [C.POINTER(LmCfg), _I, C.POINTER(_LL), C.POINTER(_LL)]),
    "mtl_lm_workspace_bytes": (_LL, [C.POINTER(LmCfg), _I, _I]),
    "mtl_lm_pass": (_I, [C.POINTER(LmCfg), _I, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _F, _ULL, _F, _P, _LL, _P, _P, _P]),
    "mtl_lm_meta_step": (_I, [C.POINTER(LmCfg), _I, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _I, _I, _P, _F, _F, _F, _F,
                              _ULL, _P, _P, _LL, _P, _P, _P]),
    "mtl_gemm": (_I, [_I, _I, _I, _I, _I, _I, _F, _P, _I, _P, _I, _F, _P, _I, _P, _I, _P, _I, _P]),
    "mtl_lowrank_pair": (_I, [_I, _I, _I, _I, _I, _I, _I, _P, _I, _P, _P, _P, _P, _P, _I, _I, _P]),
    "mtl_gemm_repeat": (_I, [_I, _I, _I, _I, _I, _I, _I, _P, _I, _P, _I, _F, _P, _I, _I, _P]),
    "mtl_debug_gemm_stamps": (_I, [_P]),
    "mtl_debug_attn_stamps": (_I, [_P]),
    "mtl_debug_gemm_span": (_I, [_P]),
    "mtl_debug_pass_buffers": (_I, [_P, _P]),
    "mtl_ln_fwd": (_I, [_P, _P, _P, _P, _P, _P, _I, _F, _ULL, _U, _P, _P, _P, _I, _I, _P]),
    "mtl_ln_bwd": (_I, [_P, _P, _P, _P, _P, _F, _ULL, _U, _P, _P, _I, _P, _P, _I, _I, _P]),
    "mtl_attn_fwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _ULL, _U, _P, _P, _P]),
    "mtl_attn_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _ULL, _U, _P, _P, _P, _P, _P]),
    "mtl_ce_fwd": (_I, [_P, _I, _P, _I, _I, _F, _P, _P, _P, _P, _P]),
    "mtl_ce_bwd": (_I, [_P, _I, _P, _P, _P, _F, _F, _P, _I, _I, _P]),
    "mtl_conv1_fwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "mtl_conv3x3_relu_fwd": (_I, [_I, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "mtl_conv3x3_relu_pool_fwd": (_I, [_I, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "mtl_conv3x3_bwd_scratch_floats": (_LL, [_I, _I, _I, _I, _I, _I]),
    "mtl_conv3x3_bwd": (_I, [_I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "mtl_conv1_wgrad": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "mtl_feat_transpose": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "mtl_embed": (_I, [_P, _P, _P, _F, _ULL, _U, _P, _P, _P, _I, _I, _I, _P]),
    "mtl_maxpool2_fwd": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "mtl_maxpool2_relu_bwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "mtl_dec_preprocess": (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _P]),
}


def get_lib():
    """Loads the library once; raises MtlError (never falls back) if it has not been built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(_LIB_PATH):
            raise MtlError(f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
        lib = C.CDLL(_LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(rc: int):
    if rc != 0:
        msg = get_lib().mtl_last_error()
        raise MtlError(f"libmtl_b200 error {rc}: {msg.decode() if msg else '?'}")
