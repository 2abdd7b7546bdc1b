"""ctypes binding of libpassport_sm100.so (C ABI declared in include/passport_sm100.h).

The library is the only compute path of this package: if it is missing or a call fails, an exception is
raised — there is deliberately no PyTorch/CPU fallback.
"""
import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpassport_sm100.so")

PP_ABI_VERSION = 16
PP_NORM_NONE, PP_NORM_BN_TRAIN, PP_NORM_BN_EVAL, PP_NORM_GN = 0, 1, 2, 3
PP_ALGO_AUTO, PP_ALGO_TCGEN05, PP_ALGO_SIMT = 0, 1, 2
PP_WS_FWD, PP_WS_BWD = 0, 1
PP_DTYPE_BF16, PP_DTYPE_TF32 = 0, 1
PP_FLAG_ACC_DW, PP_FLAG_ACC_DGAMMA, PP_FLAG_ACC_DBETA, PP_FLAG_SHARE_SM = 1, 2, 4, 8

#: every symbol include/passport_sm100.h declares (tests check the .so exports all of them)
EXPORTS = (
    "pp_version", "pp_last_error", "pp_device_info", "pp_workspace_bytes", "pp_weight_prep", "pp_key_pool",
    "pp_passport_affine_fwd", "pp_passport_affine_bwd", "pp_sign_loss_fwd", "pp_sign_loss_bwd",
    "pp_conv_block_fwd", "pp_conv_block_fwd_res", "pp_conv_block_bwd", "pp_conv_block_bwd_dz", "pp_maxpool_fwd", "pp_maxpool_bwd", "pp_conv_fwd_raw", "pp_conv_dgrad", "pp_conv_wgrad",
    "pp_sgd_step", "pp_debug_last_timeout", "pp_launch_count", "pp_profile_enable", "pp_profile_read",
    "pp_add_relu_fwd", "pp_add_relu_bwd", "pp_passport_key_grad", "pp_signature_verify",
    "pp_sgd_step_dev", "pp_ce_top1", "pp_passport_conv_fwd", "pp_passport_conv_bwd",
    "pp_debug_fused",
)


class PPConvDesc(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("O", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32),
        ("stride", C.c_int32), ("pad", C.c_int32),
        ("norm", C.c_int32), ("relu", C.c_int32), ("z_f32", C.c_int32),
        ("eps", C.c_float), ("momentum", C.c_float),
        ("algo", C.c_int32), ("groups", C.c_int32), ("flags", C.c_int32), ("dtype", C.c_int32),
    ]


PP_SIG_MAX_LAYERS = 64


class PPSigLayer(C.Structure):
    _fields_ = [
        ("w_oihw", C.c_void_p), ("S_skey", C.c_void_p), ("b_sign", C.c_void_p),
        ("O", C.c_int32), ("K", C.c_int32), ("gamma_offset", C.c_int32), ("C", C.c_int32),
    ]


_lib = None
_lock = threading.Lock()
_vp, _fp, _dp, _i, _f, _sz = C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_size_t
_desc = C.POINTER(PPConvDesc)

_PROTOS = {
    "pp_version": (C.c_int, []),
    "pp_last_error": (C.c_char_p, []),
    "pp_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "pp_workspace_bytes": (C.c_int, [_desc, _i, C.POINTER(C.c_size_t)]),
    "pp_weight_prep": (C.c_int, [_desc, _fp, _vp, _vp, _vp]),
    "pp_key_pool": (C.c_int, [_desc, _i, _fp, _dp, _vp]),
    "pp_passport_affine_fwd": (C.c_int, [_desc, _vp, _dp, _dp, _fp, _f, _fp, _fp, _fp, _fp, _vp]),
    "pp_passport_affine_bwd": (C.c_int, [_desc, _dp, _dp, _fp, _fp, _f, _fp, _fp, _fp, _fp, _i, _vp]),
    "pp_passport_key_grad": (C.c_int, [_desc, _i, _vp, _fp, _fp, _f, _fp, _fp, _fp, _dp, _fp, _fp, _vp]),
    "pp_signature_verify": (C.c_int, [_i, C.POINTER(PPSigLayer), _vp, _fp, _vp]),
    "pp_sign_loss_fwd": (C.c_int, [_i, _fp, _fp, _f, _fp, _fp, _vp]),
    "pp_sign_loss_bwd": (C.c_int, [_i, _fp, _fp, _f, _fp, _fp, _vp]),
    "pp_conv_block_bwd_dz": (C.c_int, [_desc, _vp, _vp, _vp, _fp, _fp, _fp, _fp, _vp, _vp, _fp, _fp, _vp, _vp, _sz, _vp]),
    "pp_maxpool_fwd": (C.c_int, [_i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp]),
    "pp_maxpool_bwd": (C.c_int, [_i, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "pp_conv_block_fwd": (C.c_int, [_desc, _vp, _vp, _fp, _fp, _fp, _fp, _vp, _vp, _fp, _fp, _vp, _sz, _vp]),
    "pp_conv_block_fwd_res": (C.c_int, [_desc, _vp, _vp, _fp, _fp, _fp, _fp, _vp, _vp, _fp, _fp, _vp, _vp, _sz, _vp]),
    "pp_conv_block_bwd": (C.c_int, [_desc, _vp, _vp, _vp, _vp, _fp, _fp, _fp, _fp, _vp, _fp, _fp, _fp, _vp, _sz, _vp]),
    "pp_passport_conv_fwd": (C.c_int, [_desc, _vp, _vp, _fp, _dp, _dp, _fp, _fp, _fp, _f, _fp, _fp, _vp, _vp, _fp, _fp,
                                       _fp, _fp, _fp, _fp, _vp, _sz, _vp]),
    "pp_passport_conv_bwd": (C.c_int, [_desc, _vp, _vp, _vp, _vp, _fp, _fp, _fp, _fp, _dp, _dp, _fp, _f, _fp, _vp, _fp,
                                       _fp, _fp, _vp, _sz, _vp]),
    "pp_conv_fwd_raw": (C.c_int, [_desc, _vp, _vp, _vp, _vp, _sz, _vp]),
    "pp_conv_dgrad": (C.c_int, [_desc, _vp, _vp, _vp, _vp]),
    "pp_conv_wgrad": (C.c_int, [_desc, _vp, _vp, _fp, _vp, _sz, _vp]),
    "pp_sgd_step": (C.c_int, [_sz, _fp, _fp, _fp, _f, _f, _f, _i, _vp]),
    "pp_sgd_step_dev": (C.c_int, [_sz, _fp, _fp, _fp, _fp, _vp]),
    "pp_ce_top1": (C.c_int, [_i, _i, _vp, _i, _vp, _fp, _fp, _fp, _i, _vp]),
    "pp_add_relu_fwd": (C.c_int, [_sz, _vp, _vp, _vp, _vp]),
    "pp_add_relu_bwd": (C.c_int, [_sz, _vp, _vp, _vp, _vp]),
    "pp_debug_last_timeout": (C.c_int, []),
    "pp_debug_fused": (C.c_int, [_i]),
    "pp_launch_count": (C.c_longlong, [_i]),
    "pp_profile_enable": (C.c_int, [_i]),
    "pp_profile_read": (C.c_int, [_i, _i, _i, _i, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]),
}


def load():
    """Load (once) and return the ctypes handle. Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m deepipr_b200.build` "
                "(this package has no fallback path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.pp_version() != PP_ABI_VERSION:
            raise RuntimeError(f"libpassport_sm100 ABI {lib.pp_version()} != expected {PP_ABI_VERSION}; rebuild")
        _lib = lib
    return _lib


def last_error():
    msg = load().pp_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status, what=""):
    if status != 0:
        raise RuntimeError(f"libpassport_sm100 {what} failed ({status}): {last_error()}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())
