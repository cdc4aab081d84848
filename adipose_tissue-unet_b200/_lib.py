"""ctypes binding of libadipose_b200.so (C ABI: include/adipose_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` /
``python -m adipose_unet_b200.build``.  There is no fallback: a missing library
raises ImportError-like RuntimeError on first use, and ``adp_create`` fails on a
machine without an sm_100 GPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libadipose_b200.so")

PREC_FP32, PREC_BF16, PREC_BF16_SIMT, PREC_BF16X3 = 0, 1, 2, 3
PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16, "bf16_simt": PREC_BF16_SIMT, "bf16x3": PREC_BF16X3}
BLEND_GAUSSIAN, BLEND_LINEAR = 0, 1
OPT_ADAM, OPT_ADAMW = 0, 1
OPTIMIZERS = {"adam": OPT_ADAM, "adamw": OPT_ADAMW}


class ProfRow(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("launches", C.c_int64), ("ms", C.c_double),
                ("flops", C.c_double), ("bytes", C.c_double)]


_P = C.c_void_p
_I = C.c_int
_F = C.c_float
_I64 = C.c_int64

# name -> (restype, argtypes): every symbol include/adipose_b200.h declares
SIGNATURES = {
    "adp_abi_version": (_I, []),
    "adp_last_error": (C.c_char_p, []),
    "adp_device_count": (_I, []),
    "adp_create": (_I, [_I, _I, _I, _I, C.POINTER(_P)]),
    "adp_destroy": (_I, [_P]),
    "adp_precision": (_I, [_P]),
    "adp_synchronize": (_I, [_P]),
    "adp_stream": (_P, [_P]),
    "adp_set_option": (_I, [_P, C.c_char_p, _I]),
    "adp_set_weight": (_I, [_P, C.c_char_p, _P, C.POINTER(_I64), _P, _I64]),
    "adp_get_weight": (_I, [_P, C.c_char_p, _P, _I64, _P, _I64]),
    "adp_weights_ready": (_I, [_P]),
    "adp_predict": (_I, [_P, _P, _I, _I, _F, _F, C.POINTER(_I), _I, _P]),
    "adp_predict_u8": (_I, [_P, _P, _I, _I, _I, _F, _F, C.POINTER(_I), _I, _P]),
    "adp_tta_ops": (_I, [_I, C.POINTER(_I)]),
    "adp_tta_combine": (_I, [_P, _P, _I, C.POINTER(_I), _I, _P]),
    "adp_debug_layer": (_I, [_P, C.c_char_p, _I, _P, _I64, C.POINTER(_I64)]),
    "adp_threshold_metrics": (_I, [_P, _P, _P, _I64, _F, _P, C.POINTER(_I64)]),
    "adp_threshold_sweep": (_I, [_P, _P, _P, _I64, _P, _I, _P]),
    "adp_boundary_refine": (_I, [_P, _P, _I, _I, _I, _I, _I, _F, _F, _P]),
    "adp_blend_reconstruct": (_I, [_P, _I, _P, _I, _I, _I, _P, _P, _P, _I, _I, _P]),
    "adp_wsi_begin": (_I, [_P, _I, _I, _I, _I, _I, _P]),
    "adp_wsi_push_tiles": (_I, [_P, _P, _I, _P, _P, _F, _F, C.POINTER(_I), _I]),
    "adp_wsi_push_from_slide": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _F, _F, C.POINTER(_I), _I, _I]),
    "adp_wsi_replay_deferred": (_I, [_P]),
    "adp_wsi_push_probs": (_I, [_P, _P, _I, _P, _P]),
    "adp_wsi_export": (_I, [_P, _I, _I, _P, _P]),
    "adp_wsi_import_add": (_I, [_P, _I, _I, _P, _P]),
    "adp_wsi_finalize": (_I, [_P, _I, _I, _F, _P, _P, _P, C.POINTER(_I64)]),
    "adp_wsi_end": (_I, [_P]),
    "adp_jpeg_decode": (_I, [_P, _P, _P, _I, _I, _P, _P, C.POINTER(_P), C.POINTER(_P)]),
    "adp_wsi_push_tiles_u8": (_I, [_P, _P, _I, _I, _P, _P, _F, _F, C.POINTER(_I), _I]),
    "adp_wsi_aux_begin": (_I, [_P, _I]),
    "adp_wsi_push_aux": (_I, [_P, _I, _I, _P, _I, _I, _P, _P]),
    "adp_wsi_export_u8": (_I, [_P, _I, _I, _I, _I, _I, _P]),
    "adp_wsi_export_f32": (_I, [_P, _I, _I, _I, _P]),
    "adp_wsi_finalize_auxgt": (_I, [_P, _I, _I, _I, _F, _P, _P, C.POINTER(_I64)]),
    "adp_tile_fat_percent": (_I, [_P, _P, _I, _I64, _F, _P]),
    "adp_tiff_write_lzw": (_I, [C.c_char_p, _P, _I64, _I64, _I, _I, _I]),
    "adp_loss_metrics": (_I, [_P, _P, _P, _I64, _P, C.POINTER(C.c_double)]),
    "adp_loss_metrics_ex": (_I, [_P, _P, _P, _I, _I64, _I64, _F, _F, _F, _P, C.POINTER(C.c_double)]),
    "adp_train_set_loss": (_I, [_P, _F, _F, _F]),
    "adp_train_set_deep_supervision": (_I, [_P, _I, _F, _F, _F]),
    "adp_train_outputs": (_I, [_P]),
    "adp_train_begin": (_I, [_P, _I, _I, _F, C.c_uint64]),
    "adp_train_forward": (_I, [_P, _P, _P, _I, C.POINTER(_P), C.POINTER(C.c_double)]),
    "adp_train_loss": (_I, [C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "adp_train_backward": (_I, [_P, C.POINTER(C.c_double), _I]),
    "adp_train_sums_buffer": (_I, [_P, C.POINTER(_P), C.POINTER(_I)]),
    "adp_train_sums_read": (_I, [_P, _P, _I]),
    "adp_train_grad_buffer": (_I, [_P, C.POINTER(_P), C.POINTER(_I64)]),
    "adp_train_accuracy_read": (_I, [_P, C.POINTER(C.c_double)]),
    "adp_train_grad_buckets": (_I, [_P, C.POINTER(_I64), C.POINTER(_I64), _I]),
    "adp_train_bucket_wait": (_I, [_P, _I, _P]),
    "adp_train_join": (_I, [_P, _P]),
    "adp_train_grad_read": (_I, [_P, _P, _I64]),
    "adp_train_grad_write": (_I, [_P, _P, _I64]),
    "adp_train_get_grad": (_I, [_P, C.c_char_p, _P, _I64, _P, _I64]),
    "adp_train_probs": (_I, [_P, _P, _I64]),
    "adp_train_apply": (_I, [_P, _I, _F, _F, C.c_double, C.c_double, _F, _F, _I]),
    "adp_train_step": (_I, [_P, _P, _P, _I, _I, _F, _F, _I, C.POINTER(C.c_double)]),
    "adp_train_iterations": (_I64, [_P]),
    "adp_train_end": (_I, [_P]),
    "adp_adam_update": (_I, [_P, _P, _P, _P, _P, _I64, _I64, _I, _F, C.c_double, C.c_double, _F, _F]),
    "adp_profile_enable": (_I, [_P, _I]),
    "adp_profile_reset": (_I, [_P]),
    "adp_profile_read": (_I, [_P, C.POINTER(ProfRow), _I]),
    "adp_launch_count": (_I64, [_P]),
}

_lib = None


class AdiposeError(RuntimeError):
    pass


def load():
    """Load the shared library and bind every declared symbol (no GPU needed for this)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AdiposeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU/PyTorch fallback for this path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here == ABI drift
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    if rc < 0:
        msg = load().adp_last_error()
        raise AdiposeError(f"libadipose_b200 error {rc}: {msg.decode() if msg else ''}")
    return rc


def ptr(a):
    """void* of a NumPy array (C-contiguous), a torch tensor (host or CUDA) or an int address."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):       # torch tensor
        assert a.is_contiguous()
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


def int_array(vals):
    arr = (C.c_int * len(vals))(*vals)
    return arr
