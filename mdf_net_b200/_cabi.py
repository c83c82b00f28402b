"""ctypes binding of the C ABI declared in include/mdf_b200.h.

The shared library is the product: there is no Python / PyTorch / CPU fallback.  If
libmdf_b200.so is missing this module raises at first use (build it with
`python -m mdf_net_b200.build` or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import c_char_p, c_float, c_int, c_size_t, c_void_p

from .build import LIB_PATH, TUNING_LIB_PATH

_lock = threading.Lock()
_lib = None

STATUS_NAMES = {0: "MDF_OK", -1: "MDF_ERR_INVALID_SHAPE", -2: "MDF_ERR_UNSUPPORTED", -3: "MDF_ERR_NULL_POINTER",
                -4: "MDF_ERR_WORKSPACE", -5: "MDF_ERR_NOT_DEVICE", -6: "MDF_ERR_CUDA"}

_P = c_void_p
_I = c_int
# name -> (restype, argtypes); mirrors include/mdf_b200.h one to one
SIGNATURES = {
    "mdf_abi_version": (_I, []),
    "mdf_status_string": (c_char_p, [_I]),
    "mdf_last_cuda_error": (_I, []),
    "mdf_homo_warp_workspace_bytes": (c_size_t, [_I]),
    "mdf_homo_warp_fwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, c_size_t, _P]),
    "mdf_cost_volume_workspace_bytes": (c_size_t, [_I] * 7),
    "mdf_cost_volume_fwd": (_I, [_P, _I, _P, _P, _P, _I, _P, _P, _P, _P, _P, c_float, _P, _P,
                                 _I, _I, _I, _I, _I, _I, _P, _P, c_size_t, _P]),
    "mdf_cost_volume_fwd_ex": (_I, [_P, _I, _P, _P, _P, _I, _P, _P, _P, _P, _P, c_float, _P, _P,
                                    _I, _I, _I, _I, _I, _I, _P, _P, c_size_t, _I, _P, _P, _P]),
    "mdf_fpn_out_prepped_fwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "mdf_cost_volume_prepped_workspace_bytes": (c_size_t, [_I] * 2),
    "mdf_cost_volume_fwd_prepped": (_I, [_P, _P, _P, _I, _P, _P, _P, _I, _P, _P, _P, _P, _P, c_float, _P, _P,
                                         _I, _I, _I, _I, _I, _P, _P, c_size_t, _P]),
    "mdf_variance_volume_workspace_bytes": (c_size_t, [_I] * 6),
    "mdf_variance_volume_fwd": (_I, [_P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, c_size_t, _P]),
    "mdf_softmax_regress_fwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _I, _I, _I, _I, _P]),
    "mdf_softmax_regress_fit_fwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "mdf_prob_head_fwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "mdf_prob_head_fwd_ex": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _I, _P]),
    "mdf_depth_regression_fwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "mdf_confidence_fwd": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "mdf_cost_volume_train_workspace_bytes": (c_size_t, [_I] * 7),
    "mdf_cost_volume_train_fwd": (_I, [_P, _I, _P, _P, _P, _I, _P, _P, _P, _P, _P, c_float, _P, _P, _I,
                                       _I, _I, _I, _I, _I, _I, _P, _P, _P, c_size_t, _P]),
    "mdf_cost_volume_bwd": (_I, [_P, _I, _P, _P, _P, _I, _P, _P, _P, _P, _P, c_float, _P, _P, _I,
                                 _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "mdf_bn_running_update": (_I, [_P, _I, c_float, _P, _P, _P, _P]),
    "mdf_hypos_fit_fwd": (_I, [_P, _P, _I, _P, _I, _I, _I, _I, _I, _P, _P]),
    "mdf_hypos_generate_fwd": (_I, [_P, _P, _P, _I, c_float, _I, _I, _I, _I, _I, _P, _P]),
    "mdf_geo_filter_workspace_bytes": (c_size_t, [_I]),
    "mdf_geo_filter_fwd": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P, c_float, _I, c_float, c_float, _P, _P, _P, _P, _P, _P,
                                _P, c_size_t, _P]),
    "mdf_debug_sample_positions": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P]),
}
# only exported by the tuning build (MDF_B200_TUNING=1 in the environment makes tools/ load it instead of the product)
TUNING_SIGNATURES = {"mdf_debug_read_trace": (_I, [_P, _I])}


class MdfError(RuntimeError):
    def __init__(self, fn: str, status: int, detail: str = ""):
        self.status = status
        super().__init__(f"{fn} failed: {STATUS_NAMES.get(status, status)} ({detail})")


def lib() -> ctypes.CDLL:
    """Load libmdf_b200.so once (thread safe).  Missing library is a hard error."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                path = TUNING_LIB_PATH if os.environ.get("MDF_B200_TUNING") == "1" else LIB_PATH
                if not os.path.exists(path):
                    raise ImportError(
                        f"{path} not found: the sm_100a library is the only implementation of this "
                        "package (no CPU / PyTorch fallback). Build it with `python -m mdf_net_b200.build`.")
                handle = ctypes.CDLL(path)
                for name, (res, args) in {**SIGNATURES, **TUNING_SIGNATURES}.items():
                    try:
                        fn = getattr(handle, name)
                    except AttributeError:
                        continue  # optional entry points are checked by tests/test_cabi_exports.py
                    fn.restype, fn.argtypes = res, args
                _lib = handle
    return _lib


def check(fn: str, status: int) -> None:
    if status != 0:
        l = lib()
        detail = l.mdf_status_string(status).decode()
        if status == -6:
            detail += f", cudaError={l.mdf_last_cuda_error()}"
        raise MdfError(fn, status, detail)


def ptr_array(ptrs):
    """HOST array of device pointers, as the multi-view entry points take them."""
    return (c_void_p * len(ptrs))(*ptrs)
