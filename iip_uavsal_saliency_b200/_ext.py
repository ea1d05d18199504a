"""ctypes binding of libuavsal_b200.so (the C ABI declared in include/uavsal_b200.h).

The library is loaded lazily; a missing or unloadable library is a hard error (there is no CPU or
PyTorch fallback on the product path).  Every entry point returns an int status that is mapped to
RuntimeError / ValueError here.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_void_p

from . import build as _build

_lib = None

P = c_void_p
I = c_int
L = c_int64
F = c_float
ACT = [P, L, I]          # (pointer, plane, ld)

_SIGNATURES = {
    "uavsal_device_ok": [I],
    "uavsal_set_option": [I, I],
    "uavsal_pack_weights": [P, I, I, I, P, P, P, P, F, P, I, I, I, I, P, P, P],
    "uavsal_pack_nchw_f32": [P, I, I, I, I] + ACT + [I, P],
    "uavsal_unpack_nchw_f32": ACT + [I, I, I, I, P, P],
    "uavsal_stem_conv3x3s2": [P, I, I, I, I, P, P] + ACT + [P],
    "uavsal_stem_conv3x3s2_hw": [P, I, I, I, I, P, P] + ACT + [P],
    "uavsal_dw3x3": ACT + [I, I, I, I, I, I, P, P, I] + ACT + [P],
    "uavsal_expand_dw3x3": ACT + [I, I, I, I, P, I, P, I, I, P, P] + ACT + [P],
    "uavsal_dw_project": [P, I, I, I, I, I, P, P, P, I, I, P, I, I] + ACT + ACT + [P],
    "uavsal_mbconv_fused": ACT + [I, I, I, I, P, I, P, I, P, P, P, I, P, I, I] + ACT + ACT + [P],
    "uavsal_dw_project32_hw": [P, I, I, I, I, P, P, P, P] + ACT + [P],
    "uavsal_pw_gemm": ACT + [I, I, P, I, I, P, I, I] + ACT + ACT + [P],
    "uavsal_pw_gemm_simt": ACT + [I, I, P, I, P, I] + ACT + ACT + [P],
    "uavsal_conv3x3": ACT + [I, I, I, I, P, I, P, I, I] + ACT + [P],
    "uavsal_conv3x3_simt": ACT + [I, I, I, I, P, I, P, I] + ACT + [P],
    "uavsal_bilinear_ac": ACT + [I, I, I, I] + ACT + [I, I, I, I, I, P],
    "uavsal_tdiff_cat": ACT + [I, I, I] + ACT + [I, P],
    "uavsal_ctx_sum": ACT + [I, I, I, I] + ACT + [P],
    "uavsal_add": ACT + ACT + [L, I] + ACT + [P],
    "uavsal_twa_sequence": ACT + ACT + [I, I, I, I, P, P, I, P] + ACT + [I, P, P],
    "uavsal_convlstm_sequence": ACT + ACT + [P, I, I, I, I, I, I, P, P, P, I] + ACT + [P],
    "uavsal_dw3x3_dot_sigmoid": [P, I, I, I, I, I, P, P, P, F, P, P, P],
    "uavsal_dw3x3_dot_sigmoid_q16": [P, I, I, I, I, I, P, P, P, F, P, P, P],
    "uavsal_dot_sigmoid": ACT + [L, I, P, F, P, P],
    "uavsal_post_u8": [P, I, I, I, I, I, P, P, P],
    "uavsal_post_f32": [P, I, I, I, I, I, P, P, P],
    "uavsal_conv_first": [P, I, I, I, I, I, I, P, P, I] + ACT + [P],
    "uavsal_maxpool": ACT + [I, I, I, I, I, I, I] + ACT + [P],
    "uavsal_add_act": ACT + ACT + [L, I, I] + ACT + [P],
    "uavsal_metrics4": [P, P, I, I, I, I, P, P, P],
    "uavsal_letterbox_u8": [P, I, I, I, P, I, I, I, P],
    "uavsal_auc_judd": [P, P, I, I, I, P, P, L, P],
    "uavsal_auc_sampled": [P, P, I, I, I, P, P, I, I, c_double, P, P],
}

EXPORTS = ["uavsal_version", "uavsal_arch", "uavsal_last_error", "uavsal_auc_judd_workspace", "uavsal_twa_sync_bytes"] + list(_SIGNATURES)


class UavsalError(RuntimeError):
    pass


def lib_path() -> str:
    return _build.LIB


def load():
    """Load (building first if the .so is missing) and return the ctypes library handle."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("UAVSAL_LIB") or _build.LIB              # UAVSAL_LIB: developer override (A/B of two builds)
    if not os.path.exists(path):
        path = _build.build()
    try:
        lib = ctypes.CDLL(path)
    except OSError as e:  # pragma: no cover - environment dependent
        raise UavsalError("cannot load %s: %s (no fallback path exists)" % (path, e)) from e
    lib.uavsal_version.restype = c_int
    lib.uavsal_arch.restype = c_char_p
    lib.uavsal_last_error.restype = c_char_p
    lib.uavsal_auc_judd_workspace.argtypes = [I, I, I]
    lib.uavsal_auc_judd_workspace.restype = c_int64
    lib.uavsal_twa_sync_bytes.argtypes = [I, I, I]
    lib.uavsal_twa_sync_bytes.restype = ctypes.c_size_t
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_int
    _lib = lib
    # developer knob: UAVSAL_OPTIONS="key=value,key=value" applies uavsal_set_option at load time (A/B runs of bench.py / tools)
    for kv in filter(None, os.environ.get("UAVSAL_OPTIONS", "").split(",")):
        k, v = kv.split("=")
        lib.uavsal_set_option(int(k), int(v, 0))
    return lib


def check(rc: int, what: str = "") -> None:
    if rc == 0:
        return
    msg = load().uavsal_last_error().decode("utf-8", "replace")
    if rc < 0:
        raise ValueError("uavsal-b200 %s: %s (status %d)" % (what, msg, rc))
    raise UavsalError("uavsal-b200 %s: CUDA error %d: %s" % (what, rc, msg))


def call(name: str, *args) -> None:
    check(getattr(load(), name)(*args), name)
