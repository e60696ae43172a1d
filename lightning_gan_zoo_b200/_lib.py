"""ctypes binding of libhologan_b200.so (the C ABI declared in include/hologan_b200.h).

There is NO fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
import threading

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libhologan_b200.so")

HG_F32, HG_BF16 = 0, 1
HG_NCDHW, HG_NDHWC, HG_PROJ = 0, 1, 2
HG_BORDER_REFERENCE, HG_BORDER_ZERO = 0, 1
HG_TUNE_CTA1024 = 0x100        # OR into `border`: 1024-thread CTAs for the NCDHW 16^3 rotate (same results)

_c_int, _c_void_p, _c_float, _c_ll = ctypes.c_int, ctypes.c_void_p, ctypes.c_float, ctypes.c_longlong

# name -> argtypes; every function returns int (hg_status_t) unless listed in _RESTYPES
SIGNATURES = {
    "hg_abi_version": [],
    "hg_last_error": [],
    "hg_set_option": [ctypes.c_char_p, _c_int],
    "hg_get_option": [ctypes.c_char_p, _c_void_p],
    "hg_rotate_fwd": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int,
                      _c_int, _c_int, _c_void_p],
    "hg_rotate_bwd_workspace_bytes": [_c_int, _c_int, _c_int],
    "hg_rotate_bwd": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_ll, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                      _c_int, _c_void_p],
    "hg_adain_cl_workspace_bytes": [_c_int, _c_int, _c_int, _c_int, _c_int],
    "hg_adain_cl_fwd": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_ll, _c_int, _c_int,
                        _c_int, _c_int, _c_int, _c_int, _c_float, _c_float, _c_int, _c_void_p],
    "hg_adain_cl_bwd": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                        _c_void_p, _c_ll, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_float, _c_int, _c_void_p],
    "hg_adain_act_fwd": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int,
                         _c_ll, _c_int, _c_float, _c_float, _c_int, _c_int, _c_void_p],
    "hg_adain_act_bwd": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                         _c_void_p, _c_int, _c_int, _c_int, _c_ll, _c_int, _c_int, _c_float, _c_int, _c_int, _c_void_p],
    "hg_convt_pack_weight": [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p],
    "hg_convt_fwd": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_float,
                     _c_void_p],
    "hg_convt_stats_floats": [_c_int, _c_int, _c_int, _c_int, _c_int],
    "hg_convt_fwd_stats": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                           _c_float, _c_void_p],
    "hg_adain_cl_fwd_stats": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int,
                              _c_int, _c_int, _c_int, _c_float, _c_float, _c_int, _c_void_p],
    "hg_convt_dgrad": [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p],
    "hg_convt_wgrad_workspace_bytes": [_c_int, _c_int, _c_int, _c_int, _c_int, _c_int],
    "hg_convt_wgrad": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_ll, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                       _c_int, _c_int, _c_int, _c_void_p],
    "hg_act_bwd_bias_workspace_bytes": [_c_ll, _c_int],
    "hg_act_bwd_bias": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_ll, _c_ll, _c_int, _c_float, _c_void_p],
    "hg_gemm_bf16_nt": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_ll, _c_float, _c_void_p],
    "hg_linear_relu_fwd": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p],
    "hg_linear_relu_bwd": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int,
                           _c_int, _c_int, _c_void_p],
    "hg_linear_relu_group_fwd": [_c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p],
    "hg_linear_relu_group_bwd": [_c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int,
                                 _c_void_p],
    "hg_final_conv_tanh_fwd": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p],
    "hg_final_conv_tanh_bwd_workspace_bytes": [_c_int, _c_int, _c_int, _c_int],
    "hg_final_conv_tanh_bwd": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                               _c_ll, _c_int, _c_int, _c_int, _c_int, _c_void_p],
    "hg_conv5s2_workspace_bytes": [_c_int, _c_int, _c_int, _c_int],
    "hg_conv5s2_pack_weight": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p],
    "hg_conv5s2_fwd": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_ll, _c_int, _c_int, _c_int, _c_int, _c_void_p],
    "hg_conv5s2_dx": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_ll, _c_int, _c_int, _c_int, _c_int, _c_void_p],
    "hg_conv5s2_dw": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_ll, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p],
    "hg_dconv0_fwd": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_float, _c_void_p],
    "hg_dconv0_bwd_workspace_bytes": [_c_int, _c_int],
    "hg_dconv0_bwd": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_ll, _c_int,
                      _c_int, _c_int, _c_int, _c_float, _c_int, _c_void_p],
    "hg_dheads_workspace_bytes": [_c_int, _c_int, _c_int, _c_int],
    "hg_dheads_fwd": [_c_void_p] * 11 + [_c_ll, _c_int, _c_int, _c_int, _c_int, _c_float, _c_void_p],
    "hg_dheads_bwd": [_c_void_p] * 16 + [_c_ll, _c_int, _c_int, _c_int, _c_int, _c_float, _c_void_p],
    "hg_head128_fwd": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p],
    "hg_head128_bwd_workspace_bytes": [_c_int, _c_int],
    "hg_head128_bwd": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_ll, _c_int,
                       _c_int, _c_int, _c_int, _c_void_p],
    "hg_adam_step": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_ll, _c_void_p, _c_void_p, _c_float, _c_float, _c_float,
                     _c_float, _c_void_p],
    "hg_adam_tick": [_c_void_p, _c_void_p, _c_float, _c_float, _c_void_p],
    "hg_adam_apply": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_ll, _c_void_p, _c_float, _c_float, _c_float, _c_float, _c_void_p],
    "hg_gan_loss_fwd": [_c_void_p, _c_int, _c_float, _c_float, _c_void_p, _c_int, _c_float, _c_float, _c_void_p, _c_void_p,
                        _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p],
    "hg_gan_loss_bwd": [_c_void_p, _c_void_p, _c_int, _c_float, _c_float, _c_void_p, _c_int, _c_float, _c_float, _c_void_p,
                        _c_void_p, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p],
    "hg_spectral_norm_state_floats": [_c_int, _c_int, _c_int],
    "hg_spectral_norm_workspace_bytes": [_c_int, _c_void_p, _c_void_p, _c_void_p],
    "hg_spectral_norm_fwd": [_c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                             _c_int, _c_int, _c_float, _c_int, _c_void_p, _c_ll, _c_void_p],
    "hg_spectral_norm_bwd": [_c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int,
                             _c_int, _c_void_p, _c_ll, _c_void_p],
}
_RESTYPES = {"hg_last_error": ctypes.c_char_p, "hg_rotate_bwd_workspace_bytes": ctypes.c_longlong,
             "hg_convt_wgrad_workspace_bytes": ctypes.c_longlong, "hg_act_bwd_bias_workspace_bytes": ctypes.c_longlong,
             "hg_adain_cl_workspace_bytes": ctypes.c_longlong,
             "hg_final_conv_tanh_bwd_workspace_bytes": ctypes.c_longlong,
             "hg_spectral_norm_state_floats": ctypes.c_longlong, "hg_spectral_norm_workspace_bytes": ctypes.c_longlong,
             "hg_convt_stats_floats": ctypes.c_longlong, "hg_conv5s2_workspace_bytes": ctypes.c_longlong, "hg_dconv0_bwd_workspace_bytes": ctypes.c_longlong,
             "hg_head128_bwd_workspace_bytes": ctypes.c_longlong,
             "hg_dheads_workspace_bytes": ctypes.c_longlong}

_lib = None
_lock = threading.Lock()
launch_count = 0      # number of ABI calls that enqueued kernels (bench.py reports it as gpu_launches)


class HologanB200Error(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise HologanB200Error(
                f"{LIB_PATH} not found: build it with `python -m lightning_gan_zoo_b200.build` "
                "(there is no CPU / PyTorch fallback for the HoloGAN hot path)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)       # AttributeError if the .so is stale
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, _c_int)
        if lib.hg_abi_version() != 2:
            raise HologanB200Error("libhologan_b200.so ABI version mismatch; rebuild")
        _lib = lib
    return _lib


def call(name: str, *args) -> None:
    """Invoke an ABI function and raise HologanB200Error with hg_last_error() on failure."""
    global launch_count
    lib = load()
    rc = getattr(lib, name)(*args)
    launch_count += 1
    if rc != 0:
        msg = lib.hg_last_error()
        raise HologanB200Error(f"{name} failed ({rc}): {msg.decode() if msg else ''}")


def set_option(name: str, value: int) -> int:
    """Set a tuning option of the library (include/hologan_b200.h: hg_set_option); returns the previous value."""
    lib = load()
    old = ctypes.c_int(0)
    if lib.hg_get_option(name.encode(), ctypes.byref(old)) != 0 or lib.hg_set_option(name.encode(), int(value)) != 0:
        raise HologanB200Error(lib.hg_last_error().decode())
    return old.value


def get_option(name: str) -> int:
    lib = load()
    v = ctypes.c_int(0)
    if lib.hg_get_option(name.encode(), ctypes.byref(v)) != 0:
        raise HologanB200Error(lib.hg_last_error().decode())
    return v.value
