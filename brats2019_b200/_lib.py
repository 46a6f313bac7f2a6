"""ctypes binding of libbrats_b200.so (the C ABI declared in include/brats_b200.h).

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is
raised.  Build the library with `python -c "import __graft_entry__ as g; g.build()"` or
`python -m brats2019_b200.build`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200_LIB_PATH") or os.path.join(_HERE, "libbrats_b200.so")   # override: A/B perf tests only

c_void_p, c_int, c_float, c_size_t, c_ll = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_longlong


class ConvDesc(C.Structure):
    """struct b200_conv_desc"""
    _fields_ = [("mode", c_int), ("epi", c_int), ("N", c_int), ("D", c_int), ("H", c_int), ("W", c_int),
                ("Cin_a", c_int), ("Cin_b", c_int), ("Cout", c_int)]


class WgradDesc(C.Structure):
    """struct b200_wgrad_desc"""
    _fields_ = [("mode", c_int), ("N", c_int), ("D", c_int), ("H", c_int), ("W", c_int), ("Cout", c_int),
                ("Cin", c_int)]


class PackJob(C.Structure):
    """struct b200_pack_job"""
    _fields_ = [("desc", ConvDesc), ("kind", c_int), ("Cout_w", c_int), ("Cin_w", c_int), ("taps_w", c_int),
                ("ci_off", c_int), ("K_real", c_int), ("N_real", c_int), ("w", c_void_p), ("packed", c_void_p)]


P = c_void_p
_PROTOS = {
    # name: (restype, argtypes)
    "b200_last_error": (C.c_char_p, []),
    "b200_device_check": (c_int, [c_int]),
    "b200_num_sms": (c_int, []),
    "b200_act_guard_rows": (c_ll, [c_int, c_int, c_int]),
    "b200_act_plane_rows": (c_ll, [c_int, c_int, c_int, c_int]),
    "b200_pack_input": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "b200_pack_input_t": (c_int, [P, c_int, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "b200_conv_packed_weight_bytes": (c_size_t, [C.POINTER(ConvDesc)]),
    "b200_conv_ctas": (c_int, [C.POINTER(ConvDesc)]),
    "b200_conv_pack_weight": (c_int, [C.POINTER(ConvDesc), c_int, P, c_int, c_int, c_int, c_int, c_int, c_int, P, P]),
    "b200_pack_table_entry_bytes": (c_size_t, []),
    "b200_pack_table_build": (c_int, [C.POINTER(PackJob), c_int, P, c_size_t, C.POINTER(c_int)]),
    "b200_pack_table_run": (c_int, [P, c_int, c_int, P]),
    "b200_conv_run": (c_int, [C.POINTER(ConvDesc), P, P, P, P, P, c_int, P, P, P, P, c_int, P]),
    "b200_conv_supports_gnbwd": (c_int, [C.POINTER(ConvDesc)]),
    "b200_conv_run_gnbwd": (c_int, [C.POINTER(ConvDesc), P, P, P, P, P, c_int, P, P, P, P, c_int, P, P, P]),
    "b200_conv_run_gn": (c_int, [C.POINTER(ConvDesc), P, P, P, P, P, P, P, c_float, P]),
    "b200_wgrad_workspace_bytes": (c_size_t, [C.POINTER(WgradDesc)]),
    "b200_wgrad_run": (c_int, [C.POINTER(WgradDesc), P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "b200_gn_finalize": (c_int, [P, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P, P, P]),
    "b200_gn_finalize_coef": (c_int, [P, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P, P, c_int, P, P, P, P]),
    "b200_gn_apply": (c_int, [P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "b200_gn_backward_workspace_floats": (c_size_t, [c_int, c_int]),
    "b200_gn_backward_form": (c_int, [c_int, c_int, c_int, c_int, c_int]),
    "b200_gn_backward": (c_int, [P, P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "b200_gn_backward_folded_workspace_floats": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "b200_gn_backward_folded": (c_int, [P, P, P, P, P, P, P, c_int, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "b200_upsample2x": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "b200_upsample2x_backward_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "b200_upsample2x_backward": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "b200_space_to_depth": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "b200_depth_to_space": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "b200_add": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "b200_sigmoid_backward_workspace_floats": (c_size_t, [c_int, c_int, c_int]),
    "b200_sigmoid_backward": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "b200_dice_workspace_floats": (c_size_t, [c_int, c_int]),
    "b200_dice_sums": (c_int, [P, P, P, P, c_int, c_int, c_ll, P]),
    "b200_dice_sums_t": (c_int, [P, P, c_int, P, P, c_int, c_int, c_ll, P]),
    "b200_dice_backward_t": (c_int, [P, P, c_int, P, P, c_float, P, c_int, c_int, c_ll, P]),
    "b200_bce_sum_t": (c_int, [P, P, c_int, c_float, P, P, c_ll, P]),
    "b200_bce_backward_t": (c_int, [P, P, c_int, P, c_float, C.c_double, P, c_ll, P]),
    "b200_dice_loss": (c_int, [P, c_int, c_float, P, P]),
    "b200_dice_backward": (c_int, [P, P, P, P, c_float, P, c_int, c_int, c_ll, P]),
    "b200_bce_workspace_floats": (c_size_t, []),
    "b200_bce_sum": (c_int, [P, P, c_float, P, P, c_ll, P]),
    "b200_bce_loss": (c_int, [P, C.c_double, P, P]),
    "b200_bce_backward": (c_int, [P, P, P, c_float, C.c_double, P, c_ll, P]),
    "b200_adam_step": (c_int, [P, P, P, P, P, C.POINTER(c_ll), C.POINTER(c_ll), c_int, P, P, P, C.c_double, C.c_double, c_float,
                               c_float, c_int, c_float, P]),
    "b200_conv_plan_debug": (c_int, [C.POINTER(ConvDesc), C.POINTER(c_int), c_int]),
    "b200_wgrad_plan_debug": (c_int, [C.POINTER(WgradDesc), C.POINTER(c_int), c_int]),
    "b200_wgrad_line_plan_debug": (c_int, [C.POINTER(WgradDesc), C.POINTER(c_int), c_int]),
    "b200_wgrad_line_plan_debug2": (c_int, [C.POINTER(WgradDesc), c_int, c_int, C.POINTER(c_int), c_int]),
    "b200_wgrad_march_plan_debug": (c_int, [C.POINTER(WgradDesc), C.POINTER(c_int), c_int]),
    "b200_march_plan_debug": (c_int, [C.POINTER(ConvDesc), C.POINTER(c_int), c_int]),
    "b200_band_plan_debug": (c_int, [C.POINTER(ConvDesc), C.POINTER(c_int), c_int]),
    "b200_march_prof_read": (c_int, [C.POINTER(C.c_ulonglong), c_int]),
}
EXPORTED_SYMBOLS = tuple(_PROTOS.keys())

_lib = None


def lib():
    """Load (once) and return the ctypes handle; raises if the extension was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "brats2019_b200: %s not found - the CUDA extension is not built and there is no fallback. "
                "Run `python -c 'import __graft_entry__ as g; g.build()'`." % LIB_PATH)
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(h, name)       # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def check(status, what):
    if status != 0:
        raise RuntimeError("%s failed: %s" % (what, lib().b200_last_error().decode()))
