"""ctypes binding of libpolarb200.so (include/polar_b200.h). Thin: plain pointers and sizes only.

There is no CPU fallback: if the CUDA library is missing or cannot be loaded, importing this module
raises, loudly."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
import os
LIB_PATH = Path(os.environ['PB_LIB']) if os.environ.get('PB_LIB') else PKG / 'libpolarb200.so'   # PB_LIB: developer A/B builds

PB_MAX_ROIS, PB_MAX_VARS, PB_MAX_BINS = 32, 4, 128
FLAG_OUT_F64, FLAG_NO_HIST, FLAG_EXACT_CHI2, FLAG_FORCE_UNFUSED, FLAG_PROJECT_ONLY, FLAG_HIST_MATCH, FLAG_STRIP_MARCH = 1, 2, 4, 8, 16, 32, 64
METHODS = {'1PF': 0, 'CARS': 1, 'SRS': 2, 'SHG': 3, '2PF': 4, '4POLAR 2D': 5, '4POLAR 3D': 6}
DTYPES = {'uint8': 0, 'uint16': 1, 'float32': 2, 'float64': 3}


class PbParams(C.Structure):
    _fields_ = [('method', C.c_int32), ('dtype', C.c_int32), ('N', C.c_int32), ('H', C.c_int32), ('W', C.c_int32),
                ('bin_h', C.c_int32), ('bin_w', C.c_int32), ('flags', C.c_uint32), ('dark', C.c_double),
                ('removed_intensity', C.c_double), ('rotation_stick', C.c_double), ('rotation_fig', C.c_double),
                ('reference_angle', C.c_double), ('chi2_threshold', C.c_double), ('ilow', C.c_double)]


class PbOutputs(C.Structure):
    _fields_ = [('var', C.c_void_p * PB_MAX_VARS), ('intmap', C.c_void_p), ('intensity', C.c_void_p),
                ('rho_angle', C.c_void_p), ('rho_contour', C.c_void_p), ('mask', C.c_void_p), ('hist', C.c_void_p)]


EXPORTS = {
    'pb_abi_version': (C.c_int, []),
    'pb_create': (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    'pb_destroy': (C.c_int, [C.c_void_p]),
    'pb_last_error': (C.c_char_p, [C.c_void_p]),
    'pb_set_diskcone': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    'pb_set_invk': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    'pb_set_basis': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'pb_set_hist': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    'pb_set_rois': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    'pb_analyze': (C.c_int, [C.c_void_p, C.POINTER(PbParams), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(PbOutputs), C.c_void_p]),
    'pb_analyze_host': (C.c_int, [C.c_void_p, C.POINTER(PbParams), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(PbOutputs)]),
    'pb_get_intensity': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    'pb_field': (C.c_int, [C.c_void_p, C.POINTER(PbParams), C.c_void_p, C.c_void_p, C.c_void_p]),
    'pb_binning': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'pb_compute_fields': (C.c_int, [C.c_void_p, C.POINTER(PbParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(PbOutputs), C.c_void_p]),
    'pb_fit_maps': (C.c_int, [C.c_void_p, C.POINTER(PbParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'pb_histo': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_long, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    'pb_stats': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_long, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p]),
    'pb_stats_batch': (C.c_int, [C.c_void_p, C.POINTER(PbParams), C.c_int, C.c_int, C.POINTER(PbOutputs), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'pb_set_diskcones': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    'pb_select_diskcone': (C.c_int, [C.c_void_p, C.c_int]),
    'pb_compute_dark': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_void_p]),
    'pb_lut_apply': (C.c_int, [C.c_void_p, C.POINTER(PbParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(PbOutputs), C.c_void_p]),
    'pb_stack_from_hwn': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    'pb_rescale_stack': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_long, C.c_void_p, C.POINTER(C.c_double), C.c_void_p]),
    'pb_compact': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.POINTER(C.c_long), C.c_void_p]),
    'pb_launch_count': (C.c_long, [C.c_void_p]),
    'pb_plan': (C.c_char_p, [C.c_void_p, C.POINTER(PbParams)]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen libpolarb200.so and type every export of include/polar_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(f'{LIB_PATH} is missing: build it with `python -m polarimetry_b200.build` '
                          '(nvcc, sm_100a). polarimetry_b200 has no CPU or PyTorch fallback.')
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export it
        fn.restype, fn.argtypes = res, args
    if lib.pb_abi_version() != 1:
        raise ImportError('libpolarb200.so ABI version mismatch')
    _lib = lib
    return lib


class PbError(RuntimeError):
    pass


def check(lib, ctx, rc: int, what: str):
    if rc != 0:
        msg = lib.pb_last_error(ctx)
        raise PbError(f'{what} failed ({rc}): {msg.decode() if msg else ""}')
