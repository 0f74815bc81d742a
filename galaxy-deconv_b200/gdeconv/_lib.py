"""ctypes binding of libgdeconv.so (C ABI declared in include/gdeconv.h).

There is no fallback: if the shared library is missing it is built with nvcc (gdeconv/build.py); if that fails
the import fails.  Nothing here or above it computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

ARCH_G, ARCH_U = 0, 1
PREC_FP32_SIMT, PREC_FP16_UMMA, PREC_FP16_SIMT = 0, 1, 2
LLH_GAUSSIAN, LLH_POISSON = 0, 1
SOLVER_RL, SOLVER_WIENER, SOLVER_TIKHONOV_ID, SOLVER_TIKHONOV_LAP = 0, 1, 2, 3
PRECISIONS = {'fp32_simt': PREC_FP32_SIMT, 'fp16_umma': PREC_FP16_UMMA, 'fp16_simt': PREC_FP16_SIMT}

#: every symbol include/gdeconv.h declares (tests/test_abi.py checks the header against this list and the .so)
SYMBOLS = ('gd_version', 'gd_last_error', 'gd_pack_weights', 'gd_free_weights', 'gd_workspace_bytes', 'gd_workspace_init',
           'gd_admm_forward', 'gd_resunet_forward', 'gd_subnet_forward', 'gd_fft_solver', 'gd_conv_fft', 'gd_moments_e',
           'gd_launch_count', 'gd_debug_geom', 'gd_debug_tapgemm', 'gd_profile_begin', 'gd_profile_end', 'gd_pack_xdense', 'gd_free_xdense', 'gd_xdense_workspace_bytes',
           'gd_xdense_forward', 'gd_tikhonet_forward')


class GdTensorDesc(C.Structure):
    _fields_ = [('name', C.c_char_p), ('data', C.c_void_p), ('ndim', C.c_int), ('shape', C.c_int64 * 4)]


def _load():
    path = _build.LIB
    if not os.path.exists(path) or (os.environ.get('GDECONV_REBUILD') and not _build.is_current()):
        _build.build()
    lib = C.CDLL(path)
    vp, i, f, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    lib.gd_version.restype = i
    lib.gd_last_error.restype = C.c_char_p
    lib.gd_launch_count.restype = C.c_uint64
    lib.gd_pack_weights.argtypes = [i, i, C.POINTER(GdTensorDesc), i, i, i, C.POINTER(vp)]
    lib.gd_free_weights.argtypes = [vp]
    lib.gd_free_weights.restype = None
    lib.gd_workspace_bytes.argtypes = [i, i, i]
    lib.gd_workspace_bytes.restype = sz
    lib.gd_workspace_init.argtypes = [vp, sz, i, i, i, vp]
    lib.gd_admm_forward.argtypes = [vp, i, i, vp, vp, vp, vp, vp, vp, i, vp, sz, vp]
    lib.gd_resunet_forward.argtypes = [vp, vp, vp, i, vp, sz, vp]
    lib.gd_subnet_forward.argtypes = [vp, vp, vp, vp, i, vp]
    lib.gd_fft_solver.argtypes = [i, i, f, vp, vp, vp, vp, i, vp]
    lib.gd_conv_fft.argtypes = [vp, vp, vp, i, i, vp]
    lib.gd_moments_e.argtypes = [vp, vp, i, vp]
    lib.gd_profile_begin.restype = None
    lib.gd_pack_xdense.argtypes = [C.POINTER(GdTensorDesc), i, C.c_char_p, i, C.POINTER(vp)]
    lib.gd_free_xdense.argtypes = [vp]
    lib.gd_free_xdense.restype = None
    lib.gd_xdense_workspace_bytes.argtypes = [i]
    lib.gd_xdense_workspace_bytes.restype = sz
    lib.gd_xdense_forward.argtypes = [vp, vp, vp, i, vp, sz, i, vp]
    lib.gd_tikhonet_forward.argtypes = [vp, i, f, vp, vp, vp, vp, i, vp, sz, i, vp]
    lib.gd_profile_end.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    lib.gd_debug_geom.argtypes = [i, i, C.POINTER(C.c_int * 7)]
    lib.gd_debug_tapgemm.argtypes = [i, i, i, i, i, i, i, vp, vp, vp, vp]
    for name in SYMBOLS:
        getattr(lib, name)          # AttributeError here = the .so does not match the header
    return lib


lib = _load()
LIB_PATH = _build.LIB


def check(rc):
    if rc != 0:
        raise RuntimeError(f'libgdeconv error {rc}: {lib.gd_last_error().decode(errors="replace")}')
