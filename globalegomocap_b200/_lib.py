"""ctypes binding of libgem_b200.so (include/gem_b200.h).

There is no CPU fallback: if the shared library has not been built, or the call
fails, this raises.  Build with ``python -m globalegomocap_b200.build``.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GEM_B200_LIB") or os.path.join(HERE, "libgem_b200.so")     # (env: a build variant under test)


class GemError(RuntimeError):
    pass


class EnergyWeights(C.Structure):
    _fields_ = [("w3d", C.c_float), ("smooth", C.c_float), ("bone", C.c_float), ("vae", C.c_float),
                ("reproj", C.c_float)]


class LbfgsParams(C.Structure):
    _fields_ = [("lr", C.c_double), ("max_iter", C.c_int32), ("max_eval", C.c_int32),
                ("tolerance_grad", C.c_double), ("tolerance_change", C.c_double)]


class Layer(C.Structure):
    _fields_ = [("w_d", C.c_void_p), ("bias_d", C.c_void_p), ("taps", C.c_int32), ("k", C.c_int32), ("n", C.c_int32)]


class VaeWeights(C.Structure):
    _fields_ = [("dec", Layer * 6), ("dec_bwd", Layer * 6), ("enc", Layer * 6)]


_P = C.c_void_p
_I = C.c_int
_SIGNATURES = {
    "gem_version": (C.c_int, []),
    "gem_last_error": (C.c_char_p, []),
    "gem_ctx_create": (C.c_int, [C.POINTER(_P), _I, _I, _I, _I, _I, _I, _I, _I]),
    "gem_ctx_destroy": (C.c_int, [_P]),
    "gem_ctx_set_camera": (C.c_int, [_P, C.POINTER(C.c_double), _I, C.c_double, C.c_double]),
    "gem_ctx_set_skeleton": (C.c_int, [_P, C.POINTER(C.c_int32), _I]),
    "gem_ctx_set_vae": (C.c_int, [_P, _I, C.POINTER(VaeWeights)]),
    "gem_ctx_scratch_bytes": (C.c_int64, [_P]),
    "gem_ctx_set_gemm_mode": (C.c_int, [_P, _I]),
    "gem_ctx_set_chunks": (C.c_int, [_P, _I]),
    "gem_ctx_set_slices": (C.c_int, [_P, _I, C.POINTER(C.c_int32)]),
    "gem_ctx_set_ready_events": (C.c_int, [_P, _I, C.POINTER(C.c_int32), C.POINTER(C.c_void_p)]),
    "gem_solve_windows": (C.c_int, [_P, _P, _I, _P, _P, _P, _P, _P, _P, _P, C.POINTER(EnergyWeights),
                                    C.POINTER(EnergyWeights), C.POINTER(LbfgsParams), _P, _P, _P, _P, _P, _P, _P]),
    "gem_ctx_set_texel_cache": (C.c_int, [_P, _I]),
    "gem_ctx_set_heat_layout": (C.c_int, [_P, _I]),
    "gem_ctx_texel_cache_stats": (C.c_int, [_P, _I, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "gem_ctx_launch_count": (C.c_int64, [_P]),
    "gem_ctx_set_profiling": (C.c_int, [_P, _I]),
    "gem_ctx_read_profile": (C.c_int, [_P, _I, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_float),
                                        C.POINTER(C.c_int32)]),
    "gem_energy_grad": (C.c_int, [_P, _P, _I, _P, _P, _P, _P, _P, _P, C.POINTER(EnergyWeights), _P, _P, _P, _P]),
    "gem_gemm": (C.c_int, [_P, _P, _I, _I, _I, _P, _I, _P, _P, _I, _P, _I, _I]),
    "gem_decode": (C.c_int, [_P, _P, _I, _I, _P, _P]),
    "gem_decode_vjp": (C.c_int, [_P, _P, _I, _I, _P, _P]),
    "gem_encode": (C.c_int, [_P, _P, _I, _I, _P, _P, _P, _P, _P]),
    "gem_lbfgs_begin": (C.c_int, [_P, _P, _I, _P, C.POINTER(LbfgsParams)]),
    "gem_lbfgs_trial": (_P, [_P]),
    "gem_lbfgs_advance": (C.c_int, [_P, _P, _I, _P, _P]),
    "gem_lbfgs_x": (_P, [_P]),
    "gem_lbfgs_stats": (C.c_int, [_P, _P, _I, _P, _P, _P, _P, _P, _P, _I]),
    "gem_solve_stage": (C.c_int, [_P, _P, _I, _I, _P, _P, _P, _P, _P, _P, C.POINTER(EnergyWeights),
                                  C.POINTER(LbfgsParams), _P, _P, _P, _P, _P]),
    "gem_relative_global": (C.c_int, [_P, _P, _I, _P, _I, _P, _P, _P]),
    "gem_to_global": (C.c_int, [_P, _P, _I, _P, _I, _P, _P]),
    "gem_merge_windows": (C.c_int, [_P, _P, _I, _I, _P, _P]),
    "gem_gaussian_smooth": (C.c_int, [_P, _P, _I, _I, C.c_double, _P, _P]),
    "gem_pose_align_errors": (C.c_int, [_P, _I, _I, _P, _P, C.POINTER(C.c_int32), C.POINTER(C.c_double), _P, _P, _P]),
    "gem_lift_skeleton": (C.c_int, [_P, _I, _I, _I, _I, _P, _P, C.POINTER(C.c_double), _I, C.c_double, C.c_double, _I, _I,
                                    _P, _P, _P, _P]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)
_lib = None


def load():
    """Loads the library once; raises GemError if it is missing (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GemError(f"{LIB_PATH} not found: build it with `python -m globalegomocap_b200.build` "
                       "(there is no CPU fallback for this path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().gem_last_error()
        raise GemError(f"gem_b200 error {rc}: {msg.decode() if msg else ''}")
