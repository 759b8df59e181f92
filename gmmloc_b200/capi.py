"""ctypes binding of libgmmloc_b200.so — exactly the entry points of include/gmmloc_b200.h.

There is no CPU fallback: if the CUDA library is missing this module raises at import of the
library handle, and every compute call fails with GMMLOC_ECUDA on a box without a GPU.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

PKG = Path(__file__).resolve().parent
# GMMLOC_B200_LIB: load another build of the SAME library (tools/build_variant.sh: A/B timing of compile-time variants)
LIB_PATH = Path(os.environ.get("GMMLOC_B200_LIB") or (PKG / "lib" / "libgmmloc_b200.so"))

OK, EINVAL, ENOMEM, ECUDA, ENUMERIC, EIO = range(6)
MODE_JOINT, MODE_POSE_ONLY, MODE_POINTS_ONLY = 0, 1, 2
FRAME_NUMERIC = 1
VIEW_OVERFLOW = 1
ASSOC_AUTO, ASSOC_FFMA2, ASSOC_TENSOR, ASSOC_GRID = 0, 1, 2, 3
GN_FP64, GN_MIXED, GN_WORLD, GN_STREAM = 0, 1, 2, 3
GN_DEFAULT = GN_STREAM  # what a new context runs (gmmloc_ctx_create)
OPT_GN_PREP, OPT_E2E_CHUNK, OPT_E2E_SHORT_FIRST, OPT_BA_CLUSTER, OPT_E2E_TAPER = 0, 1, 2, 3, 4


class GridOptions(C.Structure):
    """gmmloc_grid_options_t (gmmloc_map_create_ex)."""
    _fields_ = [("enable", C.c_int32), ("levels", C.c_int32), ("cell_size", C.c_double), ("tmax", C.c_double * 3)]


class Camera(C.Structure):
    """gmmloc_camera_t (gmmloc_ros/cfg/v1.yaml:5-17)."""
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("bf", C.c_double), ("width", C.c_int32), ("height", C.c_int32)]


class Params(C.Structure):
    """gmmloc_params_t."""
    _fields_ = [("tri_lambda2", C.c_double), ("ba_lambda2", C.c_double), ("tri_str_thresh", C.c_double),
                ("tri_check_str_chi2", C.c_int32), ("mdist2_gate", C.c_double), ("chi2_mono", C.c_double),
                ("chi2_stereo", C.c_double), ("gn_iters", C.c_int32), ("robust", C.c_int32),
                ("factor_mode", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("points", C.c_int64), ("flagged", C.c_int64), ("kernel_launches", C.c_int64), ("flagged_brute", C.c_int64)]


class GmmlocError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"gmmloc status {status}: {msg}")
        self.status = status


# every symbol include/gmmloc_b200.h declares: name -> (restype, argtypes)
_vp, _i, _d, _i64 = C.c_void_p, C.c_int, C.c_double, C.c_int64
SYMBOLS = {
    "gmmloc_version": (C.c_char_p, []),
    "gmmloc_status_string": (C.c_char_p, [_i]),
    "gmmloc_last_error": (C.c_char_p, []),
    "gmmloc_default_params": (None, [C.POINTER(Params)]),
    "gmmloc_sigma2_inv": (C.c_float, [_i]),
    "gmmloc_map_create": (_i, [_vp, _vp, _i, _d, _i, C.POINTER(_vp)]),
    "gmmloc_map_create_ex": (_i, [_vp, _vp, _i, _d, _i, _vp, C.POINTER(_vp)]),
    "gmmloc_map_query_point": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "gmmloc_ctx_set_option": (_i, [_vp, _i, _i64]),
    "gmmloc_frames_solve_chunks": (C.c_int32, [C.c_int32] * 5 + [_vp, C.c_int32]),
    "gmmloc_map_load_gmm": (_i, [C.c_char_p, _d, _i, C.POINTER(_vp)]),
    "gmmloc_map_save_gmm": (_i, [_vp, C.c_char_p]),
    "gmmloc_map_destroy": (None, [_vp]),
    "gmmloc_map_size": (_i, [_vp]),
    "gmmloc_map_device": (_i, [_vp]),
    "gmmloc_map_get_components": (_i, [_vp] * 10),
    "gmmloc_map_neighbors": (_i, [_vp, _d, _vp, _vp, _vp, _i64, C.POINTER(_i64)]),
    "gmmloc_map_build_neighbors": (_i, [_vp, _d, C.POINTER(_i64)]),
    "gmmloc_map_scan_bound": (_i, [_vp, C.POINTER(_d), C.POINTER(_d)]),
    "gmmloc_ctx_create": (_i, [_i, C.POINTER(_vp)]),
    "gmmloc_ctx_destroy": (None, [_vp]),
    "gmmloc_ctx_stream": (_vp, [_vp]),
    "gmmloc_ctx_synchronize": (_i, [_vp]),
    "gmmloc_ctx_stats": (_i, [_vp, C.POINTER(Stats)]),
    "gmmloc_ctx_kernel_times": (_i, [_vp, C.POINTER(C.c_float * 3)]),
    "gmmloc_ctx_set_assoc_kernel": (_i, [_vp, _i]),
    "gmmloc_ctx_last_assoc_kernel": (_i, [_vp]),
    "gmmloc_ctx_last_gn_kernel": (_i, [_vp]),
    "gmmloc_ctx_set_gn_precision": (_i, [_vp, _i]),
    "gmmloc_host_alloc": (_i, [C.c_size_t, C.POINTER(_vp)]),
    "gmmloc_host_free": (None, [_vp]),
    "gmmloc_associate": (_i, [_vp, _vp, _vp, _i, _i, _d, _vp, _vp]),
    "gmmloc_associate_async": (_i, [_vp, _vp, _vp, _i64, _d, _vp, _vp]),
    "gmmloc_frames_solve": (_i, [_vp, _vp, C.POINTER(Camera), C.POINTER(Params)] + [_vp] * 4 + [_i, _i] + [_vp] * 5),
    "gmmloc_frames_solve_async": (_i, [_vp, _vp, C.POINTER(Camera), C.POINTER(Params)] + [_vp] * 4 + [_i, _i] + [_vp] * 5),
    "gmmloc_poses_optimize": (_i, [_vp, C.POINTER(Camera)] + [_vp] * 5 + [_i, _i] + [_vp] * 3),
    "gmmloc_triangulation_solve": (_i, [_vp, _vp, C.POINTER(Camera), C.POINTER(Params)] + [_vp] * 7 + [_i, _i] + [_vp] * 5),
    "gmmloc_render_views": (_i, [_vp, _vp, C.POINTER(Camera), _vp, _i, _i] + [_vp] * 6),
    "gmmloc_search_correspondence": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "gmmloc_check_map_association": (_i, [_vp, _vp, C.POINTER(Camera), C.POINTER(Params)] + [_vp] * 4 + [_i, _i, _i, _vp, _vp]),
    "gmmloc_keyframes_associate": (_i, [_vp, _vp, C.POINTER(Camera), C.POINTER(Params)] + [_vp] * 4 + [_i, _i, _i, _i] + [_vp] * 5),
    "gmmloc_local_ba": (_i, [_vp, _vp, C.POINTER(Camera), C.POINTER(Params), _i] + [_vp] * 17),
    "gmmloc_points_solve": (_i, [_vp, _vp, C.POINTER(Camera), C.POINTER(Params)] + [_vp] * 6 + [_i, _i] + [_vp] * 4),
}

_lib = None


def lib() -> C.CDLL:
    """Loads the CUDA library; raises (loudly) when it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m gmmloc_b200.build` "
                "(gmmloc_b200 has no CPU fallback)")
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status: int) -> None:
    if status != OK:
        raise GmmlocError(status, lib().gmmloc_last_error().decode(errors="replace"))


def default_params(**kw) -> Params:
    p = Params()
    lib().gmmloc_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def v1_camera() -> Camera:
    """Intrinsics shipped with the reference (gmmloc_ros/cfg/v1.yaml:5-17)."""
    return Camera(435.2046959714599, 435.2046959714599, 367.4517211914062, 252.2008514404297,
                  47.90639384423901, 752, 480)
