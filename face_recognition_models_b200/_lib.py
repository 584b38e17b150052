"""ctypes binding of libmargin_head.so (C ABI declared in include/margin_head.h).

There is no CPU fallback: if the shared library is missing or the device is not sm_100, every
entry point raises.  The library is built in-tree by ``__graft_entry__.build()`` (or
``make -C face_recognition_models_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmargin_head.so")

_lib: Optional[C.CDLL] = None


class MhConfig(C.Structure):
    """Mirror of ``mh_config`` in include/margin_head.h."""
    _fields_ = [
        ("family", C.c_int32), ("easy_margin", C.c_int32), ("sphere_m", C.c_int32), ("plus", C.c_int32),
        ("s", C.c_float), ("m", C.c_float), ("mv_weight", C.c_float), ("momentum", C.c_float),
        ("h", C.c_float), ("t_alpha", C.c_float),
        ("l_margin", C.c_float), ("u_margin", C.c_float), ("l_a", C.c_float), ("u_a", C.c_float),
        ("sphere_lambda", C.c_float), ("reserved", C.c_float),
    ]


class MhStepWs(C.Structure):
    """Mirror of ``mh_step_ws`` in include/margin_head.h (workspace descriptor of mh_step_forward / mh_step_backward)."""
    _fields_ = [
        ("B", C.c_int64), ("B_pad", C.c_int64), ("C", C.c_int64), ("C_pad", C.c_int64), ("ld", C.c_int64),
        ("layout", C.c_int32), ("x_dtype", C.c_int32),
        ("w_hat", C.c_void_p), ("inv_norm", C.c_void_p), ("x_hat", C.c_void_p), ("x_hat32", C.c_void_p),
        ("xnorm", C.c_void_p), ("t_raw", C.c_void_p), ("label_local", C.c_void_p), ("rowp", C.c_void_p),
        ("stats_tiles", C.c_void_p), ("n_tiles", C.c_int64), ("merge_scratch", C.c_void_p), ("stats", C.c_void_p),
        ("rowout", C.c_void_p), ("bc", C.c_void_p), ("xs", C.c_void_p), ("rho", C.c_void_p), ("gty", C.c_void_p),
        ("dxhat_part", C.c_void_p), ("part_splits", C.c_int64), ("dxhat_full", C.c_void_p), ("gscal", C.c_void_p),
        ("r_colsum", C.c_void_p), ("rpart", C.c_void_p), ("rflag", C.c_void_p), ("dx_sync", C.c_void_p),
        ("pw_ready", C.c_void_p), ("prog", C.c_void_p), ("guard", C.c_void_p), ("graph_cache", C.c_void_p),
    ]


# enums of include/margin_head.h
FAMILY = dict(arcface=0, cosface=1, sphereface=2, mv_am=3, mv_arc=4, curricularface=5, adaface=6,
              elastic_cos=7, elastic_arc=8, magface=9, vpl_arcface=10)
LAYOUT_CD, LAYOUT_DC = 0, 1
DT_F32, DT_BF16, DT_F16 = 0, 1, 2
RP = dict(SCALE=0, THR=1, ZT=2, DZT=3, T=4, DZT_DN=5, DLG_DN=6, NORMS=7)
RP_PLANES = 8
ST_PLANES = 4
RO = dict(LSE2=0, LOSS=1, CNT=2, AUX0=3, AUX1=4)
RO_PLANES = 5
STATE_FLOATS = 8
MERGE_BLOCKS = 64
DX_SYNC_INTS = 64
TILE = 128
NTILE = 256
D = 512

_vp, _i64, _i32, _f32 = C.c_void_p, C.c_int64, C.c_int, C.c_float
_cfgp = C.POINTER(MhConfig)

# name -> argtypes; every function returns int (mh_status) unless listed in _RESTYPES
SIGNATURES = {
    "mh_device_check": [],
    "mh_prologue_w": [_vp, _i32, _i64, _i64, _vp, _i64, _vp, _vp, _vp],
    "mh_sgd_step_w": [_vp, _i32, _i64, _i64, _vp, _vp, _f32, _f32, _f32, _vp, _vp, _vp, _i64, _vp, _vp],
    "mh_prologue_x": [_vp, _i32, _i64, _i64, _vp, _vp, _i32, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp],
    "mh_row_params": [_cfgp, _i64, _vp, _vp, _vp, _vp, _i32, _vp, _i64, _vp],
    "mh_tc_forward": [_cfgp, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp],
    "mh_tc_forward_pw": [_cfgp, _vp, _i64, _i64, _vp, _i32, _i64, _vp, _vp, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp,
                         C.POINTER(C.c_int), _vp],
    "mh_tc_backward_g": [_cfgp, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp],
    "mh_tc_backward_dw_fused": [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _i32, _vp, _i32, _vp, _i64, _vp],
    "mh_tc_backward_dw_proj": [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _i32, _vp, _i64, _vp, _vp, _vp],
    "mh_tc_backward_dxdw": [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _i32, _vp, _i64, _vp, C.POINTER(C.c_int), _vp, _vp, _vp, _vp],
    "mh_tc_backward_dx_stash": [_cfgp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _i64, _vp, _vp, C.POINTER(C.c_int), _vp, _vp],
    "mh_vpl_mix": [_vp, _vp, _vp, C.c_float, _i64, _i64, _vp, _vp, _vp],
    "mh_pair_cosine": [_vp, _vp, _i32, _i64, _i64, _i64, _i64, _vp, _vp],
    "mh_step_forward": [_cfgp, C.POINTER(MhStepWs), _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp],
    "mh_step_backward": [_cfgp, C.POINTER(MhStepWs), _i32, _vp, _vp, _vp, _vp, _vp, _vp],
    "mh_step_cache_create": [C.POINTER(C.c_void_p)],
    "mh_step_cache_destroy": [_vp],
    "mh_step_cache_stats": [_vp, C.POINTER(C.c_int64)],
    "mh_tc_fixref_ok": [_cfgp, _i64],
    "mh_tc_stash_ok": [_cfgp, _i64],
    "mh_tc_stash_guarded_ok": [_cfgp, _i64],
    "mh_tc_forward_ex": [_cfgp, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _i32, _vp, _i32, _vp],
    "mh_tc_backward_g_ex": [_cfgp, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp],
    "mh_merge_stats_ex": [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _i32, _vp],
    "mh_finalize_rows_ex": [_vp, _i64, _vp, _i64, _i64, _i64, _i32, _vp, _i64, _vp, _vp, _f32, _vp, _vp, _i32, _vp],
    "mh_stash_prep_ex": [_cfgp, _vp, _i64, _vp, _i64, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp],
    "mh_stash_prep": [_cfgp, _vp, _i64, _vp, _i64, _vp, _i64, _i64, _vp, _vp, _vp, _vp],
    "mh_stash_dx_combine": [_vp, _i32, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _vp],
    "mh_stash_dw_target": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _i64, _vp],
    "mh_tc_backward_dx": [_vp, _i64, _i64, _vp, _vp, C.POINTER(C.c_int), _vp, _vp],
    "mh_tc_backward_dw": [_vp, _i64, _i64, _vp, _vp, _vp],
    "mh_sgemm_strided": [_i64, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _vp],
    "mh_dense_forward": [_cfgp, _vp, _i64, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp],
    "mh_dense_backward_dc": [_cfgp, _vp, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "mh_merge_stats": [_vp, _i64, _i64, _i64, _vp, _vp, _vp],
    "mh_finalize_rows": [_vp, _i64, _vp, _i64, _i64, _i64, _i32, _vp, _i64, _vp, _vp, _vp],
    "mh_make_gscal": [_vp, _vp, _i64, _vp, _vp],
    "mh_norm_backward_x": [_vp, _i32, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _i32, _vp],
    "mh_norm_backward_w": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _i64, _vp],
}
_RESTYPES = {"mh_version": C.c_char_p, "mh_last_error": C.c_char_p, "mh_fwd_num_tiles": C.c_int64}
EXPORTED = sorted(list(SIGNATURES) + ["mh_version", "mh_last_error", "mh_fwd_num_tiles", "mh_tc_schedule_tiles"])


class MarginHeadError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the shared library (once). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("MH_LIB") or LIB_PATH      # MH_LIB: an experimental build (scripts/build_variant.sh)
    if not os.path.exists(path):
        raise MarginHeadError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for the margin head.")
    lib = C.CDLL(path)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.mh_version.restype = C.c_char_p
    lib.mh_version.argtypes = []
    lib.mh_last_error.restype = C.c_char_p
    lib.mh_last_error.argtypes = []
    lib.mh_fwd_num_tiles.restype = C.c_int64
    lib.mh_fwd_num_tiles.argtypes = [C.c_int64]
    lib.mh_tc_schedule_tiles.restype = C.c_int64
    lib.mh_tc_schedule_tiles.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64]
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().mh_last_error().decode("utf-8", "replace")
        raise MarginHeadError(f"{what} failed (status {status}): {msg}")


# When set to a list, every call is bracketed by CUDA events on the current (launching) stream and
# (name, start_event, end_event, n_kernel_launches) is appended: bench.py's per-kernel roofline timing.
PROFILE = None


def _launches(name: str, args) -> int:
    if name in ("mh_merge_stats", "mh_merge_stats_ex"):
        return 2 if int(args[1]) >= 256 else 1
    if name in ("mh_tc_forward", "mh_tc_forward_ex"):
        return 2                      # statistics-identity fill + the tensor-core kernel
    if name == "mh_tc_backward_dx":
        return 0 if not getattr(args[4], "value", None) else 1
    if name == "mh_tc_backward_dx_stash":
        return 0 if not getattr(args[9], "value", None) else 1
    if name == "mh_tc_backward_dxdw":
        return 0 if not getattr(args[11], "value", None) else 1
    if name == "mh_step_forward":
        # x prologue, row terms, identity fill, forward, merge, finalize (+ W prologue; + the four gated launches of the
        # guarded stash's fallback: identity fill, general forward, merge, finalize)
        return 6 + (1 if int(args[8]) else 0) + (4 if int(args[9]) == 2 else 0)
    if name == "mh_step_backward":
        stash, dx, dw = int(args[2]), bool(getattr(args[6], "value", None)), bool(getattr(args[7], "value", None))
        return (1 + (1 if stash else 1) + ((3 if stash else 2) if dx else 0) + ((3 if stash else 2) if dw else 0)
                + (1 if stash == 2 else 0))
    return 1


def call(name: str, *args) -> None:
    """Call an int-returning entry point and raise MarginHeadError on a non-zero status."""
    lib = load()
    if PROFILE is None:
        check(getattr(lib, name)(*args), name)
        return
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(getattr(lib, name)(*args), name)
    e1.record()
    PROFILE.append((name, e0, e1, _launches(name, args)))
