"""ctypes binding of include/lbm2d.h -- the reference-side stub a maintainer would add
(see INTEGRATION.md).  Fails loudly when the CUDA library is missing: there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

from . import _build


class LbmParams(C.Structure):
    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32), ("warmup_steps", C.c_int32),
        ("nu", C.c_double), ("rho_in", C.c_double), ("rho_out", C.c_double),
        ("c_smag", C.c_double), ("s_ghost", C.c_double),
        ("sponge_in", C.c_int32), ("sponge_out", C.c_int32), ("sponge_top", C.c_int32), ("sponge_bot", C.c_int32),
        ("sponge_strength", C.c_double),
        ("bc_type", C.c_int32 * 4), ("bc_value", (C.c_float * 2) * 4),
        ("arith", C.c_int32), ("obstacle_mode", C.c_int32), ("device", C.c_int32), ("kernel", C.c_int32),
        ("nx_global", C.c_int32), ("slab_x0", C.c_int32),
    ]


class LbmExportConfig(C.Structure):
    _fields_ = [("x0", C.c_int32), ("x1", C.c_int32), ("y0", C.c_int32), ("y1", C.c_int32),
                ("target_w", C.c_int32), ("target_h", C.c_int32)]


class LbmDeviceView(C.Structure):
    _fields_ = [
        ("f_cur", C.c_void_p), ("f_prev", C.c_void_p), ("rho", C.c_void_p), ("ux", C.c_void_p), ("uy", C.c_void_p),
        ("cell_code", C.c_void_p), ("nx_local", C.c_int32), ("ny", C.c_int32), ("pitch", C.c_int32),
        ("plane_stride", C.c_int64), ("stream", C.c_void_p),
    ]


ABI_VERSION = 2  # LBM2D_ABI_VERSION of include/lbm2d.h
COMM_ID_BYTES = 128
PEER_HANDLE_BYTES = 256
ARITH = {"fast": 0, "strict": 1}
KERNEL = {"auto": 0, "register": 1, "tma": 2}
EXPORTS = {
    "lbm_abi_version": (C.c_int, []),
    "lbm_last_error": (C.c_char_p, []),
    "lbm_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "lbm_create": (C.c_int, [C.POINTER(LbmParams), C.c_void_p, C.POINTER(C.c_void_p)]),
    "lbm_destroy": (C.c_int, [C.c_void_p]),
    "lbm_init": (C.c_int, [C.c_void_p]),
    "lbm_run": (C.c_int, [C.c_void_p, C.c_int]),
    "lbm_synchronize": (C.c_int, [C.c_void_p]),
    "lbm_step_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "lbm_get_force": (C.c_int, [C.c_void_p, C.c_void_p]),
    "lbm_get_max_velocity": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "lbm_get_vel": (C.c_int, [C.c_void_p, C.c_void_p]),
    "lbm_get_rho": (C.c_int, [C.c_void_p, C.c_void_p]),
    "lbm_get_mask": (C.c_int, [C.c_void_p, C.c_void_p]),
    "lbm_get_moments": (C.c_int, [C.c_void_p, C.c_void_p]),
    "lbm_get_viz_fields": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "lbm_get_f": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "lbm_export_configure": (C.c_int, [C.c_void_p, C.POINTER(LbmExportConfig)]),
    "lbm_export_layout": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "lbm_export_frame": (C.c_int, [C.c_void_p, C.c_void_p]),
    "lbm_export_frame_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "lbm_export_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]),
    "lbm_comm_unique_id": (C.c_int, [C.c_void_p]),
    "lbm_comm_connect": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "lbm_static_mask": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                  C.POINTER(C.c_int32)]),
    "lbm_peer_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "lbm_peer_connect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "lbm_device_view": (C.c_int, [C.c_void_p, C.POINTER(LbmDeviceView)]),
    "lbm_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "lbm_graph_replay_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "lbm_selftest_arith": (C.c_int, [C.c_int64, C.c_uint64, C.POINTER(C.c_int64)]),
    "lbm_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "lbm_host_free": (C.c_int, [C.c_void_p]),
}

_lib = None


class LbmError(RuntimeError):
    pass


def load(build_if_missing: bool = True):
    """dlopen lib/liblbm2d.so (building it first if the sources are newer) and bind every export."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("LBM2D_LIB", _build.LIB_PATH)  # LBM2D_LIB: experiment builds (tuning sweeps)
    if path == _build.LIB_PATH and build_if_missing and _build.needs_build():
        _build.build_library()
    if not os.path.exists(path):
        raise LbmError(f"{path} is missing: build it with __graft_entry__.build() (no CPU fallback exists)")
    lib = C.CDLL(path)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.lbm_abi_version() != ABI_VERSION:
        raise LbmError("liblbm2d.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().lbm_last_error()
        raise LbmError(f"lbm2d error {rc}: {msg.decode() if msg else '?'}")
