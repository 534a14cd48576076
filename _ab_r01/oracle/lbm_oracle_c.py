"""ctypes front-end of oracle/lbm_oracle.c  --  TEST INFRASTRUCTURE ONLY (see the C file header).

``OracleLBMC`` has the same surface as ``OracleLBM`` (and so as the reference class) but runs
the three passes in C with OpenMP; parameter parsing is inherited from the numpy oracle so
both restatements derive their constants from one place.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .lbm_oracle_np import OracleLBM

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liblbm_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C oracle with the committed Makefile (gcc, strict IEEE flags)."""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
        os.path.join(_HERE, "lbm_oracle.c")
    ):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


def _params_struct(ctype):
    class P(C.Structure):
        _fields_ = [
            ("nx", C.c_int32), ("ny", C.c_int32), ("warmup_steps", C.c_int32), ("les_on", C.c_int32),
            ("bc_type", C.c_int32 * 4), ("frame_count", C.c_int32),
            ("tau0", ctype), ("tau0_sq", ctype), ("cs_sq_factor", ctype),
            ("rho_in", ctype), ("rho_out", ctype), ("bc_value", (ctype * 2) * 4),
            ("S_base", ctype * 9), ("w", ctype * 9), ("invM", ctype * 81),
        ]

    return P


_P32 = _params_struct(C.c_float)
_P64 = _params_struct(C.c_double)


def load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_ramp_f32.restype = C.c_float
        _lib.oracle_ramp_f64.restype = C.c_double
        _lib.oracle_max_velocity_f32.restype = C.c_float
        _lib.oracle_max_velocity_f64.restype = C.c_double
    return _lib


def set_threads(n: int) -> None:
    """OpenMP thread count for the following calls (libgomp honours omp_set_num_threads)."""
    gomp = C.CDLL("libgomp.so.1")
    gomp.omp_set_num_threads(int(n))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleLBMC(OracleLBM):
    def __init__(self, config, mask_data=None, dtype=np.float32, exact_inv_m=False):
        super().__init__(config, mask_data=mask_data, dtype=dtype, exact_inv_m=exact_inv_m)
        self._lib = load()
        F = self.F
        self._suf = "_f32" if F is np.float32 else "_f64"
        p = (_P32 if F is np.float32 else _P64)()
        p.nx, p.ny = self.nx, self.ny
        p.warmup_steps = int(self.warmup_steps)
        p.les_on = int(self.C_smag > 0.001)
        p.frame_count = 0
        for d in range(4):
            p.bc_type[d] = int(self.bc_type[d])
            p.bc_value[d][0] = float(self.bc_value[d, 0])
            p.bc_value[d][1] = float(self.bc_value[d, 1])
        p.tau0 = float(F(self.tau_0))
        p.tau0_sq = float(F(self.tau_0**2))
        p.cs_sq_factor = float(F(self.Cs_sq_factor))
        p.rho_in = float(F(self.rho_in_target))
        p.rho_out = float(F(self.rho_out_target))
        for k in range(9):
            p.S_base[k] = float(self.S_base[k])
            p.w[k] = float(self.w[k])
        flat = np.ascontiguousarray(self.invM, dtype=F).ravel()
        for k in range(81):
            p.invM[k] = float(flat[k])
        self._p = p
        self.mask = np.ascontiguousarray(self.mask, dtype=np.float32)

    def _fn(self, name):
        return getattr(self._lib, name + self._suf)

    @property
    def frame_count(self):
        return int(self._p.frame_count) if hasattr(self, "_p") else 0

    @frame_count.setter
    def frame_count(self, v):
        if hasattr(self, "_p"):
            self._p.frame_count = int(v)

    def init(self):
        self._fn("oracle_init")(C.byref(self._p), _ptr(self.f_old), _ptr(self.f_new), _ptr(self.rho), _ptr(self.vel))

    def collide_and_stream(self):
        self._fn("oracle_collide_and_stream")(
            C.byref(self._p), _ptr(self.f_old), _ptr(self.f_new), _ptr(self.damp_x), _ptr(self.damp_y))

    def update_macro_var(self):
        self._fn("oracle_update_macro_var")(
            C.byref(self._p), _ptr(self.f_old), _ptr(self.f_new), _ptr(self.rho), _ptr(self.vel))

    def apply_bc(self):
        self._fn("oracle_apply_bc")(C.byref(self._p), _ptr(self.f_old), _ptr(self.rho), _ptr(self.vel), _ptr(self.mask))

    def run_step(self, steps=1):
        self._fn("oracle_run_step")(
            C.byref(self._p), _ptr(self.f_old), _ptr(self.f_new), _ptr(self.rho), _ptr(self.vel),
            _ptr(self.mask), _ptr(self.damp_x), _ptr(self.damp_y), C.c_int(int(steps)))

    def get_force(self):
        out = np.zeros(2, self.F)
        self._fn("oracle_force")(C.byref(self._p), _ptr(self.f_new), _ptr(self.mask), _ptr(out))
        self.force_sum[:] = out
        return out

    def get_max_velocity(self):
        return float(self._fn("oracle_max_velocity")(C.byref(self._p), _ptr(self.vel)))

    def compute_moments_for_output(self):
        self._fn("oracle_moments")(C.byref(self._p), _ptr(self.f_new), _ptr(self.moments_field))
