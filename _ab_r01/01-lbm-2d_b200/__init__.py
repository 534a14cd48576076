"""B200-native D2Q9 MRT-LES lattice-Boltzmann time step (drop-in for the hot path of
ms-112-scott/01-lbm-2d).  The directory name is not a Python identifier; import it with

    import importlib; lbm = importlib.import_module("01-lbm-2d_b200")

Contents: csrc/ (sm_100a kernels + the C ABI of include/lbm2d.h), _capi.py (ctypes binding),
solver.py (`LBM2D_MRT_LES`, the reference's solver class API).
"""
from ._build import LIB_PATH, build_library  # noqa: F401
from ._capi import LbmError, load as load_library  # noqa: F401
from .solver import LBM2D_MRT_LES  # noqa: F401

__all__ = ["LBM2D_MRT_LES", "build_library", "load_library", "LbmError", "LIB_PATH"]
