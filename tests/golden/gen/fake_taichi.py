"""A minimal pure-Python stand-in for the subset of Taichi 1.7 that the reference solver uses.

Purpose: execute the UNMODIFIED reference source
(/root/reference/src/lbm_mrt_les/core/LBM2D_MRT_LES.py) in a container where Taichi cannot be
installed, cell by cell in plain Python loops, to produce golden vectors that pin the oracle
(oracle/lbm_oracle_np.py, oracle/lbm_oracle.c) to the reference's own code rather than to a
reading of it.  Used only by tests/golden/gen/make_ti_shim_fixtures.py at fixture-generation
time; the committed .npz files are what the tests read.

Typing model (Taichi default_fp=f32, default_ip=i32):
  * field elements and vectors are np.float32; NumPy's weak-scalar promotion (NEP 50) makes a
    Python float meeting an f32 round to f32 first -- the same thing Taichi does with a
    Python-scope constant;
  * kernel loop indices are ``TiInt`` so that ``int / int`` is an f32 true division;
  * ``Vec.sum()`` / ``Vec.norm()`` / ``Mat @ Vec`` evaluate left to right in f32.
Parallel for-loops run sequentially (the reference kernels are race-free apart from the
force/max atomics, whose summation order is unspecified in Taichi anyway).
"""
from __future__ import annotations

import itertools
import math as _pymath
import sys
import types

import numpy as np

F32 = np.float32


class TiInt(int):
    """i32 loop index: arithmetic stays integral, '/' is Taichi's f32 true division."""

    def __add__(self, o):
        return TiInt(int(self) + int(o)) if isinstance(o, int) else NotImplemented

    __radd__ = __add__

    def __sub__(self, o):
        return TiInt(int(self) - int(o)) if isinstance(o, int) else NotImplemented

    def __rsub__(self, o):
        return TiInt(int(o) - int(self)) if isinstance(o, int) else NotImplemented

    def __mul__(self, o):
        return TiInt(int(self) * int(o)) if isinstance(o, int) else NotImplemented

    __rmul__ = __mul__

    def __truediv__(self, o):
        return F32(F32(int(self)) / F32(o))

    def __rtruediv__(self, o):
        return F32(F32(o) / F32(int(self)))


class Vec(np.ndarray):
    """f32 (or i32) vector with Taichi's sequential reductions."""

    def sum(self):  # noqa: A003  (sequential, starts at 0)
        s = self.dtype.type(0)
        for x in np.asarray(self):
            s = s + x
        return s

    def to_numpy(self):
        return np.array(self)

    def norm(self):
        a = np.asarray(self)
        s = a.dtype.type(0)
        for x in a:
            s = s + x * x
        return np.sqrt(s)


def _vec(values, dtype=F32):
    return np.array(values, dtype=dtype).view(Vec)


class Mat:
    def __init__(self, rows, dtype):
        self.a = np.array(rows, dtype=dtype)

    def __getitem__(self, idx):
        v = self.a[idx]
        return int(v) if np.issubdtype(self.a.dtype, np.integer) else v

    def __matmul__(self, vec):
        v = np.asarray(vec, dtype=F32)
        out = np.zeros(self.a.shape[0], F32)
        for r in range(self.a.shape[0]):
            s = F32(self.a[r, 0]) * v[0]
            for c in range(1, self.a.shape[1]):
                s = s + F32(self.a[r, c]) * v[c]
            out[r] = s
        return out.view(Vec)


class _ScalarField:
    def __init__(self, dtype, shape):
        self.shape = tuple(shape) if isinstance(shape, (tuple, list)) else (shape,)
        if shape == ():
            self.shape = ()
        self.arr = np.zeros(self.shape, dtype=dtype)

    def __getitem__(self, idx):
        if idx is None:
            return self.arr[()]
        return self.arr[idx]

    def __setitem__(self, idx, val):
        if idx is None:
            self.arr[()] = val
        else:
            self.arr[idx] = val

    def __iter__(self):
        for idx in itertools.product(*(range(n) for n in self.shape)):
            yield tuple(TiInt(i) for i in idx)

    def fill(self, v):
        self.arr[...] = v

    def from_numpy(self, a):
        self.arr[...] = a

    def to_numpy(self):
        return self.arr.copy()


class _VectorField(_ScalarField):
    def __init__(self, n, dtype, shape):
        super().__init__(dtype, shape)
        self.n = n
        self.arr = np.zeros(self.shape + (n,), dtype=dtype)

    def __getitem__(self, idx):
        if idx is None:
            return self.arr.view(Vec)
        return self.arr[idx].view(Vec)

    def __setitem__(self, idx, val):
        if idx is None:
            self.arr[...] = val
        else:
            self.arr[idx] = val


def _field(dtype, shape):
    return _ScalarField(dtype, shape)


class _VectorNS:
    def __call__(self, values):
        return _vec(values)

    @staticmethod
    def field(n, dtype, shape):
        return _VectorField(n, dtype, shape)


def _vector_type(n, dtype):
    dt = F32 if dtype in (float, F32) else dtype

    def ctor(*args):
        if len(args) == 1 and np.isscalar(args[0]):
            return np.full(n, args[0], dtype=dt).view(Vec)
        if len(args) == 1:
            return np.array(args[0], dtype=dt).view(Vec)
        assert len(args) == n
        return np.array(args, dtype=dt).view(Vec)

    return ctor


def _matrix_type(n, m, dtype):
    return lambda rows: Mat(rows, dtype)


def _ndrange(*specs):
    rs = [range(*s) if isinstance(s, tuple) else range(s) for s in specs]
    for idx in itertools.product(*rs):
        yield tuple(TiInt(i) for i in idx)


def _identity_decorator(fn=None, **_kw):
    return fn


def _atomic_max(_a, _b):
    raise NotImplementedError("atomic_max on a Python local cannot be emulated; compute max outside")


def install():
    """Register fake ``taichi`` and ``taichi.math`` modules in sys.modules."""
    ti = types.ModuleType("taichi")
    tm = types.ModuleType("taichi.math")

    ti.init = lambda **kw: None
    ti.gpu = ti.cpu = ti.cuda = "fake"
    ti.INFO = "info"
    ti.f32 = F32
    ti.i32 = np.int32
    ti.data_oriented = lambda cls: cls
    ti.kernel = _identity_decorator
    ti.func = _identity_decorator
    ti.static = lambda x: x
    ti.ndrange = _ndrange
    ti.field = lambda dtype, shape: _field(dtype, shape)
    ti.Vector = _VectorNS()
    ti.atomic_max = _atomic_max
    ti.types = types.SimpleNamespace(vector=_vector_type, matrix=_matrix_type)

    tm.vec2 = lambda a, b: _vec([a, b])
    tm.dot = lambda a, b: F32(a[0]) * F32(b[0]) + F32(a[1]) * F32(b[1])
    tm.sqrt = lambda x: np.sqrt(F32(x))
    tm.min = lambda a, b: F32(min(F32(a), F32(b)))
    tm.max = lambda a, b: F32(max(F32(a), F32(b)))
    # correctly rounded f32 cosine (see oracle/lbm_oracle_np.py header)
    tm.cos = lambda x: F32(_pymath.cos(float(F32(x))))
    ti.math = tm

    sys.modules["taichi"] = ti
    sys.modules["taichi.math"] = tm
    return ti, tm
