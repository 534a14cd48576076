"""`LBM2D_MRT_LES` -- host-side mirror of the reference solver class, backed by the sm_100a library.

Same constructor, methods, attributes and error behaviour as
`src/lbm_mrt_les/core/LBM2D_MRT_LES.py:10` of the reference (cited as ref:LINE), so
`run_one_case.py:48-49`, `simulation_ops.py:101-103,146,179` and `run_one_case.py:152` run
unchanged against it.  All arithmetic happens on the GPU behind the C ABI of include/lbm2d.h;
this file only parses the config and moves numpy arrays.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math as _math
import threading
import weakref

import numpy as np

from . import _capi


class _FieldShim:
    """Stand-in for the Taichi fields callers touch via `.to_numpy()` (ref:152 of run_one_case.py)."""

    def __init__(self, getter):
        self._getter = getter

    def to_numpy(self):
        return self._getter()


class _PinnedPool:
    """Page-locked frame buffers behind `get_moments_numpy()`.

    The reference contract (ref:739-741): every call returns a fresh array the caller owns -- the
    run loop queues it to the writer thread (io/lbm_writer.py:260-287, queue depth 5) and immediately asks for
    the next one.  A pageable `np.empty` destination costs first-touch page faults and a staged copy (172 ms per
    604 MB frame at 8192x2048); here the array is a view of a pinned buffer (`lbm_host_alloc`), filled by one DMA,
    and a finalizer returns the buffer to the pool when the LAST reference to the array (views included) is dropped.
    Two live arrays never share memory.  `max_buffers` bounds the pinned memory (queue depth 5 + the frame being
    written + the one being filled = 7); beyond it the call falls back to a pageable array."""

    def __init__(self, lib, max_buffers=7, max_bytes=8 << 30):
        self._lib, self._free, self._sizes = lib, {}, {}
        self._max, self._max_bytes, self._lock = max_buffers, max_bytes, threading.Lock()

    def take(self, shape):
        nbytes = int(np.prod(shape)) * 4
        with self._lock:
            ptr = next((p for p in self._free.get(nbytes, [])), None)
            if ptr is not None:
                self._free[nbytes].remove(ptr)
            elif len(self._sizes) < self._max and sum(self._sizes.values()) + nbytes <= self._max_bytes:
                out = C.c_void_p()
                if self._lib.lbm_host_alloc(nbytes, C.byref(out)) != 0:
                    return None
                ptr = out.value
                self._sizes[ptr] = nbytes
            else:
                return None
        buf = (C.c_float * (nbytes // 4)).from_address(ptr)
        weakref.finalize(buf, self._give_back, ptr, nbytes)
        return np.ctypeslib.as_array(buf).reshape(shape)   # the array's base chain keeps `buf` alive

    def free_count(self, shape):
        with self._lock:
            return len(self._free.get(int(np.prod(shape)) * 4, []))

    def _give_back(self, ptr, nbytes):
        with self._lock:
            if ptr in self._sizes:
                self._free.setdefault(nbytes, []).append(ptr)

    def close(self):
        """Frees the buffers that are back in the pool; buffers still owned by live arrays stay valid (and pinned)
        until the process ends."""
        with self._lock:
            for lst in self._free.values():
                for ptr in lst:
                    self._lib.lbm_host_free(C.c_void_p(ptr))
                    self._sizes.pop(ptr, None)
            self._free.clear()
            self._sizes = {}


class LBM2D_MRT_LES:
    _PIN_THRESHOLD = 64 << 20   # bytes: frames at least this large come from the pinned pool
    _EXPORT_PIN_THRESHOLD = 1 << 20   # export frames (9, H, W) at least this large likewise

    def __init__(self, config, mask_data=None, *, arith="strict", kernel="auto", device=None, slab=None,
                 obstacle_mode="refill"):
        """ref:13-29.  `config` is the per-case YAML dict; missing keys raise KeyError like the
        reference.  `mask_data`: bool/float (nx, ny), True/1 = solid, None = all fluid (ref:107-111).

        Extensions (keyword-only, absent from the reference): `arith` = "strict" (default: every operation rounded
        as the reference writes it, bit-identical to the fp32 oracle; 95 % of the HBM roofline) | "fast" (FMA,
        re-associated transforms, SFU reciprocals: tolerance-level parity, ~5 % faster), `kernel` = "auto" | "register" | "tma",
        `device` = CUDA ordinal, `slab` =
        (x0, nx_owned) to own a column range of a larger global domain (multi-GPU), `obstacle_mode` =
        "refill" (the reference's wet-node refill, ref:452-455) | "bounce_back" (half-way bounce-back on the
        solid links, solids frozen at rest -- not reference behaviour; single GPU, default kernel).
        """
        self.config = config
        self._init_params()
        self._lib = _capi.load()

        nx_owned, x0 = self.nx, 0
        if slab is not None:
            x0, nx_owned = int(slab[0]), int(slab[1])
        self._x0, self._nx_owned = x0, nx_owned
        west_halo, east_halo = x0 > 0, x0 + nx_owned < self.nx
        nx_local = nx_owned + int(west_halo) + int(east_halo)
        self._lo = x0 - int(west_halo)  # global x of local column 0

        p = _capi.LbmParams()
        p.nx, p.ny = nx_owned, self.ny
        p.warmup_steps = int(self.warmup_steps)
        p.nu, p.rho_in, p.rho_out = float(self.nu), float(self.rho_in_target), float(self.rho_out_target)
        p.c_smag, p.s_ghost = float(self.C_smag), float(self.S_other)
        zones = self.config["domain_zones"]
        p.sponge_in, p.sponge_out = int(zones["sponge_in"]), int(zones["sponge_out"])
        p.sponge_top, p.sponge_bot = int(zones["sponge_top"]), int(zones["sponge_bot"])
        p.sponge_strength = float(self.sponge_strength)
        bc_cfg = self.config["boundary_condition"]  # ref:114-119
        bc_type = np.array(bc_cfg["type"], dtype=np.int32)
        bc_value = np.array(bc_cfg["value"], dtype=np.float32)
        for d in range(4):
            p.bc_type[d] = int(bc_type[d])
            p.bc_value[d][0] = float(bc_value[d][0])
            p.bc_value[d][1] = float(bc_value[d][1])
        p.arith = _capi.ARITH[arith]
        p.kernel = _capi.KERNEL[kernel]
        p.obstacle_mode = {"refill": 0, "bounce_back": 1}[obstacle_mode]
        p.device = -1 if device is None else int(device)
        p.nx_global, p.slab_x0 = self.nx, x0
        self._params = p

        mask_ptr = None
        if mask_data is not None:  # ref:108-109: mask_data.astype(np.float32); solid where == 1.0
            m = np.asarray(mask_data)
            if m.shape != (self.nx, self.ny):
                raise ValueError(f"mask_data shape {m.shape} != (nx, ny) = {(self.nx, self.ny)}")
            m = (m.astype(np.float32) == 1.0)[self._lo:self._lo + nx_local]
            self._mask_u8 = np.ascontiguousarray(m, dtype=np.uint8)
            mask_ptr = self._mask_u8.ctypes.data_as(C.c_void_p)
        h = C.c_void_p()
        _capi.check(self._lib.lbm_create(C.byref(p), mask_ptr, C.byref(h)))
        self._h = h
        self._pool = None
        self._frame_pool = None

        # Taichi-field look-alikes (ref:99-128); only `.to_numpy()` is supported
        self.vel = _FieldShim(lambda: self._get("lbm_get_vel", (self._nx_owned, self.ny, 2)))
        self.rho = _FieldShim(lambda: self._get("lbm_get_rho", (self._nx_owned, self.ny)))
        self.mask = _FieldShim(lambda: self._get("lbm_get_mask", (self._nx_owned, self.ny), need_init=False))
        self.f_old = _FieldShim(lambda: self._get_f(0))
        self.f_new = _FieldShim(lambda: self._get_f(1))
        self.moments_field = _FieldShim(self.get_moments_numpy)
        self.frame_count = _FieldShim(lambda: np.array(self.step_count(), dtype=np.int32))
        self.force_sum = _FieldShim(self.get_force)

    # ref:32-94 -- same attribute names, same strict indexing, same derived quantities
    def _init_params(self):
        sim_cfg = self.config["simulation"]
        self.name = sim_cfg["name"]
        self.nx = sim_cfg["nx"]
        self.ny = sim_cfg["ny"]
        self.steps_per_frame = sim_cfg["compute_step_size"]
        self.warmup_steps = sim_cfg["warmup_steps"]
        self.nu = sim_cfg["nu"]
        self.tau_0 = 3.0 * self.nu + 0.5
        self.characteristic_length = sim_cfg["characteristic_length"]
        self.rho_in_target = sim_cfg["rho_in"]
        self.rho_out_target = sim_cfg["rho_out"]
        delta_rho = self.rho_in_target - self.rho_out_target
        u_char = _math.sqrt(2.0 / 3.0 * delta_rho) if delta_rho > 1e-9 else 0.01
        self.Re = (u_char * self.characteristic_length) / self.nu if self.nu > 0 else float("inf")
        print(
            f"[Solver] Initialized: target rho_in={self.rho_in_target}, rho_out={self.rho_out_target}, "
            f"u_est={u_char:.5f}, Re_est={self.Re:.1f}"
        )
        self.C_smag = sim_cfg["smagorinsky_constant"]
        self.Cs_sq_factor = 18.0 * (self.C_smag**2)
        self.S_other = sim_cfg["ghost_moments_s"]
        self.viz_sigma = self.config["outputs"]["gui"]["gaussian_sigma"]
        zones = self.config["domain_zones"]
        self.sponge_w_in = max(1, zones["sponge_in"])
        self.sponge_w_out = max(1, zones["sponge_out"])
        self.sponge_w_top = max(1, zones["sponge_top"])
        self.sponge_w_bot = max(1, zones["sponge_bot"])
        self.sponge_strength = zones["sponge_strength"]

    # ------------------------------------------------------------------ reference API
    def init(self):
        """ref:235-241"""
        _capi.check(self._lib.lbm_init(self._h))

    def run_step(self, steps=1):
        """ref:552-573.  Asynchronous; the getters synchronise."""
        _capi.check(self._lib.lbm_run(self._h, int(steps)))

    def get_force(self):
        """ref:644-646 -> np.float32[2]"""
        out = np.zeros(2, np.float32)
        _capi.check(self._lib.lbm_get_force(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def get_max_velocity(self):
        """ref:656-660 -> python float"""
        v = C.c_float()
        _capi.check(self._lib.lbm_get_max_velocity(self._h, C.byref(v)))
        return float(v.value)

    def get_physical_fields(self):
        """ref:207-212 -> (vel (nx,ny,2) f32, mask (nx,ny) f32), fresh host copies"""
        return self.vel.to_numpy(), self.mask.to_numpy()

    def get_moments_numpy(self):
        """ref:739-741 -> fresh (nx,ny,9) f32 array owned by the caller (it is queued to the writer thread).
        Large frames are views of pinned pool buffers filled by one DMA (see _PinnedPool)."""
        shape = (self._nx_owned, self.ny, 9)
        out = None
        if int(np.prod(shape)) * 4 >= self._PIN_THRESHOLD:
            if self._pool is None:
                self._pool = _PinnedPool(self._lib)
            out = self._pool.take(shape)
        if out is None:
            out = np.empty(shape, np.float32)
        _capi.check(self._lib.lbm_get_moments(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    # ------------------------------------------------------------------ on-device export reduction
    def export_configure(self, x0, x1, y0, y1, target_w, target_h):
        """ROI crop + INTER_AREA target of the reference writer (io/lbm_writer.py:37-58); resets the statistics."""
        cfg = _capi.LbmExportConfig(int(x0), int(x1), int(y0), int(y1), int(target_w), int(target_h))
        _capi.check(self._lib.lbm_export_configure(self._h, C.byref(cfg)))
        dlo, dhi, th = C.c_int32(), C.c_int32(), C.c_int32()
        _capi.check(self._lib.lbm_export_layout(self._h, C.byref(dlo), C.byref(dhi), C.byref(th)))
        self.export_columns = (int(dlo.value), int(dhi.value))   # this rank's columns of the global frame
        self._export_shape = (9, int(target_h), int(dhi.value - dlo.value))
        if self._nx_owned == self.nx and self._pinned_frames():   # slabs gather their frames GPU to GPU (slab.py)
            # page-locking is slow (~0.4 ms per MB): the first frames' buffers are made here, at set-up, not inside the run loop
            held = [self._frame_pool.take(self._export_shape) for _ in range(3 - self._frame_pool.free_count(self._export_shape))]
            del held

    def _pinned_frames(self):
        if int(np.prod(self._export_shape)) * 4 < self._EXPORT_PIN_THRESHOLD:
            return False
        if self._frame_pool is None:
            self._frame_pool = _PinnedPool(self._lib, max_bytes=1 << 30)
        return True

    def export_frame(self, want_frame=True):
        """One export frame (9, H, W): moments -> crop -> INTER_AREA on the GPU, statistics accumulated there."""
        out = None
        if want_frame:   # like get_moments_numpy(): a fresh caller-owned array (it is queued to the writer thread), pinned when large
            if self._pinned_frames():
                out = self._frame_pool.take(self._export_shape)
            if out is None:
                out = np.empty(self._export_shape, np.float32)
        _capi.check(self._lib.lbm_export_frame(self._h, out.ctypes.data_as(C.c_void_p) if want_frame else None))
        return out

    def export_frame_device(self):
        """The same frame left on the GPU: (device pointer or None, shape).  Valid until the next export call."""
        ptr = C.c_void_p()
        _capi.check(self._lib.lbm_export_frame_device(self._h, C.byref(ptr)))
        return ptr.value, self._export_shape

    def export_stats(self):
        """Running accumulators of io/lbm_writer.py:176-210 -> dict (float64 arrays) + count."""
        c, h, w = self._export_shape
        rs, vs, vo = np.empty((c, h, w)), np.empty((h, w)), np.empty((h, w))
        mn, mx, cnt = np.empty(9), np.empty(9), C.c_int64()
        p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        _capi.check(self._lib.lbm_export_stats(self._h, p(rs), p(vs), p(vo), p(mn), p(mx), C.byref(cnt)))
        return {"running_sum": rs, "running_vel_sq_sum": vs, "sum_abs_vor": vo, "global_min": mn, "global_max": mx,
                "running_count": int(cnt.value)}

    def static_mask_fields(self, x0, x1, y0, y1, target_w, target_h):
        """(2, H, W) float32 `static_mask` of the case file (io/lbm_writer.py:74-110: nearest-resized ROI mask + signed
        distance field), computed on the device from the resident mask; bit-identical to cv2 + scipy.  The degenerate
        masks (no solid / no fluid pixel in the ROI), where scipy's output is an artefact, go through scipy itself."""
        out = np.empty((2, int(target_h), int(target_w)), np.float32)
        deg = C.c_int32(0)
        _capi.check(self._lib.lbm_static_mask(self._h, int(x0), int(x1), int(y0), int(y1), int(target_w), int(target_h),
                                              out.ctypes.data_as(C.c_void_p), C.byref(deg)))
        if deg.value:
            from .device_writer import static_mask_host

            return static_mask_host(self.mask.to_numpy() == 1.0, x0, x1, y0, y1, target_w, target_h)
        return out

    # ------------------------------------------------------------------ extras
    def get_viz_fields(self, sigma=None):
        """(vel_mag, vorticity), (nx, ny) float32 each: the numeric part of the reference's video frame
        (`visualization/Taichi_Gui_Viz.py:22-34` -- scipy gaussian_filter of both velocity components, |u|,
        np.gradient vorticity) computed on the device, bit-identical to scipy / numpy on `vel.to_numpy()`.
        `sigma` defaults to the config's `outputs.gui.gaussian_sigma`; <= 0 switches the filter off."""
        sigma = float(self.viz_sigma if sigma is None else sigma)
        radius, wptr = 0, None
        if sigma > 0:   # scipy/ndimage/_filters.py: _gaussian_kernel1d with truncate = 4.0
            radius = int(4.0 * sigma + 0.5)
            x = np.arange(-radius, radius + 1)
            phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
            weights = np.ascontiguousarray((phi / phi.sum())[radius:], dtype=np.float64)
            wptr = weights.ctypes.data_as(C.c_void_p)
        mag = np.empty((self._nx_owned, self.ny), np.float32)
        vor = np.empty((self._nx_owned, self.ny), np.float32)
        _capi.check(self._lib.lbm_get_viz_fields(self._h, wptr, radius, mag.ctypes.data_as(C.c_void_p),
                                                 vor.ctypes.data_as(C.c_void_p)))
        return mag, vor

    def step_count(self) -> int:
        v = C.c_int64()
        _capi.check(self._lib.lbm_step_count(self._h, C.byref(v)))
        return int(v.value)

    def synchronize(self):
        _capi.check(self._lib.lbm_synchronize(self._h))

    def launch_count(self) -> int:
        v = C.c_int64()
        _capi.check(self._lib.lbm_launch_count(self._h, C.byref(v)))
        return int(v.value)

    def graph_replay_count(self) -> int:
        """run_step() batches that were replayed as one CUDA graph (launch-bound grids)."""
        v = C.c_int64()
        _capi.check(self._lib.lbm_graph_replay_count(self._h, C.byref(v)))
        return int(v.value)

    def device_view(self):
        v = _capi.LbmDeviceView()
        _capi.check(self._lib.lbm_device_view(self._h, C.byref(v)))
        return v

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.lbm_destroy(h)
        for name in ("_pool", "_frame_pool"):
            pool = getattr(self, name, None)
            setattr(self, name, None)
            if pool is not None:
                pool.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _get(self, fn, shape, need_init=True):
        out = np.empty(shape, np.float32)
        _capi.check(getattr(self._lib, fn)(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def _get_f(self, which):
        out = np.empty((self._nx_owned, self.ny, 9), np.float32)
        _capi.check(self._lib.lbm_get_f(self._h, which, out.ctypes.data_as(C.c_void_p)))
        return out
