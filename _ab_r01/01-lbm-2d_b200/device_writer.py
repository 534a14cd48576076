"""`DeviceLBMCaseWriter` -- the reference's HDF5 case writer with its per-frame work moved to the GPU.

Same constructor arguments, datasets, attributes and statistics as `LBMCaseWriter`
(`src/lbm_mrt_les/io/lbm_writer.py:11-251`, cited as writer:LINE): ROI crop, per-channel INTER_AREA
down-sampling to `save_resolution_height`, `turbulence` frames, `static_mask` (mask + SDF),
`mean_vel_field`, `mean_vel_sq_field`, `sum_vor`, `stats_min/max/mean`, `config_json`.  The difference is
where the work happens: `append_from_solver(solver)` asks the solver for the already cropped and resized
(9, H, W) frame (`lbm_export_frame`), so ~11 MB instead of the 604 MB (nx, ny, 9) array cross PCIe per
export at 8192x2048, and the running sums live on the device until `finalize()`.

HDF5 output needs h5py (as in the reference); without it the same datasets go to an .npz next to the
requested path, so the statistics stay usable and testable.
"""
from __future__ import annotations

import json
import os

import numpy as np

try:
    import h5py
except Exception:  # pragma: no cover - h5py is absent from the build image
    h5py = None


class DeviceLBMCaseWriter:
    def __init__(self, file_path, config, nx, ny, channels=9, mask_data=None, solver=None):
        os.makedirs(os.path.dirname(os.path.abspath(file_path)), exist_ok=True)
        self.file_path, self.config, self.nx, self.ny, self.channels = file_path, config, nx, ny, channels
        self.is_closed = False
        zones = config["domain_zones"]  # writer:26-31, strict indexing
        sponge_in, sponge_out = zones["sponge_in"], zones["sponge_out"]
        sponge_top, sponge_bot, buffer = zones["sponge_top"], zones["sponge_bot"], zones["buffer"]
        self.x0, self.x1 = sponge_in, nx - sponge_out - buffer          # writer:37
        self.y0, self.y1 = sponge_bot + buffer, ny - sponge_top - buffer  # writer:38
        self.crop_w, self.crop_h = self.x1 - self.x0, self.y1 - self.y0
        if self.crop_w <= 0 or self.crop_h <= 0:
            raise ValueError(f"[Error] Crop area is invalid! W={self.crop_w}, H={self.crop_h}. Check your domain_zones config.")
        save_res_h = config["outputs"]["dataset"]["save_resolution_height"]
        scale = save_res_h / self.crop_h                                   # writer:55-58
        self.target_w, self.target_h = int(self.crop_w * scale), save_res_h
        self.compression = config["outputs"]["dataset"]["compression"]
        self.frames = []
        self.static_mask = None
        if mask_data is not None:
            self.static_mask = self._static_mask(np.asarray(mask_data))
        self._solver = None
        if solver is not None:
            self.attach(solver)

    def attach(self, solver):
        solver.export_configure(self.x0, self.x1, self.y0, self.y1, self.target_w, self.target_h)
        self._solver = solver

    def _static_mask(self, mask):
        """writer:74-110: nearest-resized mask + signed distance field (fluid positive), host, once per case."""
        import cv2
        import scipy.ndimage

        hw = mask[self.x0:self.x1, self.y0:self.y1].transpose(1, 0).astype(np.float32)
        small = cv2.resize(hw, (self.target_w, self.target_h), interpolation=cv2.INTER_NEAREST)
        small = (small > 0.5).astype(np.float32)
        sdf = scipy.ndimage.distance_transform_edt(1 - small) - scipy.ndimage.distance_transform_edt(small)
        return np.stack([small, sdf], axis=0).astype(np.float32)

    def append_from_solver(self, solver=None):
        if self.is_closed:
            return
        if solver is not None and solver is not self._solver:
            self.attach(solver)
        self.frames.append(self._solver.export_frame())

    def append(self, moment_data):
        raise TypeError("DeviceLBMCaseWriter takes frames from the solver: use append_from_solver(solver); "
                        "for host (nx, ny, 9) arrays use the reference's LBMCaseWriter")

    def finalize(self):
        if self.is_closed:
            return None
        self.is_closed = True
        st = self._solver.export_stats() if self._solver is not None else {"running_count": 0}
        frames = np.stack(self.frames, axis=0) if self.frames else None
        if getattr(self._solver, "world", 1) > 1:
            # x-slabs: every rank holds a column range of the global frame; assemble on rank 0
            sv = self._solver
            frames = sv.gather_columns(frames if frames is not None else np.zeros((0, 9, self.target_h, 0), np.float32))
            parts = {k: sv.gather_columns(st[k]) for k in ("running_sum", "running_vel_sq_sum", "sum_abs_vor")}
            mins = sv.gather_columns(st["global_min"][:, None])
            maxs = sv.gather_columns(st["global_max"][:, None])
            if sv.rank != 0:
                self.result, self.attrs = None, None
                return None
            st = dict(st, **parts, global_min=mins.min(axis=1), global_max=maxs.max(axis=1))
        out = {}
        if self.static_mask is not None:
            out["static_mask"] = self.static_mask
        if st["running_count"] > 0:
            n = st["running_count"]
            mean_field = (st["running_sum"] / n).astype(np.float32)              # writer:224-233
            out.update(
                turbulence=frames if frames is not None else np.zeros((0, 9, self.target_h, self.target_w), np.float32),
                mean_vel_field=mean_field,
                mean_vel_sq_field=(st["running_vel_sq_sum"] / n).astype(np.float32),
                sum_vor=st["sum_abs_vor"].astype(np.float32),
            )
            attrs = {"stats_min": st["global_min"], "stats_max": st["global_max"],
                     "stats_mean": np.mean(mean_field, axis=(1, 2))}
        else:
            attrs = {}
        meta = dict(self.config)
        meta["_dataset_info"] = {"original_crop": [self.crop_w, self.crop_h], "saved_resolution": [self.target_w, self.target_h],
                                 "resize_algo": "INTER_AREA (per channel, on device)"}
        attrs["config_json"] = json.dumps(meta, default=str)
        self.result, self.attrs = out, attrs
        if h5py is not None:
            with h5py.File(self.file_path, "w", libver="latest") as f:
                for k, v in out.items():
                    kw = {"compression": self.compression} if k in ("static_mask", "turbulence") else {}
                    if k == "turbulence":
                        kw["chunks"] = (1, self.channels, self.target_h, self.target_w)
                    f.create_dataset(k, data=v, **kw)
                for k, v in attrs.items():
                    f.attrs[k] = v
        else:
            np.savez(os.path.splitext(self.file_path)[0] + ".npz", **out,
                     **{f"attr_{k}": np.asarray(v) for k, v in attrs.items()})
        return out

    def close(self):
        return self.finalize()
