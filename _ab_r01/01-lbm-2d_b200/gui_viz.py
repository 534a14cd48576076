"""`DeviceGuiViz` -- the reference's `Taichi_Gui_Viz` (`src/lbm_mrt_les/visualization/Taichi_Gui_Viz.py:6-51`, cited
as viz:LINE) with the numeric part of `process_frame` moved to the GPU.

The reference pulls `vel (nx, ny, 2)` and `mask (nx, ny)` to the host for every GUI / video frame and runs two scipy
gaussian filters, a norm and two np.gradient calls on them (viz:22-34) -- about a second of host time per frame at
8192x2048 -- before colouring and resizing (viz:36-51).  Here `solver.get_viz_fields()` delivers the filtered |u| and the
vorticity, bit-identical to scipy / numpy, and only the colouring stays on the host: the reference's own
`colorize_velocity` / `colorize_vorticity` / `apply_resize` (matplotlib / cv2 based) are passed in, so this module adds no
dependency and changes no pixel.
"""
from __future__ import annotations

import numpy as np


class DeviceGuiViz:
    def __init__(self, width, height, viz_sigma=1.0, u_norm_max=0.15, vorticity_range=0.03, max_display_size=1024, *,
                 colorize_velocity=None, colorize_vorticity=None, apply_resize=None):
        """Positional arguments as viz:7-20.  The three callables are the reference's
        `visualization.color_utils.colorize_velocity`, `colorize_vorticity` and `utils.apply_resize`; without them
        `process_frame_from_solver` returns the two scalar fields side by side instead of an RGB image."""
        self.width, self.height = width, height
        self.viz_sigma, self.u_norm_max, self.vorticity_range = viz_sigma, u_norm_max, vorticity_range
        self._cv, self._cw, self._resize = colorize_velocity, colorize_vorticity, apply_resize
        self._mask = None

    def process_frame_from_solver(self, solver):
        vel_mag, vor = solver.get_viz_fields(self.viz_sigma)          # viz:24-34, on the device
        if self._cv is None or self._cw is None:
            return np.concatenate((vel_mag, vor), axis=1)
        if self._mask is None:                                        # static: fetched once, not per frame
            self._mask = solver.mask.to_numpy()
        vel_img = self._cv(vel_mag, u_norm_max=self.u_norm_max, mask=self._mask)           # viz:37-41
        vor_img = self._cw(vor, vorticity_range=self.vorticity_range, mask=self._mask)     # viz:43-47
        combined = np.concatenate((vel_img, vor_img), axis=1)                                # viz:50
        return self._resize(combined, self.height, self.width) if self._resize else combined

    def process_frame(self, vel_raw, mask_np):
        raise TypeError("DeviceGuiViz takes its fields from the solver: use process_frame_from_solver(solver); "
                        "for host arrays use the reference's Taichi_Gui_Viz")
