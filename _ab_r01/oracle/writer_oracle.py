"""NumPy restatement of the reference's export statistics  --  TEST INFRASTRUCTURE ONLY.

Follows `/root/reference/src/lbm_mrt_les/io/lbm_writer.py` (cited as writer:LINE): ROI crop
(writer:37-42), per-channel `cv2.resize(..., INTER_AREA)` to `save_resolution_height` (writer:52-58,
:150-163) and the running statistics of `append` / `finalize` (writer:176-251).  `area_resize` restates
OpenCV 4.x's INTER_AREA for single-channel float32 shrinking (cv::resize -> resizeArea_ /
ResizeArea_Invoker and the integer-scale resizeAreaFast_ path, modules/imgproc/src/resize.cpp of the
opencv 4.13 the reference pins through cv2) operation for operation; tests check it against cv2 itself
bit for bit, so the CUDA export can be checked against either.
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32


def roi_and_target(config, nx, ny):
    """writer:26-58 -> (slice_x, slice_y, target_w, target_h)"""
    z = config["domain_zones"]
    x0, x1 = z["sponge_in"], nx - z["sponge_out"] - z["buffer"]
    y0, y1 = z["sponge_bot"] + z["buffer"], ny - z["sponge_top"] - z["buffer"]
    crop_w, crop_h = x1 - x0, y1 - y0
    if crop_w <= 0 or crop_h <= 0:
        raise ValueError(f"[Error] Crop area is invalid! W={crop_w}, H={crop_h}. Check your domain_zones config.")
    target_h = config["outputs"]["dataset"]["save_resolution_height"]
    scale = target_h / crop_h
    return slice(x0, x1), slice(y0, y1), int(crop_w * scale), target_h


def area_tab(ssize, dsize):
    """computeResizeAreaTab: list of (si, di, alpha_f32) in table order; scale as cv::resize derives it."""
    inv = float(dsize) / float(ssize)
    scale = 1.0 / inv
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1, sx2 = math.ceil(fsx1), math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((sx1 - 1, dx, F((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((sx, dx, F(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((sx2, dx, F(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return scale, tab


def area_resize(img, dw, dh):
    """cv2.resize(img (H, W) f32, (dw, dh), interpolation=cv2.INTER_AREA) for dw <= W, dh <= H."""
    img = np.ascontiguousarray(img, F)
    H, W = img.shape
    assert dw <= W and dh <= H, "INTER_AREA restated for shrinking only"
    sx_scale, xtab = area_tab(W, dw)
    sy_scale, ytab = area_tab(H, dh)
    ix, iy = int(round(sx_scale)), int(round(sy_scale))
    eps = np.finfo(np.float64).eps
    if abs(sx_scale - ix) < eps and abs(sy_scale - iy) < eps:
        return _area_fast(img, dw, dh, ix, iy)
    out = np.zeros((dh, dw), F)
    xs = np.array([t[0] for t in xtab]); xd = np.array([t[1] for t in xtab]); xa = np.array([t[2] for t in xtab], F)
    total = np.zeros(dw, F)
    prev = ytab[0][1]
    for sy, dy, beta in ytab:
        buf = np.zeros(dw, F)
        prod = img[sy, xs] * xa                     # f32 products, then sequential f32 accumulation per dx
        for k in range(len(xs)):
            buf[xd[k]] = buf[xd[k]] + prod[k]
        if dy != prev:
            out[prev] = total
            total = beta * buf
            prev = dy
        else:
            total = total + beta * buf
    out[prev] = total
    return out


def _area_fast(img, dw, dh, ix, iy):
    """resizeAreaFast_: integer scales.  2x2 uses the SIMD kernel ((a+b) + (c+d)) * 0.25; every other
    area is the scalar loop over the area in row-major order, unrolled by four the way OpenCV does
    (`sum += S[k] + S[k+1] + S[k+2] + S[k+3]`), times 1/area."""
    H, W = img.shape
    out = np.zeros((dh, dw), F)
    w_lim = min(dw, W // ix)  # columns the fast loop covers; remaining ones use the clipped generic tail
    if ix == 2 and iy == 2:
        a, b = img[0:2 * dh:2, 0:2 * dw:2], img[0:2 * dh:2, 1:2 * dw:2]
        c, d = img[1:2 * dh:2, 0:2 * dw:2], img[1:2 * dh:2, 1:2 * dw:2]
        return ((a + b) + (c + d)) * F(0.25)
    scale = F(1.0 / (ix * iy))
    for dy in range(dh):
        for dx in range(w_lim):
            vals = img[dy * iy:(dy + 1) * iy, dx * ix:(dx + 1) * ix].ravel()
            s, k = F(0), 0
            while k <= len(vals) - 4:
                s = s + (((vals[k] + vals[k + 1]) + vals[k + 2]) + vals[k + 3])
                k += 4
            while k < len(vals):
                s = s + vals[k]
                k += 1
            out[dy, dx] = s * scale
    return out


class WriterOracle:
    """Statistics part of LBMCaseWriter (no HDF5): append() frames, finalize() -> dict of datasets / attrs."""

    def __init__(self, config, nx, ny, channels=9):
        self.slice_x, self.slice_y, self.target_w, self.target_h = roi_and_target(config, nx, ny)
        self.channels = channels
        self.frames = []
        self.running_sum = np.zeros((channels, self.target_h, self.target_w), np.float64)
        self.running_vel_sq_sum = np.zeros((self.target_h, self.target_w), np.float64)
        self.sum_abs_vor = np.zeros((self.target_h, self.target_w), np.float64)
        self.running_count = 0
        self.global_min = np.full(channels, np.inf)
        self.global_max = np.full(channels, -np.inf)

    def resize_frame(self, moment_data, resize=area_resize):
        crop = moment_data[self.slice_x, self.slice_y, :]          # writer:144
        hwc = crop.transpose(1, 0, 2)                               # (H, W, C), writer:148
        chans = [resize(np.ascontiguousarray(hwc[:, :, i]), self.target_w, self.target_h) for i in range(self.channels)]
        return np.stack(chans, axis=0).astype(F)                   # (C, H, W), writer:166-170

    def append(self, moment_data, resize=area_resize):
        self.append_frame(self.resize_frame(moment_data, resize))

    def append_frame(self, data_final):
        self.frames.append(data_final)
        self.running_sum += data_final                              # writer:178-179
        self.running_count += 1
        self.global_min = np.minimum(self.global_min, data_final.min(axis=(1, 2)))
        self.global_max = np.maximum(self.global_max, data_final.max(axis=(1, 2)))
        rho, jx, jy = data_final[0], data_final[3], data_final[5]  # writer:189-199
        rho_safe = np.maximum(rho, 1e-6)
        u, v = jx / rho_safe, jy / rho_safe
        self.running_vel_sq_sum += u**2 + v**2
        vor = np.gradient(v, axis=1) - np.gradient(u, axis=0)      # writer:205-210
        self.sum_abs_vor += np.abs(vor)

    def finalize(self):
        if self.running_count == 0:
            return {}
        mean_field = (self.running_sum / self.running_count).astype(F)     # writer:224-233
        return {
            "turbulence": np.stack(self.frames, axis=0),
            "mean_vel_field": mean_field,
            "mean_vel_sq_field": (self.running_vel_sq_sum / self.running_count).astype(F),
            "sum_vor": self.sum_abs_vor.astype(F),
            "stats_min": self.global_min, "stats_max": self.global_max,
            "stats_mean": np.mean(mean_field, axis=(1, 2)),
        }
