"""CPU: the writer oracle against cv2 itself (bit-exact INTER_AREA restatement) and against the
statistics formulas of the reference writer (io/lbm_writer.py:135-251) written out directly."""
import cv2
import numpy as np
import pytest

from helpers import make_config
from oracle.writer_oracle import WriterOracle, area_resize, roi_and_target


@pytest.mark.parametrize("shape", [(48, 220, 8, 36), (60, 100, 20, 33), (64, 64, 32, 32), (64, 96, 16, 32), (30, 41, 7, 9),
                                   (96, 96, 32, 32), (50, 50, 50, 50), (120, 77, 40, 25), (36, 60, 12, 20), (35, 60, 7, 12),
                                   (10, 70, 5, 10), (256, 1173, 43, 196)])
def test_area_resize_is_bit_identical_to_cv2(shape):
    H, W, dh, dw = shape
    img = (np.random.default_rng(H * W).standard_normal((H, W)) * 0.1 + 1).astype(np.float32)
    assert np.array_equal(area_resize(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_AREA))


def test_writer_oracle_matches_the_reference_formulas_with_cv2():
    nx, ny = 96, 64
    cfg = make_config(nx, ny, sponge=(6, 14, 5, 5), buffer=4, save_h=16)
    rng = np.random.default_rng(3)
    w = WriterOracle(cfg, nx, ny)
    sx, sy, tw, th = roi_and_target(cfg, nx, ny)
    assert (sx, sy, tw, th) == (slice(6, 78), slice(9, 55), int(72 * (16 / 46)), 16)
    rs, vs, vo, frames = np.zeros((9, th, tw)), np.zeros((th, tw)), np.zeros((th, tw)), []
    for _ in range(3):
        m = rng.standard_normal((nx, ny, 9)).astype(np.float32) * 0.05
        m[..., 0] += 1.0
        w.append(m)
        hwc = m[sx, sy, :].transpose(1, 0, 2)  # what the reference does, with cv2
        d = np.stack([cv2.resize(hwc[:, :, i], (tw, th), interpolation=cv2.INTER_AREA) for i in range(9)], 2).transpose(2, 0, 1)
        frames.append(d)
        rs += d
        rho_safe = np.maximum(d[0], 1e-6)
        u, v = d[3] / rho_safe, d[5] / rho_safe
        vs += u**2 + v**2
        vo += np.abs(np.gradient(v, axis=1) - np.gradient(u, axis=0))
    out = w.finalize()
    assert np.array_equal(out["turbulence"], np.stack(frames))
    assert np.array_equal(out["mean_vel_field"], (rs / 3).astype(np.float32))
    assert np.array_equal(out["mean_vel_sq_field"], (vs / 3).astype(np.float32))
    assert np.array_equal(out["sum_vor"], vo.astype(np.float32))
    assert np.array_equal(out["stats_min"], np.stack(frames).min(axis=(0, 2, 3)).astype(np.float64))
