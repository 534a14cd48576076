"""GPU: on-device export reduction (ROI crop + INTER_AREA + running statistics) against the writer oracle
fed with the CPU oracle's moments, and against cv2 directly."""
import importlib
import os

import cv2
import numpy as np
import pytest

from helpers import cylinder_mask, make_config, random_blocks_mask, rel_linf
from oracle.lbm_oracle_c import OracleLBMC
from oracle.writer_oracle import WriterOracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    return importlib.import_module("01-lbm-2d_b200")


CASES = [  # (nx, ny, sponge, buffer, save_h): general path, integer fast path (3x3 and 2x2), identity
    (200, 96, (8, 30, 6, 6), 4, 24),
    (142, 70, (6, 16, 3, 3), 2, 20),
    (154, 60, (10, 20, 6, 6), 0, 16),
    (150, 44, (10, 20, 6, 6), 0, 16),
    (90, 40, (5, 9, 4, 4), 1, 30),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}x{c[1]}_h{c[4]}")
def test_export_frames_and_stats_bit_exact_in_strict_mode(pkg, case, tmp_path):
    nx, ny, sponge, buffer, save_h = case
    cfg = make_config(nx, ny, rho_in=1.02, nu=0.02, warmup=20, sponge=sponge, buffer=buffer, save_h=save_h, compute_step_size=15)
    mask = cylinder_mask(nx, ny, nx // 3, ny // 2, max(3, ny // 10)) | random_blocks_mask(nx, ny, 4, seed=nx, keep_in=nx // 4, keep_out=nx // 3)
    ref = OracleLBMC(cfg, mask)
    ref.init()
    s = pkg.LBM2D_MRT_LES(cfg, mask_data=mask, arith="strict")
    s.init()
    dw = importlib.import_module("01-lbm-2d_b200.device_writer")
    w = dw.DeviceLBMCaseWriter(str(tmp_path / "case.h5"), cfg, nx, ny, mask_data=mask, solver=s)
    wo = WriterOracle(cfg, nx, ny)
    assert (w.target_w, w.target_h) == (wo.target_w, wo.target_h)
    for _ in range(4):
        ref.run_step(15), s.run_step(15)
        wo.append(ref.get_moments_numpy())
        w.append_from_solver(s)
        m = ref.get_moments_numpy()[wo.slice_x, wo.slice_y, 3].T  # cv2 itself on one channel
        assert np.array_equal(w.last_frame[3], cv2.resize(np.ascontiguousarray(m), (w.target_w, w.target_h), interpolation=cv2.INTER_AREA))
    got, want = w.finalize(), wo.finalize()
    for k in ("turbulence", "mean_vel_field", "mean_vel_sq_field", "sum_vor"):
        assert np.array_equal(got[k], want[k]), k
    for k in ("stats_min", "stats_max", "stats_mean"):
        assert np.array_equal(np.asarray(w.attrs[k]), want[k]), k
    assert got["static_mask"].shape == (2, w.target_h, w.target_w)
    assert os.path.exists(str(tmp_path / "case.h5")) or os.path.exists(str(tmp_path / "case.npz"))


def test_export_fast_arithmetic_within_tolerance_and_run_loop_uses_it(pkg, tmp_path):
    nx, ny = 256, 96
    cfg = make_config(nx, ny, rho_in=1.01, nu=0.01, warmup=50, sponge=(8, 24, 4, 4), buffer=4, save_h=22, compute_step_size=20)
    cfg["outputs"]["start_record_step"] = 40
    mask = cylinder_mask(nx, ny, 64, 48, 8)
    ref = OracleLBMC(cfg, mask)
    ref.init()
    wo = WriterOracle(cfg, nx, ny)
    for step in range(20, 101, 20):
        ref.run_step(20)
        if step >= 40:
            wo.append(ref.get_moments_numpy())
    s = pkg.LBM2D_MRT_LES(cfg, mask_data=mask, arith="fast")
    s.init()
    dw = importlib.import_module("01-lbm-2d_b200.device_writer")
    ops = importlib.import_module("01-lbm-2d_b200.simulation_ops")
    w = dw.DeviceLBMCaseWriter(str(tmp_path / "c.h5"), cfg, nx, ny, mask_data=mask, solver=s)
    meta = ops.run_simulation_loop(cfg, s, None, None, None, w, max_steps=100, progress=False)
    assert meta["status"] == "Success" and w.n_frames == 4
    got, want = w.finalize(), wo.finalize()
    assert rel_linf(got["turbulence"], want["turbulence"]) <= 1e-5
    assert rel_linf(got["mean_vel_field"], want["mean_vel_field"]) <= 1e-5
    assert np.abs(got["mean_vel_sq_field"] - want["mean_vel_sq_field"]).max() <= 1e-6


@pytest.mark.parametrize("nx,ny,roi,target", [
    (300, 140, (20, 260, 10, 130), (160, 80)),      # ratio 1.5
    (257, 99, (5, 250, 3, 97), (113, 43)),          # awkward fractional ratios
    (512, 128, (16, 448, 8, 120), (246, 64)),       # BASELINE configs[0] geometry
    (200, 96, (0, 200, 0, 96), (200, 96)),          # identity
])
def test_static_mask_and_sdf_on_the_device_match_cv2_and_scipy(pkg, nx, ny, roi, target):
    """(f)-4 / row a19: nearest-resized ROI mask + signed distance field computed from the resident mask, bit-identical to
    the reference's cv2.INTER_NEAREST + scipy.ndimage.distance_transform_edt chain (io/lbm_writer.py:74-110)."""
    dw = importlib.import_module("01-lbm-2d_b200.device_writer")
    mask = cylinder_mask(nx, ny, nx // 3, ny // 2, ny // 7) | random_blocks_mask(nx, ny, 14, seed=ny, smin=2, smax=17, keep_in=0, keep_out=0)
    mask[roi[0] + 3, roi[2]:roi[3]] = True                      # a wall across the ROI, solids on the ROI border
    s = pkg.LBM2D_MRT_LES(make_config(nx, ny), mask_data=mask)
    x0, x1, y0, y1 = roi
    got = s.static_mask_fields(x0, x1, y0, y1, *target)
    want = dw.static_mask_host(mask, x0, x1, y0, y1, *target)
    assert got.shape == want.shape == (2, target[1], target[0]) and got.dtype == np.float32
    assert np.array_equal(got[0], want[0]), "mask"
    assert np.array_equal(got[1], want[1]), float(np.abs(got[1] - want[1]).max())
    # degenerate ROI (no solid inside): scipy's artefact is reproduced through the host path
    empty = pkg.LBM2D_MRT_LES(make_config(nx, ny), mask_data=np.zeros((nx, ny), bool))
    assert np.array_equal(empty.static_mask_fields(x0, x1, y0, y1, *target),
                          dw.static_mask_host(np.zeros((nx, ny), bool), x0, x1, y0, y1, *target))


def test_large_export_frames_are_fresh_pinned_arrays(pkg):
    """Export frames of a megabyte and more come from a page-locked pool (one DMA, no page faults) and are still fresh
    caller-owned arrays: the writer thread keeps up to 5 queued, so two live frames never share memory, a dropped frame's
    buffer is reused, and beyond the pool's 7 buffers the call falls back to pageable arrays."""
    nx, ny = 1200, 400
    cfg = make_config(nx, ny, rho_in=1.01, nu=0.02, warmup=10, sponge=(8, 32, 4, 4), buffer=0, save_h=196, compute_step_size=5)
    s = pkg.LBM2D_MRT_LES(cfg, mask_data=cylinder_mask(nx, ny, 300, 200, 30), arith="strict")
    s.init()
    s.export_configure(8, nx - 32, 4, ny - 4, int((nx - 40) * (196 / (ny - 8))), 196)
    assert int(np.prod(s._export_shape)) * 4 >= s._EXPORT_PIN_THRESHOLD
    s.run_step(5)
    ref = s.export_frame().copy()
    live = [s.export_frame() for _ in range(9)]          # the solver state is unchanged: the same frame nine times
    ptrs = [a.ctypes.data for a in live]
    assert len(set(ptrs)) == 9 and all(np.array_equal(a, ref) for a in live)
    first = ptrs[0]
    del live[0]                                           # its buffer goes back to the pool ...
    again = s.export_frame()
    assert again.ctypes.data == first and np.array_equal(again, ref)   # ... and is what the next frame gets
    s.close()
    assert np.array_equal(again, ref)                     # frames outlive the solver
