"""CPU tests of DeviceLBMCaseWriter's host side with a scripted solver: both containers (HDF5 through the h5py stand-in,
and the explicitly named raw container), streaming (one frame in memory, frames on disk before finalize), the dataset /
attribute set of the reference writer (io/lbm_writer.py:69-133, 212-251), and the VALUES of `static_mask` (row a19)."""
import importlib
import json
import os

import numpy as np
import pytest

import fake_h5py
from helpers import cylinder_mask, make_config

dw = importlib.import_module("01-lbm-2d_b200.device_writer")


class ScriptedSolver:
    """export_* interface of LBM2D_MRT_LES with deterministic frames."""

    def __init__(self):
        self.n = 0

    def export_configure(self, x0, x1, y0, y1, tw, th):
        self.shape = (9, th, tw)
        self.sum = np.zeros(self.shape)

    def export_frame(self, want_frame=True):
        self.n += 1
        f = (np.arange(np.prod(self.shape), dtype=np.float32).reshape(self.shape) % 7 + self.n).astype(np.float32)
        self.sum += f
        return f

    def export_stats(self):
        c, h, w = self.shape
        return {"running_sum": self.sum, "running_vel_sq_sum": np.full((h, w), 2.0 * self.n), "sum_abs_vor": np.full((h, w), 3.0),
                "global_min": np.arange(9.0), "global_max": np.arange(9.0) + 10, "running_count": self.n}


def _case(tmp_path, container, monkeypatch):
    nx, ny = 120, 60
    cfg = make_config(nx, ny, sponge=(6, 14, 3, 3), buffer=2, save_h=20)
    mask = cylinder_mask(nx, ny, 40, 30, 8)
    if container == "h5":
        monkeypatch.setattr(dw, "h5py", fake_h5py)
    return cfg, mask, dw.DeviceLBMCaseWriter(str(tmp_path / "case.h5"), cfg, nx, ny, mask_data=mask, solver=ScriptedSolver(),
                                             container=container)


@pytest.mark.parametrize("container", ["h5", "h5lite", "raw"])
def test_frames_are_streamed_and_the_reference_dataset_set_is_written(tmp_path, monkeypatch, container):
    cfg, mask, w = _case(tmp_path, container, monkeypatch)
    assert not hasattr(w, "frames")                      # nothing accumulates on the host
    for i in range(5):
        w.append_from_solver()
        w.flush()                                        # the worker thread has appended it
        if container == "h5":                            # the frame is in the file BEFORE finalize (writer:167-169)
            d = fake_h5py.FILES[str(tmp_path / "case.h5")]["turbulence"]
            assert d.shape[0] == i + 1 and d.maxshape[0] is None and d.chunks == (1, 9, w.target_h, w.target_w)
            assert d.compression == "lzf"
        elif container == "h5lite":                      # ... and at the end of the real file, one chunk per frame
            fb = 9 * w.target_h * w.target_w * 4
            assert os.path.getsize(tmp_path / "case.h5") == 512 + 2 * w.target_h * w.target_w * 4 + (i + 1) * fb
            tail = np.fromfile(tmp_path / "case.h5", np.float32, offset=os.path.getsize(tmp_path / "case.h5") - fb)
            assert np.array_equal(tail.reshape(9, w.target_h, w.target_w), w.last_frame)
        else:
            assert os.path.getsize(tmp_path / "case.turbulence.f32") == (i + 1) * 9 * w.target_h * w.target_w * 4
    out = w.finalize()
    assert w.finalize() is None                          # idempotent, like writer:213-214
    assert set(out) == {"static_mask", "turbulence", "mean_vel_field", "mean_vel_sq_field", "sum_vor", "attrs"}
    assert set(out["attrs"]) == {"config_json", "stats_min", "stats_max", "stats_mean"}
    assert out["turbulence"].shape == (5, 9, w.target_h, w.target_w) and out["turbulence"].dtype == np.float32
    assert np.array_equal(np.asarray(out["turbulence"][4]), w.last_frame)
    assert np.array_equal(out["mean_vel_field"], (w._solver.sum / 5).astype(np.float32))
    assert out["mean_vel_sq_field"].dtype == np.float32 and np.all(out["mean_vel_sq_field"] == 2.0)
    assert np.array_equal(out["attrs"]["stats_mean"], np.mean(out["mean_vel_field"], axis=(1, 2)))
    again = dw.read_case(str(tmp_path / "case"))
    assert np.array_equal(np.asarray(again["turbulence"]), np.asarray(out["turbulence"]))


def test_no_frames_means_no_statistics_like_the_reference(tmp_path, monkeypatch):
    _, _, w = _case(tmp_path, "h5", monkeypatch)
    out = w.finalize()
    assert set(out) == {"static_mask", "turbulence", "attrs"} and out["turbulence"].shape[0] == 0 and out["attrs"] == {}


def test_h5_container_needs_h5py_and_auto_says_what_it_does(tmp_path, monkeypatch, capsys):
    monkeypatch.setattr(dw, "h5py", None)
    monkeypatch.setattr(dw, "_warned", False)
    cfg = make_config(120, 60, sponge=(6, 14, 3, 3), buffer=2, save_h=20)
    with pytest.raises(ImportError):
        dw.DeviceLBMCaseWriter(str(tmp_path / "a.h5"), cfg, 120, 60, container="h5")
    out = dw.DeviceLBMCaseWriter(str(tmp_path / "a.h5"), cfg, 120, 60).finalize()
    assert "container='h5lite'" in capsys.readouterr().err
    with open(tmp_path / "a.h5", "rb") as f:                   # without h5py the case file is still HDF5
        assert f.read(8) == b"\x89HDF\r\n\x1a\n"
    assert out["turbulence"].shape[0] == 0 and not os.path.exists(tmp_path / "a.npz")


def _edt_brute(feature):
    """distance of every pixel to the nearest pixel where `feature` is False (scipy's convention), float64."""
    ys, xs = np.nonzero(~feature)
    yy, xx = np.mgrid[0:feature.shape[0], 0:feature.shape[1]]
    d2 = ((yy[..., None] - ys) ** 2 + (xx[..., None] - xs) ** 2).min(axis=-1)
    return np.where(feature, np.sqrt(d2.astype(np.float64)), 0.0)


def test_static_mask_values_row_a19():
    """(2, H, W): channel 0 = nearest-resized ROI mask (transposed to image order), channel 1 = signed distance,
    positive in the fluid, negative in solids (writer:74-110) -- checked against an independent brute-force EDT."""
    nx, ny = 150, 90
    mask = cylinder_mask(nx, ny, 60, 40, 14)
    mask[100:112, 20:50] = True
    x0, x1, y0, y1, tw, th = 10, 130, 6, 84, 60, 39            # ratio 2.0, target (H, W) = (39, 60)
    sm = dw.static_mask_host(mask, x0, x1, y0, y1, tw, th)
    assert sm.shape == (2, th, tw) and sm.dtype == np.float32
    hw = mask[x0:x1, y0:y1].T                                    # image order (H, W)
    want_small = hw[(np.arange(th) * 2)[:, None], (np.arange(tw) * 2)[None, :]]   # INTER_NEAREST at ratio 2: pixel 2 i
    assert np.array_equal(sm[0], want_small.astype(np.float32))
    solid = want_small.astype(bool)
    sdf = _edt_brute(~solid) - _edt_brute(solid)
    assert np.array_equal(sm[1], sdf.astype(np.float32))
    assert (sm[1][~solid] > 0).all() and (sm[1][solid] < 0).all()
    assert sm[1][solid].min() <= -6 and abs(sm[1][0, 0] - np.hypot(*np.argwhere(solid).min(axis=0))) < 12   # deep inside / far away


def test_upsampling_target_takes_the_reference_host_path(tmp_path):
    """`save_resolution_height` above the ROI height: the device reduction covers INTER_AREA shrinking only, so the writer
    computes such frames the way the reference does (cv2 on get_moments_numpy()) instead of refusing the config."""
    import cv2

    nx, ny = 64, 40
    cfg = make_config(nx, ny, sponge=(4, 8, 2, 2), buffer=1, save_h=50)      # ROI 51 x 34 -> target 75 x 50

    class MomentSolver:
        def __init__(self):
            self.n = 0

        def export_configure(self, *a):
            raise AssertionError("the device reduction must not be asked to up-sample")

        def get_moments_numpy(self):
            self.n += 1
            rng = np.random.default_rng(self.n)
            m = rng.standard_normal((nx, ny, 9)).astype(np.float32) * 0.01
            m[..., 0] += 1.0
            return m

    w = dw.DeviceLBMCaseWriter(str(tmp_path / "up.h5"), cfg, nx, ny, solver=MomentSolver(), container="h5lite")
    assert (w.target_w, w.target_h) == (75, 50) and w.target_h > w.crop_h
    for _ in range(3):
        w.append_from_solver()
    out = w.finalize()
    m1 = MomentSolver().get_moments_numpy()
    want = cv2.resize(np.ascontiguousarray(m1[4:55, 3:37, 3].T), (75, 50), interpolation=cv2.INTER_AREA)
    assert out["turbulence"].shape == (3, 9, 50, 75) and np.array_equal(out["turbulence"][0, 3], want)
    assert np.array_equal(out["mean_vel_field"], (np.asarray(out["turbulence"], np.float64).sum(axis=0) / 3).astype(np.float32))
    assert json.loads(out["attrs"]["config_json"])["_dataset_info"]["resize_algo"].endswith("host)")
