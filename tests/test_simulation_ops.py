"""Host logic of the run loop (no GPU): the stability fuse and the loop's bookkeeping, with a scripted
stand-in solver.  Mirrors reference behaviour documented in SURVEY.md 3.3 (core/simulation_ops.py)."""
import importlib

import numpy as np
import pytest

from helpers import make_config

ops = importlib.import_module("01-lbm-2d_b200.simulation_ops")


def test_check_stability_branches():
    ok = ops.check_stability
    assert ok([0.1, -0.2], 0.05, 10) == (True, "")
    assert not ok([np.nan, 0.0], 0.05, 10)[0] and "NaN" in ok([np.nan, 0.0], 0.05, 10)[1]
    assert not ok([0.0, np.inf], 0.05, 10)[0]
    assert not ok([2e6, 0.0], 0.05, 10)[0] and "exploded" in ok([2e6, 0.0], 0.05, 10)[1]
    assert not ok([0.0, 0.0], float("nan"), 10)[0]
    assert ok([0.0, 0.0], 0.3, 1000, warmup_step=1000)[0]           # threshold only AFTER the warm-up
    assert not ok([0.0, 0.0], 0.3, 1001, warmup_step=1000)[0]
    assert ok([0.0, 0.0], 0.25, 5000)[0]                             # strictly greater than


class ScriptedSolver:
    Re, nx, ny = 123.0, 8, 4

    def __init__(self, max_v_by_step):
        self.steps, self.calls, self.max_v_by_step = 0, [], max_v_by_step

    def run_step(self, n):
        self.steps += n
        self.calls.append(("run", n))

    def get_force(self):
        return np.array([1.0, -1.0], np.float32)

    def get_max_velocity(self):
        return self.max_v_by_step(self.steps)

    def get_moments_numpy(self):
        self.calls.append(("moments", self.steps))
        return np.zeros((self.nx, self.ny, 9), np.float32)


class ListWriter:
    def __init__(self):
        self.frames = []

    def append(self, m):
        self.frames.append(m)


def _cfg(step=10, warmup=20, start_record=30):
    cfg = make_config(8, 4, compute_step_size=step, warmup=warmup)
    cfg["outputs"]["start_record_step"] = start_record
    return cfg


def test_loop_runs_to_max_steps_and_exports_at_the_interval():
    s, w = ScriptedSolver(lambda n: 0.1), ListWriter()
    meta = ops.run_simulation_loop(_cfg(), s, None, None, None, w, max_steps=60, progress=False)
    assert meta["status"] == "Success" and meta["final_steps"] == 60 and meta["target_steps"] == 60
    assert meta["re_val"] == 123.0 and meta["u_max"] == 0.0 and meta["D"] == 8.0
    assert [c for c in s.calls if c[0] == "run"] == [("run", 10)] * 6
    assert [c[1] for c in s.calls if c[0] == "moments"] == [30, 40, 50, 60]   # start_record_step honoured
    assert len(w.frames) == 4


def test_loop_fails_on_velocity_after_warmup_and_on_nan():
    s = ScriptedSolver(lambda n: 0.3)  # above 0.25: tolerated during warm-up (20 steps), fatal after
    meta = ops.run_simulation_loop(_cfg(), s, None, None, None, None, max_steps=100, progress=False)
    assert meta["status"] == "Failed" and meta["final_steps"] == 30 and "exceeded" in meta["reason"]
    s = ScriptedSolver(lambda n: float("nan") if n >= 20 else 0.0)
    meta = ops.run_simulation_loop(_cfg(), s, None, None, None, None, max_steps=100, progress=False)
    assert meta["status"] == "Failed" and meta["final_steps"] == 20 and "NaN" in meta["reason"]


def test_loop_reports_exceptions_as_error_status():
    class Boom(ScriptedSolver):
        def run_step(self, n):
            raise RuntimeError("device lost")

    meta = ops.run_simulation_loop(_cfg(), Boom(lambda n: 0.0), None, None, None, None, max_steps=10, progress=False)
    assert meta["status"] == "Error" and "device lost" in meta["reason"] and meta["final_steps"] == 0


def test_video_frames_take_the_device_fields_when_the_viz_offers_them():
    """`DeviceGuiViz` (process_frame_from_solver) is asked instead of get_physical_fields + process_frame; a plain viz
    object keeps the reference's call sequence (ops:146-152)."""
    import importlib

    gv = importlib.import_module("01-lbm-2d_b200.gui_viz")

    class S(ScriptedSolver):
        def get_viz_fields(self, sigma):
            self.calls.append(("viz_fields", sigma))
            return np.full((self.nx, self.ny), 0.1, np.float32), np.zeros((self.nx, self.ny), np.float32)

        def get_physical_fields(self):
            self.calls.append(("fields", self.steps))
            return np.zeros((self.nx, self.ny, 2), np.float32), np.zeros((self.nx, self.ny), np.float32)

    class Rec:
        frames = []

        def write_frame(self, f):
            self.frames.append(f.shape)

    class HostViz:
        def process_frame(self, vel, mask):
            return np.zeros((8, 8, 3), np.float32)

    cfg = _cfg()
    cfg["outputs"]["video"]["enable"] = True
    cfg["outputs"]["video"]["interval_steps"] = 20
    s, rec = S(lambda n: 0.1), Rec()
    colour = lambda f, mask=None, **kw: np.repeat(f[..., None], 3, axis=2)  # noqa: E731
    viz = gv.DeviceGuiViz(16, 4, viz_sigma=1.5, colorize_velocity=colour, colorize_vorticity=colour)
    s.mask = type("M", (), {"to_numpy": staticmethod(lambda: np.zeros((8, 4), np.float32))})()
    ops.run_simulation_loop(cfg, s, viz, rec, None, None, max_steps=60, progress=False)
    assert [c for c in s.calls if c[0] == "viz_fields"] == [("viz_fields", 1.5)] * 2      # steps 40, 60 (>= start_record 30)
    assert not [c for c in s.calls if c[0] == "fields"] and rec.frames == [(8, 8, 3)] * 2   # (nx, 2 ny, 3) = (8, 8, 3), transposed
    s2 = S(lambda n: 0.1)
    ops.run_simulation_loop(cfg, s2, HostViz(), Rec(), None, None, max_steps=60, progress=False)
    assert len([c for c in s2.calls if c[0] == "fields"]) == 2 and not [c for c in s2.calls if c[0] == "viz_fields"]
    with pytest.raises(TypeError):
        viz.process_frame(None, None)


def test_gui_frames_draw_the_zone_overlay_twice_like_the_reference():
    """ops:152-158: set_image, the overlay for both halves of the window when `show_zone_overlay` is on, show()."""

    class Gui:
        running, log = True, []

        def set_image(self, img):
            self.log.append("img")

        def show(self):
            self.log.append("show")

    class HostViz:
        def process_frame(self, vel, mask):
            return np.zeros((8, 8, 3), np.float32)

    class S(ScriptedSolver):
        def get_physical_fields(self):
            return np.zeros((self.nx, self.ny, 2), np.float32), np.zeros((self.nx, self.ny), np.float32)

    cfg = _cfg()
    cfg["outputs"]["gui"].update(enable=True, interval_steps=20, show_zone_overlay=True)
    gui, overlay = Gui(), []
    draw = lambda g, zones, y_offset=0.0: (overlay.append((y_offset, zones["roi_x_start"], zones["roi_x_end"])), g.log.append("ovl"))  # noqa: E731
    ops.run_simulation_loop(cfg, S(lambda n: 0.1), HostViz(), None, gui, None, max_steps=40, progress=False, draw_zone_overlay=draw)
    assert gui.log == ["img", "ovl", "ovl", "show"] * 2 and [o[0] for o in overlay] == [0.0, 0.5] * 2
    z = cfg["domain_zones"]
    assert overlay[0][1:] == (z["sponge_in"] + z["buffer"], cfg["simulation"]["nx"] - z["sponge_out"] - z["buffer"])
    # a frame is due but there is no viz object: the reference fails inside the loop and reports status "Error"
    meta = ops.run_simulation_loop(cfg, S(lambda n: 0.1), None, None, Gui(), None, max_steps=40, progress=False)
    assert meta["status"] == "Error"
