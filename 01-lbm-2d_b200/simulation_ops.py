"""Taichi-free mirror of the reference's per-case run loop -- the only caller of the hot path.

Same functions, arguments, return values and decision logic as
`src/lbm_mrt_les/core/simulation_ops.py` of the reference (cited as ops:LINE): `check_stability`
(ops:9-57) is the numerical fuse fed by `get_force()` / `get_max_velocity()` after every batch, and
`run_simulation_loop` (ops:60-242) drives `solver.run_step(compute_step_size)` and the viz / video /
dataset sinks at their intervals.  The reference file imports `taichi` only for the name; this one
does not, so the loop runs wherever the CUDA library does.
"""
from __future__ import annotations

import time
import traceback

import numpy as np

try:  # progress bar is cosmetic
    from tqdm import tqdm
except Exception:  # pragma: no cover
    tqdm = None


def check_stability(forces, max_v, step_count, v_threshold=0.25, f_threshold=1e6, warmup_step=1000):
    """ops:9-57.  Returns (is_stable, reason).  NaN/Inf and force blow-up always trip; the velocity
    threshold only after `warmup_step` steps."""
    fx, fy = forces[0], forces[1]
    if np.isnan(fx) or np.isnan(fy) or np.isinf(fx) or np.isinf(fy):
        return False, f"Force becomes NaN/Inf at step {step_count} (Fx={fx}, Fy={fy})"
    if abs(fx) > f_threshold or abs(fy) > f_threshold:
        return False, f"Force exploded (> {f_threshold:.1e}) at step {step_count} (Fx={fx:.2e}, Fy={fy:.2e})"
    if np.isnan(max_v) or np.isinf(max_v):
        return False, f"Velocity field contains NaN/Inf at step {step_count}"
    if step_count > warmup_step and max_v > v_threshold:
        return False, f"Velocity {max_v:.4f} exceeded stability threshold ({v_threshold}) at step {step_count}"
    return True, ""


class _NoBar:
    def set_postfix(self, **_):
        pass

    def update(self, _):
        pass

    def close(self):
        pass


def get_zone_config(config):
    """utils/config_utils.py:22-50 of the reference: sponge widths and the ROI rectangle for the GUI overlay."""
    nx, ny = config["simulation"]["nx"], config["simulation"]["ny"]
    z = config["domain_zones"]
    return {"sponge_in": z["sponge_in"], "sponge_out": z["sponge_out"], "sponge_top": z["sponge_top"], "sponge_bot": z["sponge_bot"],
            "roi_x_start": z["sponge_in"] + z["buffer"], "roi_x_end": nx - z["sponge_out"] - z["buffer"],
            "roi_y_start": z["sponge_bot"] + z["buffer"], "roi_y_end": ny - z["sponge_top"] - z["buffer"], "nx": nx, "ny": ny}


def run_simulation_loop(config, solver, viz, recorder, gui, writer, max_steps, progress=True, draw_zone_overlay=None):
    """ops:60-242.  Returns the metadata dict the batch runner reads (status / reason / final_steps /
    target_steps / re_val / u_max / D / nu).

    `draw_zone_overlay(gui, zones, y_offset=...)`: the reference's `utils.draw_zone_overlay` (visualization/viz_utils.py:52;
    Taichi GUI line drawing, outside this package) -- pass it to get the overlay of ops:155-157; without it the overlay
    is skipped, everything else is the reference's decision logic."""
    sim_cfg = config["simulation"]
    out_cfg = config["outputs"]
    zones = get_zone_config(config)   # ops:67
    step = sim_cfg["compute_step_size"]
    gui_every = out_cfg["gui"]["interval_steps"]
    vid_every = out_cfg["video"]["interval_steps"]
    data_every = out_cfg["dataset"]["interval_steps"]
    start_record = out_cfg.get("start_record_step", 0)
    warmup = sim_cfg["warmup_steps"]
    profiling = out_cfg["enable_profiling"]

    done = 0
    status, reason = "Success", "Reached max_steps"
    bar = tqdm(total=max_steps, unit="step") if (tqdm is not None and progress) else _NoBar()
    timings = {}
    try:
        while done < max_steps:
            t_loop = time.perf_counter()
            if gui and not gui.running:  # ops:92-96
                status, reason = "Aborted", "GUI closed by user"
                break

            t0 = time.perf_counter()  # ops:100-105: advance, then the two per-batch diagnostics
            solver.run_step(step)
            forces = solver.get_force()
            max_v = solver.get_max_velocity()
            done += step
            timings["compute"] = (time.perf_counter() - t0) * 1e3

            ok, why = check_stability(forces, max_v, done, warmup_step=warmup)  # ops:113-124
            if not ok:
                status, reason = "Failed", why
                print(f"\n[CRITICAL] Simulation Failed: {why}")
                break
            bar.set_postfix(Fx=f"{forces[0]:.2e}", Fy=f"{forces[1]:.2e}", MaxV=f"{max_v:.4f}")
            bar.update(step)

            # ops:131-168: visualisation sinks
            t0 = time.perf_counter()
            gui_frame = out_cfg["gui"]["enable"] and done % gui_every == 0
            vid_frame = out_cfg["video"]["enable"] and done % vid_every == 0 and done >= start_record
            img = None
            if gui_frame or vid_frame:   # like ops:145-148, a frame without a `viz` is an error (status "Error")
                if hasattr(viz, "process_frame_from_solver"):   # device-side fields (gui_viz.DeviceGuiViz)
                    img = viz.process_frame_from_solver(solver)
                else:
                    vel, mask = solver.get_physical_fields()
                    img = viz.process_frame(vel, mask)
            if gui_frame and gui:
                gui.set_image(img)
                if out_cfg["gui"]["show_zone_overlay"] and draw_zone_overlay is not None:   # ops:155-157
                    draw_zone_overlay(gui, zones, y_offset=0.0)
                    draw_zone_overlay(gui, zones, y_offset=0.5)
                gui.show()
            if vid_frame and recorder:
                recorder.write_frame(np.transpose(img, (1, 0, 2)))
            timings["viz"] = (time.perf_counter() - t0) * 1e3

            # ops:173-189: dataset sink
            t0 = time.perf_counter()
            data_frame = out_cfg["dataset"]["enable"] and done % data_every == 0 and done >= start_record
            if data_frame and writer:
                if hasattr(writer, "append_from_solver"):  # device-side crop / resize / statistics
                    writer.append_from_solver(solver)
                else:
                    writer.append(solver.get_moments_numpy())
            timings["hdf5_io"] = (time.perf_counter() - t0) * 1e3

            if profiling and (done // step) % 10 == 0:  # ops:194-209
                total = (time.perf_counter() - t_loop) * 1e3
                print(f"\n[Profile] Step {done} | Loop: {total:.1f}ms | compute {timings['compute']:.1f} ms | "
                      f"viz {timings['viz']:.1f} ms | dataset {timings['hdf5_io']:.1f} ms")
    except KeyboardInterrupt:
        status, reason = "Aborted", "User Interrupted (Ctrl+C)"
    except Exception as e:  # ops:216-221
        status, reason = "Error", f"Runtime Error: {e}"
        traceback.print_exc()
    finally:
        bar.close()

    return {  # ops:226-240
        "status": status,
        "reason": reason,
        "final_steps": done,
        "target_steps": max_steps,
        "re_val": float(solver.Re) if hasattr(solver, "Re") else 0.0,
        "u_max": float(np.linalg.norm(solver.u_inlet)) if hasattr(solver, "u_inlet") else 0.0,
        "D": float(config["simulation"]["characteristic_length"]),
        "nu": float(config["simulation"]["nu"]),
    }
