#!/usr/bin/env python
"""Generate tests/golden/ti_shim_*.npz by running the UNMODIFIED reference solver source
under the pure-Python Taichi stand-in (fake_taichi.py).

Run in the build container (needs /root/reference; ~10 minutes, or name the cases to (re)generate):
    python tests/golden/gen/make_ti_shim_fixtures.py

The tests never read /root/reference -- they read the committed .npz files.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fake_taichi  # noqa: E402

REF_SOLVER = "/root/reference/src/lbm_mrt_les/core/LBM2D_MRT_LES.py"
OUT_DIR = os.path.dirname(HERE)


def load_reference_class():
    fake_taichi.install()
    spec = importlib.util.spec_from_file_location("ref_LBM2D_MRT_LES", REF_SOLVER)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    # Taichi's `float` is f32 inside kernels (float(frame_count), vector(9, float))
    mod.float = np.float32
    return mod.LBM2D_MRT_LES


def make_config(nx, ny, *, bc_type, bc_value, rho_in, rho_out, nu, cs, warmup, sponge, strength=3.0,
                s_ghost=1.2, name="case"):
    return {
        "simulation": {
            "name": name, "nx": nx, "ny": ny, "compute_step_size": 1, "warmup_steps": warmup,
            "nu": nu, "characteristic_length": 4.0, "rho_in": rho_in, "rho_out": rho_out,
            "smagorinsky_constant": cs, "ghost_moments_s": s_ghost, "max_steps": 100,
        },
        "outputs": {"gui": {"gaussian_sigma": 1.0}},
        "domain_zones": {
            "sponge_in": sponge[0], "sponge_out": sponge[1], "sponge_top": sponge[2],
            "sponge_bot": sponge[3], "buffer": 0, "sponge_strength": strength,
        },
        "boundary_condition": {"type": list(bc_type), "value": [list(v) for v in bc_value]},
    }


def mask_a(nx, ny):
    m = np.zeros((nx, ny), bool)
    m[8:11, 5:8] = True          # interior block
    m[5, 0:2] = True             # touches the bottom ring + its interior neighbour
    m[nx - 1, ny - 1] = True     # top-right corner
    m[0, 6] = True               # west ring cell
    m[nx - 1, 3] = True          # east ring cell
    m[1, 1] = True               # interior cell next to the SW corner
    m[14, ny - 2:ny] = True      # touches the top ring
    m[nx - 2, 8] = True          # neighbour of an east ring cell
    return m


def mask_b(nx, ny):
    m = np.zeros((nx, ny), bool)
    m[4:7, 3:6] = True
    m[0, 0] = True
    m[nx - 1, 0] = True
    m[0, ny - 1] = True
    m[9, 1] = True
    return m


def mask_c(nx, ny):
    """street-canyon style blocks: one flush with the W ring column, one flush with the E ring column, one touching the
    top ring, one single-cell solid, two blocks one fluid cell apart (a one-cell channel)"""
    m = np.zeros((nx, ny), bool)
    m[0:3, 10:15] = True
    m[nx - 4:nx, 20:26] = True
    m[12:18, ny - 5:ny] = True
    m[22, 9] = True
    m[26:31, 6:14] = True
    m[26:31, 15:22] = True
    m[36:40, 1:4] = True
    return m


ZERO4 = [[0.0, 0.0]] * 4
CASES = {
    # template boundary types, solids touching every wall and two corners
    "default": dict(
        cfg=make_config(20, 14, bc_type=[0, 2, 1, 2], bc_value=[[0.05, 0.0]] + ZERO4[1:], rho_in=1.02,
                        rho_out=1.0, nu=0.02, cs=0.1, warmup=8, sponge=(3, 5, 2, 2)),
        mask=mask_a, snaps=(1, 2, 5, 12, 20, 40, 80)),
    # type 0 on top/bottom -> velocity-Dirichlet branch + the `ibc == 0` coordinate quirk in the W corners
    "dirichlet_tb": dict(
        cfg=make_config(16, 12, bc_type=[0, 0, 1, 0], bc_value=[[0.0, 0.0], [0.03, 0.0], [0.0, 0.0], [0.0, 0.0]],
                        rho_in=1.01, rho_out=1.0, nu=0.05, cs=0.17, warmup=5, sponge=(2, 3, 1, 1)),
        mask=mask_b, snaps=(1, 3, 10, 16, 48)),
    # reversed pressure gradient -> outlet backflow guard (ux < 0) and negative inlet ux
    "backflow": dict(
        cfg=make_config(18, 10, bc_type=[0, 2, 1, 2], bc_value=ZERO4, rho_in=1.0, rho_out=1.03, nu=0.03,
                        cs=0.1, warmup=4, sponge=(2, 2, 2, 2)),
        mask=None, snaps=(1, 4, 9, 15, 50)),
    # no-op types on W / bottom, type 1 on top (outlet formula only in the E corner), Dirichlet on E
    "noop_types": dict(
        cfg=make_config(13, 11, bc_type=[1, 1, 0, 3], bc_value=[[0.0, 0.0], [0.0, 0.0], [-0.04, 0.01], [0.0, 0.0]],
                        rho_in=1.0, rho_out=0.99, nu=0.04, cs=0.12, warmup=6, sponge=(0, 0, 0, 0)),
        mask=mask_b, snaps=(1, 2, 7, 14, 45)),
    # LES switched off (Cs <= 0.001), warmup_steps = 0 (ramp == 1 from the first step), no mask
    "les_off_warm0": dict(
        cfg=make_config(15, 9, bc_type=[0, 2, 1, 2], bc_value=ZERO4, rho_in=1.015, rho_out=1.0, nu=0.1,
                        cs=0.0005, warmup=0, sponge=(4, 4, 3, 3), strength=1.5, s_ghost=1.0),
        mask=None, snaps=(1, 2, 6, 12, 40)),
    # larger grid (ny > 32: two pitch lines per column), strong LES, all four sponges, ramp end inside the run
    "blocks_48x34": dict(
        cfg=make_config(48, 34, bc_type=[0, 2, 1, 2], bc_value=ZERO4, rho_in=1.03, rho_out=1.0, nu=0.008,
                        cs=0.17, warmup=30, sponge=(6, 10, 4, 4)),
        mask=mask_c, snaps=(1, 10, 29, 30, 31, 60, 120)),
    # long run through and far past the soft-start ramp (frame_count >> warmup_steps), pressure outlet above 1
    "long_ramp": dict(
        cfg=make_config(24, 16, bc_type=[0, 2, 1, 2], bc_value=ZERO4, rho_in=1.012, rho_out=1.004, nu=0.015,
                        cs=0.1, warmup=100, sponge=(3, 6, 2, 2), strength=2.0),
        mask=mask_a, snaps=(1, 50, 99, 100, 101, 250, 500)),
    # lid-driven cavity: free-slip W/E/bottom, Dirichlet lid on top
    "cavity": dict(
        cfg=make_config(12, 12, bc_type=[2, 0, 2, 2], bc_value=[[0.0, 0.0], [0.06, 0.0], [0.0, 0.0], [0.0, 0.0]],
                        rho_in=1.0, rho_out=1.0, nu=0.01, cs=0.1, warmup=3, sponge=(1, 1, 1, 1)),
        mask=None, snaps=(1, 5, 12, 60)),
}


def run_case(cls, name, spec):
    cfg = spec["cfg"]
    cfg["simulation"]["name"] = name
    nx, ny = cfg["simulation"]["nx"], cfg["simulation"]["ny"]
    mask = spec["mask"](nx, ny) if spec["mask"] is not None else None
    solver = cls(cfg, mask_data=mask)
    # values the reference assigns to kernel locals become f32 in Taichi (rho_in = self.rho_in_target)
    solver.rho_in_target = np.float32(solver.rho_in_target)
    solver.rho_out_target = np.float32(solver.rho_out_target)
    with np.errstate(all="ignore"):
        solver.init()
        out = {
            "config_json": json.dumps(cfg),
            "mask": (mask if mask is not None else np.zeros((nx, ny), bool)),
            "has_mask": np.array(mask is not None),
            "snaps": np.array(spec["snaps"], dtype=np.int32),
            "init_f_old": solver.f_old.to_numpy(),
        }
        done = 0
        for s in spec["snaps"]:
            solver.run_step(s - done)
            done = s
            vel = solver.vel.to_numpy()
            out[f"s{s}_f_old"] = solver.f_old.to_numpy()
            out[f"s{s}_f_new"] = solver.f_new.to_numpy()
            out[f"s{s}_rho"] = solver.rho.to_numpy()
            out[f"s{s}_vel"] = vel
            out[f"s{s}_moments"] = solver.get_moments_numpy()
            out[f"s{s}_force"] = np.asarray(solver.get_force(), dtype=np.float32)
            out[f"s{s}_frame_count"] = np.array(solver.frame_count[None])
            out[f"s{s}_max_v"] = np.array(np.sqrt(vel[..., 0] ** 2 + vel[..., 1] ** 2).max(), np.float32)
    return out


def main():
    import warnings

    warnings.simplefilter("ignore")  # 0-division in warmup_steps=0 etc. is reference behaviour
    cls = load_reference_class()
    only = sys.argv[1:]
    for name, spec in CASES.items():
        if only and name not in only:
            continue
        t0 = time.time()
        out = run_case(cls, name, spec)
        path = os.path.join(OUT_DIR, f"ti_shim_{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{name}: {time.time() - t0:.1f}s -> {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
