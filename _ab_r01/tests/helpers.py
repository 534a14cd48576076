"""Shared helpers for the test-suite: configs, masks, golden-file access, error norms."""
from __future__ import annotations

import glob
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def make_config(nx, ny, *, bc_type=(0, 2, 1, 2), bc_value=None, rho_in=1.01, rho_out=1.0, nu=0.02, cs=0.1,
                warmup=100, sponge=(4, 8, 2, 2), strength=3.0, s_ghost=1.2, L=8.0, name="case",
                compute_step_size=10, buffer=0, save_h=16):
    """A per-case YAML dict with exactly the keys the reference solver / writer read."""
    if bc_value is None:
        bc_value = [[0.05, 0.0], [0.0, 0.0], [0.0, 0.0], [0.0, 0.0]]
    return {
        "simulation": {
            "name": name, "nx": nx, "ny": ny, "compute_step_size": compute_step_size, "warmup_steps": warmup,
            "nu": nu, "characteristic_length": L, "rho_in": rho_in, "rho_out": rho_out,
            "smagorinsky_constant": cs, "ghost_moments_s": s_ghost, "max_steps": 1000,
        },
        "outputs": {
            "enable_profiling": False,
            "gui": {"enable": False, "gaussian_sigma": 1.0, "interval_steps": compute_step_size, "max_size": 512,
                    "show_zone_overlay": False},
            "video": {"enable": False, "fps": 30, "filename": "x.mp4", "interval_steps": compute_step_size},
            "dataset": {"enable": True, "compression": "lzf", "save_resolution_height": save_h,
                        "interval_steps": compute_step_size},
            "start_record_step": 0,
        },
        "domain_zones": {
            "sponge_in": sponge[0], "sponge_out": sponge[1], "sponge_top": sponge[2], "sponge_bot": sponge[3],
            "buffer": buffer, "sponge_strength": strength,
        },
        "boundary_condition": {"type": list(bc_type), "value": [list(v) for v in bc_value]},
        "mask": {"enable": True, "type": "png", "invert": False, "path": ""},
    }


def cylinder_mask(nx, ny, cx, cy, r):
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    return (i - cx) ** 2 + (j - cy) ** 2 <= r * r


def random_blocks_mask(nx, ny, n, seed, smin=2, smax=6, keep_in=2, keep_out=2):
    rng = np.random.default_rng(seed)
    m = np.zeros((nx, ny), bool)
    for _ in range(n):
        w, h = rng.integers(smin, smax + 1, 2)
        x = rng.integers(keep_in, max(keep_in + 1, nx - keep_out - w))
        y = rng.integers(0, max(1, ny - h))
        m[x:x + w, y:y + h] = True
    return m


def golden_cases():
    return sorted(glob.glob(os.path.join(GOLDEN, "ti_shim_*.npz")))


def load_golden(path):
    z = np.load(path)
    cfg = json.loads(str(z["config_json"]))
    mask = z["mask"] if bool(z["has_mask"]) else None
    return z, cfg, mask


def rel_linf(a, b):
    """max|a-b| / max|b|  -- the north_star parity norm."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))


_EX = (0, 1, 0, -1, 0, 1, -1, -1, 1)
_EY = (0, 0, 1, 0, -1, 1, 1, -1, -1)
_INV = (0, 3, 4, 1, 2, 7, 8, 5, 6)


def force_f64(f_new, mask):
    """Momentum-exchange force (reference LBM2D_MRT_LES.py:588-641) summed exactly-ish in float64, and
    S = sum of |terms|.  The reference adds fp32 terms in an unspecified (atomic) order, and the terms
    cancel to a small net force, so two correct fp32 summations may differ by ~n * eps * S."""
    solid = np.asarray(mask, bool)
    nx, ny = solid.shape
    F = np.zeros(2, np.float64)
    S = 0.0
    for k in range(1, 9):
        ex, ey = _EX[k], _EY[k]
        # solid at (i, j), fluid neighbour at (i + ex, j + ey), both in bounds
        i0, i1 = max(0, -ex), min(nx, nx - ex)
        j0, j1 = max(0, -ey), min(ny, ny - ey)
        s = solid[i0:i1, j0:j1] & ~solid[i0 + ex:i1 + ex, j0 + ey:j1 + ey]
        fv = 2.0 * f_new[i0 + ex:i1 + ex, j0 + ey:j1 + ey, _INV[k]].astype(np.float64)[s]
        F[0] += -ex * fv.sum()
        F[1] += -ey * fv.sum()
        S += np.abs(fv).sum()
    return F, S
