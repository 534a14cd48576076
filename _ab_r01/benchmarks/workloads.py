"""Seeded synthetic inputs for the BASELINE.json configs.

The reference's own generators (src/tools/hybrid_map_gen.py, map_gen/shapes.py) are unseeded and
need matplotlib; these are numpy restatements of their *rules* (shape families, size ranges,
minimum distance, blockage limit, fluid margins) with explicit seeds.  They produce the
`(config dict, bool mask (nx, ny))` pair the solver constructor takes.
"""
from __future__ import annotations

import math

import numpy as np


def base_config(nx, ny, *, name, nu, rho_in, rho_out=1.0, cs=0.1, warmup, css, sponge, buffer, L,
                save_h=256, max_steps=10000):
    """Per-case YAML as produced by the reference's config_batch_gen from master_config.yaml:49-112."""
    return {
        "simulation": {
            "name": name, "nx": nx, "ny": ny, "nu": nu, "ghost_moments_s": 1.2, "characteristic_length": L,
            "rho_in": rho_in, "rho_out": rho_out, "smagorinsky_constant": cs, "compute_step_size": css,
            "warmup_steps": warmup, "max_steps": max_steps,
        },
        "outputs": {
            "enable_profiling": False,
            "gui": {"enable": False, "max_size": 1024, "show_zone_overlay": True, "gaussian_sigma": 1.0,
                    "interval_steps": css},
            "video": {"enable": False, "fps": 30, "filename": f"{name}.mp4", "interval_steps": css},
            "dataset": {"enable": True, "compression": "lzf", "save_resolution_height": save_h, "interval_steps": css},
            "project_name": "bench", "data_save_root": "outputs", "target_rho_in": rho_in, "start_record_step": 0,
        },
        "boundary_condition": {"type": [0, 2, 1, 2], "value": [[0.05, 0.0], [0.0, 0.0], [0.0, 0.0], [0.0, 0.0]]},
        "mask": {"enable": True, "type": "png", "invert": False, "path": ""},
        "domain_zones": {"sponge_in": sponge[0], "sponge_out": sponge[1], "sponge_top": sponge[2],
                         "sponge_bot": sponge[3], "buffer": buffer, "sponge_strength": 3.0},
    }


# ---------------------------------------------------------------- shape rasterisers (bounding-box local)
def _draw_circle(mask, cx, cy, r):
    nx, ny = mask.shape
    x0, x1 = max(0, int(cx - r - 1)), min(nx, int(cx + r + 2))
    y0, y1 = max(0, int(cy - r - 1)), min(ny, int(cy + r + 2))
    if x0 >= x1 or y0 >= y1:
        return
    x, y = np.ogrid[x0:x1, y0:y1]
    mask[x0:x1, y0:y1] |= (x - cx) ** 2 + (y - cy) ** 2 <= r * r


def _poly_bbox(pts, shape, pad=0):
    nx, ny = shape
    x0, x1 = max(0, int(math.floor(pts[:, 0].min() - pad))), min(nx, int(math.ceil(pts[:, 0].max() + pad)) + 1)
    y0, y1 = max(0, int(math.floor(pts[:, 1].min() - pad))), min(ny, int(math.ceil(pts[:, 1].max() + pad)) + 1)
    return x0, x1, y0, y1


def _convex_poly_local(pts, shape, pad=0.0):
    """Boolean raster of a convex polygon (vertices in order), optionally grown by `pad` cells."""
    x0, x1, y0, y1 = _poly_bbox(pts, shape, pad)
    if x0 >= x1 or y0 >= y1:
        return (0, 0, 0, 0), None
    x, y = np.ogrid[x0:x1, y0:y1]
    inside = np.ones((x1 - x0, y1 - y0), bool)
    n = len(pts)
    area2 = sum(pts[i, 0] * pts[(i + 1) % n, 1] - pts[(i + 1) % n, 0] * pts[i, 1] for i in range(n))
    sgn = 1.0 if area2 >= 0 else -1.0
    for i in range(n):
        ax, ay = pts[i]
        bx, by = pts[(i + 1) % n]
        ex, ey = bx - ax, by - ay
        ln = math.hypot(ex, ey)
        # signed distance to the edge line, positive inside
        d = sgn * (ex * (y - ay) - ey * (x - ax)) / ln
        inside &= d >= -pad
    return (x0, x1, y0, y1), inside


def rotated_rect_points(cx, cy, w, h, angle_deg):
    a = math.radians(angle_deg)
    c, s = math.cos(a), math.sin(a)
    corners = np.array([[-w / 2, -h / 2], [w / 2, -h / 2], [w / 2, h / 2], [-w / 2, h / 2]])
    rot = np.array([[c, -s], [s, c]])
    return corners @ rot.T + np.array([cx, cy])


def triangle_points(cx, cy, size, angle_deg):
    a = math.radians(angle_deg - 90.0)
    c, s = math.cos(a), math.sin(a)
    base = np.array([[0, -size], [-size * math.sqrt(3) / 2, size / 2], [size * math.sqrt(3) / 2, size / 2]])
    rot = np.array([[c, -s], [s, c]])
    return base @ rot.T + np.array([cx, cy])


def _try_place_poly(mask, keepout, pts, min_dist):
    box, ras = _convex_poly_local(pts, mask.shape)
    if ras is None or not ras.any():
        return False
    x0, x1, y0, y1 = box
    if (keepout[x0:x1, y0:y1] & ras).any():
        return False
    mask[x0:x1, y0:y1] |= ras
    gbox, grown = _convex_poly_local(pts, mask.shape, pad=min_dist)
    keepout[gbox[0]:gbox[1], gbox[2]:gbox[3]] |= grown
    return True


# ---------------------------------------------------------------- BASELINE configs
def cylinder_512x128():
    """configs[0]: single-cylinder channel flow 512x128, Re ~ 100."""
    nx, ny = 512, 128
    mask = np.zeros((nx, ny), bool)
    _draw_circle(mask, 128, 64, 10)
    cfg = base_config(nx, ny, name="cylinder_512x128", nu=0.00894, rho_in=1.003, warmup=1000, css=100,
                      sponge=(16, 64, 8, 8), buffer=0, L=20.0, save_h=64, max_steps=10000)
    return cfg, mask


def tube_bank_2048x512(seed=0):
    """configs[1]: staggered tube bank, circles r=40, 5 columns x 4 rows in x in [0.30, 0.55] W."""
    nx, ny = 2048, 512
    mask = np.zeros((nx, ny), bool)
    cols, rows, r = 5, 4, 40
    xs = np.linspace(0.30 * nx, 0.55 * nx, cols)
    pitch_y = ny / rows
    for ci, cx in enumerate(xs):
        off = 0.5 * pitch_y if ci % 2 else 0.0
        for ri in range(rows + 1):
            cy = (ri + 0.5) * pitch_y - 0.5 * pitch_y + off
            if r + 8 < cy < ny - r - 8:
                _draw_circle(mask, cx, cy, r)
    cfg = base_config(nx, ny, name="tube_bank_2048x512", nu=0.0067, rho_in=1.01, warmup=2000, css=500,
                      sponge=(64, 256, 32, 32), buffer=32, L=80.0, save_h=256, max_steps=50000)
    return cfg, mask


def urban(nx=8192, ny=2048, seed=1, x_lo=256, x_hi_margin=1024, n_rects=100, max_attempts=400):
    """configs[2]: urban-block mask -- rotated rectangles w,h in [60,400], angle +-80 deg, min distance 30,
    row blockage <= 0.6, columns < 256 and > nx-1024 kept fluid (rules of hybrid_map_gen.py:134-176)."""
    rng = np.random.default_rng(seed)
    mask = np.zeros((nx, ny), bool)
    keepout = np.zeros((nx, ny), bool)
    blocked_rows = np.zeros(ny, bool)
    placed = 0
    max_w = 0.0
    for _ in range(max_attempts):
        if placed >= n_rects:
            break
        w, h = rng.uniform(60, 400, 2)
        ang = rng.uniform(-80, 80)
        margin = max(w, h) / 2
        cx = rng.uniform(x_lo + margin, nx - x_hi_margin - margin)
        cy = rng.uniform(margin, ny - margin)
        pts = rotated_rect_points(cx, cy, w, h, ang)
        y0, y1 = max(0, int(pts[:, 1].min())), min(ny, int(pts[:, 1].max()) + 1)
        rows = blocked_rows.copy()
        rows[y0:y1] = True
        if rows.mean() > 0.6:
            continue
        if _try_place_poly(mask, keepout, pts, 30.0):
            blocked_rows = rows
            placed += 1
            max_w = max(max_w, w)
    # blockage-aware rho_in (config_utils/blockage_adjuster.py:16-30, constants.py:26-27)
    xs, xe = max(1, int(nx * 0.05)), min(nx - 1, nx - x_hi_margin - 128)
    per_x = mask[xs:xe].mean(axis=1).astype(np.float32)
    blockage = float(np.convolve(per_x, np.ones(5, np.float32) / 5, mode="valid").max())
    open_fraction = max(0.20, 1.0 - blockage)
    rho_in = min(1.02, 1.0 + 1.5 * (0.15 * open_fraction) ** 2)
    buffer = 128
    cfg = base_config(nx, ny, name=f"urban_{nx}x{ny}", nu=0.007, rho_in=rho_in, warmup=5000, css=500,
                      sponge=(256 - buffer, 1024 - buffer, 256 - buffer, 256 - buffer), buffer=buffer,
                      L=float(max(1.0, max_w)), save_h=256, max_steps=100000)
    return cfg, mask


def random_obstacles(nx=32768, ny=8192, n_shapes=2000, seed=1234, max_attempts=None):
    """configs[3]: 1/3 circles, 1/3 rotated squares, 1/3 triangles, size 16-96, min distance 24."""
    rng = np.random.default_rng(seed)
    mask = np.zeros((nx, ny), bool)
    keepout = np.zeros((nx, ny), bool)
    placed = 0
    x_lo, x_hi = min(512, nx // 8), max(nx - 2048, nx * 3 // 4)
    for _ in range(max_attempts or 3 * n_shapes):
        if placed >= n_shapes:
            break
        kind = placed % 3
        size = rng.uniform(16, 96)
        cx, cy = rng.uniform(x_lo, x_hi), rng.uniform(128, ny - 128)
        if kind == 0:
            n = 24
            ang = np.linspace(0, 2 * math.pi, n, endpoint=False)
            pts = np.stack([cx + size / 2 * np.cos(ang), cy + size / 2 * np.sin(ang)], axis=1)
        elif kind == 1:
            pts = rotated_rect_points(cx, cy, size, size, rng.uniform(0, 90))
        else:
            pts = triangle_points(cx, cy, size / 2, rng.uniform(0, 120))
        if _try_place_poly(mask, keepout, pts, 24.0):
            placed += 1
    cfg = base_config(nx, ny, name=f"random_{nx}x{ny}", nu=0.01, rho_in=1.002, warmup=2000, css=200,
                      sponge=(128, 512, 64, 64), buffer=64, L=96.0, save_h=256, max_steps=100000)
    return cfg, mask


def sweep_case(seed, nx=1024, ny=256):
    """configs[4]: one of the 64 procedural 1024x256 masks (circles / rotated squares / triangles)."""
    cfg, mask = random_obstacles(nx, ny, n_shapes=6 + seed % 7, seed=seed)
    mask[: nx // 8] = False
    mask[nx - nx // 4:] = False
    cfg = base_config(nx, ny, name=f"sweep_{seed:02d}", nu=[0.02, 0.01, 0.0067][seed % 3], rho_in=1.005, warmup=800,
                      css=200, sponge=(32, 128, 16, 16), buffer=16, L=48.0, save_h=64, max_steps=20000)
    return cfg, mask


WORKLOADS = {
    "cylinder": cylinder_512x128,
    "tube_bank": tube_bank_2048x512,
    "urban": urban,
    "random": random_obstacles,
}
