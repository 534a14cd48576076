"""CPU restatement of the numeric part of the reference's video frame (test infrastructure only).

`Taichi_Gui_Viz.process_frame` (`src/lbm_mrt_les/visualization/Taichi_Gui_Viz.py:22-51`, cited as viz:LINE) turns
`vel (nx, ny, 2)` into two scalar fields before colouring them on the host:

    vel_x, vel_y = scipy.ndimage.gaussian_filter(vel[..., c], sigma)        viz:24-28
    vel_mag      = sqrt(vel_x**2 + vel_y**2)                                viz:31
    vor          = np.gradient(vel_x)[1] - np.gradient(vel_y)[0]            viz:32-34

scipy's filter is restated here operation for operation (scipy/ndimage/_filters.py `_gaussian_kernel1d`,
`gaussian_filter1d`, and `NI_Correlate1D` in src/ni_filters.c, scipy 1.x): separable, axis 0 then axis 1, every pass
reads float32 and accumulates in double -- centre tap first, then the symmetric pairs from the outermost inwards,
`(in[-j] + in[+j]) * w[j]` -- with mode="reflect" (edge sample repeated) and the result rounded to float32 after each
pass.  `tests/test_viz_oracle.py` checks it bit for bit against scipy itself.
"""
from __future__ import annotations

import numpy as np


def gaussian_weights(sigma: float, truncate: float = 4.0):
    """(radius, weights[0..radius]) -- weights[j] is the tap at distance j (the kernel is symmetric), float64,
    computed exactly as scipy does (the exp is numpy's, so the device takes these numbers, it does not recompute them)."""
    sd = float(sigma)
    radius = int(truncate * sd + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sd * sd) * x ** 2)
    phi = phi / phi.sum()
    return radius, np.ascontiguousarray(phi[radius:], dtype=np.float64)


def reflect_index(i, n):
    """mode="reflect": (d c b a | a b c d | d c b a), any distance past the edge."""
    i = np.mod(i, 2 * n)
    return np.where(i < n, i, 2 * n - 1 - i)


def correlate1d_symmetric(a, radius, w, axis):
    """One pass of NI_Correlate1D (symmetric branch) along `axis`: float32 in, double accumulation, float32 out."""
    a = np.asarray(a, np.float32)
    n = a.shape[axis]
    idx = np.arange(n)
    take = lambda off: np.take(a, reflect_index(idx + off, n), axis=axis).astype(np.float64)  # noqa: E731
    acc = take(0) * w[0]
    for j in range(radius, 0, -1):
        acc = acc + (take(-j) + take(j)) * w[j]
    return acc.astype(np.float32)


def gaussian_filter(a, sigma):
    if sigma <= 0:
        return np.asarray(a, np.float32)
    radius, w = gaussian_weights(sigma)
    out = np.asarray(a, np.float32)
    for axis in range(out.ndim):   # scipy: axes in order, each pass reads the previous pass's float32 output
        out = correlate1d_symmetric(out, radius, w, axis)
    return out


def viz_fields(vel, sigma):
    """(vel_mag, vor), both (nx, ny) float32, as viz:22-34."""
    vel = np.asarray(vel, np.float32)
    vx = gaussian_filter(vel[..., 0], sigma)
    vy = gaussian_filter(vel[..., 1], sigma)
    mag = np.sqrt(vx ** 2 + vy ** 2)
    vor = np.gradient(vx)[1] - np.gradient(vy)[0]
    return mag, vor
