"""The restated scipy gaussian filter / np.gradient chain of the reference's video frame
(`visualization/Taichi_Gui_Viz.py:22-34`) against scipy itself, bit for bit (CPU)."""
import numpy as np
import pytest

from oracle import viz_oracle

scipy_ndimage = pytest.importorskip("scipy.ndimage")


@pytest.mark.parametrize("shape", [(64, 48), (7, 5), (3, 40), (33, 2), (129, 65)])
@pytest.mark.parametrize("sigma", [1.0, 0.6, 2.5, 4.0])
def test_gaussian_filter_is_scipy_bit_for_bit(shape, sigma):
    rng = np.random.default_rng(hash((shape, sigma)) % 2**32)
    a = (rng.standard_normal(shape) * 0.05).astype(np.float32)
    want = scipy_ndimage.gaussian_filter(a, sigma=sigma)
    got = viz_oracle.gaussian_filter(a, sigma)
    assert got.dtype == np.float32 and np.array_equal(got, want)


def test_fields_match_the_reference_formulas_evaluated_with_scipy():
    rng = np.random.default_rng(0)
    vel = (rng.standard_normal((40, 28, 2)) * 0.03).astype(np.float32)
    for sigma in (1.0, 0.0):
        if sigma > 0:
            vx = scipy_ndimage.gaussian_filter(vel[:, :, 0], sigma=sigma)
            vy = scipy_ndimage.gaussian_filter(vel[:, :, 1], sigma=sigma)
        else:
            vx, vy = vel[:, :, 0], vel[:, :, 1]
        mag = np.sqrt(vx ** 2 + vy ** 2)
        ug, vg = np.gradient(vx), np.gradient(vy)
        vor = ug[1] - vg[0]
        m, w = viz_oracle.viz_fields(vel, sigma)
        assert np.array_equal(m, mag) and np.array_equal(w, vor) and m.dtype == np.float32 and w.dtype == np.float32
