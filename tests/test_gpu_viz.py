"""Device-side video-frame fields (`lbm_get_viz_fields`) against the oracle's restatement of the
reference chain -- scipy gaussian_filter, |u|, np.gradient vorticity (`Taichi_Gui_Viz.py:22-34`) -- which
`tests/test_viz_oracle.py` pins to scipy itself.  Bit-exact."""
import importlib

import numpy as np
import pytest

from helpers import cylinder_mask, make_config
from oracle import viz_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    return importlib.import_module("01-lbm-2d_b200")


@pytest.mark.parametrize("nx,ny", [(150, 70), (9, 5), (3, 40), (130, 4), (64, 33)])
def test_fields_bit_identical_to_scipy_and_numpy(pkg, nx, ny):
    cfg = make_config(nx, ny, rho_in=1.02, nu=0.02, warmup=10, sponge=(2, 3, 1, 1))
    mask = cylinder_mask(nx, ny, nx // 4, ny // 2, max(1, min(nx, ny) // 8)) if min(nx, ny) > 8 else None
    s = pkg.LBM2D_MRT_LES(cfg, mask_data=mask)
    s.init()
    s.run_step(200 if nx > 8 else 40)   # the 3-column channel accelerates without bound (it blows up in the oracle too)
    vel = s.vel.to_numpy()
    assert np.isfinite(vel).all() and np.abs(vel).max() > 1e-4
    for sigma in (1.0, 2.5, 0.6, 0.0):     # radius 4, 10 (longer than the short grids: repeated reflection), 2, off
        mag, vor = s.get_viz_fields(sigma)
        want_mag, want_vor = viz_oracle.viz_fields(vel, sigma)
        assert mag.dtype == np.float32 and mag.shape == (nx, ny)
        assert np.array_equal(mag, want_mag), f"|u| sigma={sigma}"
        assert np.array_equal(vor, want_vor), f"vorticity sigma={sigma}"
    d_mag, _ = s.get_viz_fields()           # default: the config's outputs.gui.gaussian_sigma
    assert np.array_equal(d_mag, viz_oracle.viz_fields(vel, cfg["outputs"]["gui"]["gaussian_sigma"])[0])


def test_device_gui_viz_frame_and_slab_restriction(pkg):
    gv = importlib.import_module("01-lbm-2d_b200.gui_viz")
    capi = importlib.import_module("01-lbm-2d_b200._capi")
    cfg = make_config(48, 24, rho_in=1.01, warmup=5)
    s = pkg.LBM2D_MRT_LES(cfg)
    s.init()
    s.run_step(50)
    frame = gv.DeviceGuiViz(96, 24, viz_sigma=1.0).process_frame_from_solver(s)
    mag, vor = viz_oracle.viz_fields(s.vel.to_numpy(), 1.0)
    assert np.array_equal(frame, np.concatenate((mag, vor), axis=1))
    slab = pkg.LBM2D_MRT_LES(cfg, slab=(0, 24))
    slab.init()
    with pytest.raises(capi.LbmError, match="lbm_comm_connect"):   # a slab needs its neighbours (tests/slab_worker.py)
        slab.get_viz_fields(1.0)
