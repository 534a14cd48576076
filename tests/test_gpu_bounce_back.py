"""Optional obstacle mode: half-way bounce-back (`obstacle_mode="bounce_back"`).  NOT reference behaviour -- the
reference refills solids with the wet-node equilibrium (ref:452-455, the default and the parity mode).  The checker
is the numpy oracle's restatement of the same rule (oracle/lbm_oracle_np.py) plus the textbook property of the
scheme: the no-slip wall sits half-way between the last fluid node and the first solid node."""
import importlib

import numpy as np
import pytest

from helpers import make_config, rel_linf
from oracle.lbm_oracle_np import OracleLBM

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def pkg():
    return importlib.import_module("01-lbm-2d_b200")


def test_strict_build_bit_exact_vs_the_oracle_rule(pkg):
    rng = np.random.default_rng(77)
    for trial in range(10):
        nx, ny = int(rng.integers(8, 48)), int(rng.integers(6, 70))
        types = [0, 2, 1, 2] if trial % 2 == 0 else [int(t) for t in rng.integers(0, 4, 4)]
        cfg = make_config(nx, ny, bc_type=types, rho_in=float(rng.uniform(1.0, 1.03)), nu=float(rng.uniform(0.01, 0.1)),
                          cs=float(rng.choice([0.0, 0.1])), warmup=int(rng.integers(0, 10)),
                          sponge=tuple(int(v) for v in rng.integers(0, 4, 4)), strength=float(rng.uniform(0, 3)))
        mask = rng.random((nx, ny)) < 0.12          # isolated solids, clusters, solids on the ring
        ref = OracleLBM(cfg, mask, obstacle_mode="bounce_back")
        s = pkg.LBM2D_MRT_LES(cfg, mask_data=mask, arith="strict", obstacle_mode="bounce_back")
        ref.init(), s.init()
        for n in (1, 2, 37):
            ref.run_step(n), s.run_step(n)
            tag = f"trial {trial} {nx}x{ny} +{n}"
            assert np.array_equal(s.f_old.to_numpy(), ref.f_old, equal_nan=True), tag
            assert np.array_equal(s.f_new.to_numpy(), ref.f_new, equal_nan=True), tag
            assert np.array_equal(s.rho.to_numpy(), ref.rho, equal_nan=True), tag
            assert np.array_equal(s.vel.to_numpy(), ref.vel, equal_nan=True), tag
            assert np.array_equal(s.get_moments_numpy(), ref.get_moments_numpy(), equal_nan=True), tag
        f = pkg.LBM2D_MRT_LES(cfg, mask_data=mask, arith="fast", obstacle_mode="bounce_back")
        f.init()
        f.run_step(40)
        if np.isfinite(ref.f_old).all():
            assert rel_linf(f.f_old.to_numpy(), ref.f_old) <= TOL, f"trial {trial} fast"


def test_solids_are_frozen_and_differ_from_the_refill_mode(pkg):
    cfg = make_config(64, 32, rho_in=1.02, nu=0.03, warmup=5)
    mask = np.zeros((64, 32), bool)
    mask[20:26, 12:20] = True
    bb = pkg.LBM2D_MRT_LES(cfg, mask_data=mask, obstacle_mode="bounce_back")
    rf = pkg.LBM2D_MRT_LES(cfg, mask_data=mask)
    bb.init(), rf.init()
    bb.run_step(300), rf.run_step(300)
    assert np.all(bb.rho.to_numpy()[mask] == 1.0) and np.all(bb.vel.to_numpy()[mask] == 0.0)
    assert not np.array_equal(bb.rho.to_numpy(), rf.rho.to_numpy())
    assert np.isfinite(bb.get_force()).all() and bb.get_force()[0] > 0     # drag points downstream


def test_wall_sits_half_way_between_fluid_and_solid_node(pkg):
    """Pressure-driven channel between two solid slabs (no LES, no sponge), steady state: the parabola through the
    fluid nodes vanishes at y = 2.5 and ny - 3.5, i.e. half a cell inside the first solid row (3 solid rows per side);
    the reference's refill rule puts it at 2.2 -- which is why the two modes are not interchangeable."""
    nx, ny = 24, 22
    cfg = make_config(nx, ny, rho_in=1.0006, rho_out=1.0, nu=0.1, cs=0.0, warmup=0, sponge=(0, 0, 0, 0), strength=0.0)
    mask = np.zeros((nx, ny), bool)
    mask[:, :3] = True
    mask[:, -3:] = True
    roots = {}
    for mode in ("bounce_back", "refill"):
        s = pkg.LBM2D_MRT_LES(cfg, mask_data=mask, obstacle_mode=mode)
        s.init()
        s.run_step(6000)
        u = s.vel.to_numpy()[nx // 2, 3:ny - 3, 0].astype(np.float64)
        roots[mode] = np.sort(np.roots(np.polyfit(np.arange(3, ny - 3), u, 2)))
    assert np.abs(roots["bounce_back"] - [2.5, ny - 3.5]).max() < 0.03
    assert np.abs(roots["refill"] - [2.5, ny - 3.5]).max() > 0.2


def test_mode_is_restricted_to_the_default_kernel(pkg):
    capi = importlib.import_module("01-lbm-2d_b200._capi")
    cfg = make_config(32, 16)
    with pytest.raises(capi.LbmError, match="bounce-back"):
        pkg.LBM2D_MRT_LES(cfg, obstacle_mode="bounce_back", kernel="tma")
    pkg.LBM2D_MRT_LES(cfg, obstacle_mode="bounce_back", slab=(0, 16)).close()   # slabs are fine (tests/slab_worker.py)
    with pytest.raises(KeyError):
        pkg.LBM2D_MRT_LES(cfg, obstacle_mode="on_node")
