"""GPU parity tests: the CUDA path, called through the C ABI (ctypes -> liblbm2d.so), against the
CPU oracle and the committed golden vectors.

Bars (BASELINE.json north_star): the default (strict) arithmetic is bit-identical to the fp32 oracle (which is itself
bit-identical to the reference source run under the Taichi stand-in) -- rho, u, f, the 9 MRT moments and max|u|.
The optional fast arithmetic is held to relative L-inf (max|a-b| / max|b|) <= 1e-5 on rho and f and, channel by
channel, to max(1e-5, 3 x fp32 noise floor) on the moments and u (helpers.fast_arith_report).  Forces are sums whose
order the reference leaves unspecified (atomics): relative to the sum of |terms|.
"""
import importlib

import numpy as np
import pytest

from helpers import (cylinder_mask, fast_arith_report, force_f64, format_report, golden_cases, load_golden, make_config,
                     random_blocks_mask, rel_linf, rel_linf_channels)
from oracle.lbm_oracle_c import OracleLBMC

pytestmark = pytest.mark.gpu
TOL = 1e-5  # north_star tolerance, relative L-inf


@pytest.fixture(scope="module")
def pkg():
    return importlib.import_module("01-lbm-2d_b200")


def _force_close(force, f_new, mask, rel=2e-6):
    """CUDA force (fp64 tree sum of the fp32 link terms) against the float64 sum over the oracle's f_new.
    Tolerance is relative to S = sum |terms| because the net force is a cancellation of O(S) terms
    (the oracle's own sequential fp32 sum carries ~1e-5 S of rounding noise, see helpers.force_f64)."""
    with np.errstate(all="ignore"):
        F, S = force_f64(f_new, mask if mask is not None else np.zeros(f_new.shape[:2], bool))
    if not np.isfinite(F).all():
        return not np.isfinite(force).all()
    return float(np.max(np.abs(np.asarray(force, np.float64) - F))) <= rel * S + 1e-9


def _assert_bit_exact(s, ref, tag=""):
    assert np.array_equal(s.f_old.to_numpy(), ref.f_old, equal_nan=True), f"f_old {tag}"
    assert np.array_equal(s.f_new.to_numpy(), ref.f_new, equal_nan=True), f"f_new {tag}"
    assert np.array_equal(s.rho.to_numpy(), ref.rho, equal_nan=True), f"rho {tag}"
    assert np.array_equal(s.vel.to_numpy(), ref.vel, equal_nan=True), f"vel {tag}"
    assert np.array_equal(s.get_moments_numpy(), ref.get_moments_numpy(), equal_nan=True), f"moments {tag}"
    mv, rv = s.get_max_velocity(), ref.get_max_velocity()
    assert mv == rv or (np.isnan(mv) and np.isnan(rv)), f"max_v {tag}: {mv} vs {rv}"
    assert _force_close(s.get_force(), ref.f_new, ref.mask), f"force {tag}"


def _assert_close(s, ref, ref64, tol=TOL, tag=""):
    """Fast (FMA / re-associated, non-default) arithmetic against the oracle, field by field and channel by channel.

    rho and f: relative L-inf <= 1e-5 against the fp32 oracle (north_star).  The nine moments and the two velocity
    components: per channel, <= max(1e-5, 3 x the distance of the reference-order fp32 oracle from its own float64
    evaluation) -- see helpers.fast_arith_report for why the momentum-like channels cannot meet a literal 1e-5 in
    ANY fp32 evaluation order.  The bit-exact strict build is the default and has no such caveat.
    """
    errs = {
        "rho": rel_linf(s.rho.to_numpy(), ref.rho),
        "f_old": rel_linf(s.f_old.to_numpy(), ref.f_old),
        "f_new": rel_linf(s.f_new.to_numpy(), ref.f_new),
    }
    assert max(errs.values()) <= tol, (tag, errs)
    ok, rows = fast_arith_report(s.get_moments_numpy(), s.vel.to_numpy(), ref, ref64, tol)
    print(f"{tag} fast per channel: {format_report(rows)}")
    assert ok, (tag, rows)
    assert abs(s.get_max_velocity() - ref.get_max_velocity()) <= 2 * abs(ref.get_max_velocity() - ref64.get_max_velocity()) + 1e-6
    assert _force_close(s.get_force(), ref.f_new, ref.mask, rel=2e-5), tag
    return errs


# ------------------------------------------------------------------ the strict kernel's inline division / square root
def test_packed_division_matches_fdiv_rn():
    """Lane2's FFMA2 Newton sequences (two cells per instruction) against __fdiv_rn / __fsqrt_rn, bit for bit, on
    2 x 10^9 random operand pairs from the guarded range, hard mantissa patterns included."""
    import ctypes as C

    capi = importlib.import_module("01-lbm-2d_b200._capi")
    lib = capi.load()
    for seed in (1, 2):
        bad = (C.c_int64 * 3)()
        capi.check(lib.lbm_selftest_arith(1_000_000_000, seed, bad))
        assert list(bad) == [0, 0, 0], list(bad)


# ------------------------------------------------------------------ golden vectors (reference under shim)
KERNELS = ("register", "tma")


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("path", golden_cases(), ids=lambda p: p.split("ti_shim_")[1][:-4])
def test_strict_build_bit_exact_vs_golden(pkg, path, kernel):
    z, cfg, mask = load_golden(path)
    s = pkg.LBM2D_MRT_LES(cfg, mask_data=mask, arith="strict", kernel=kernel)
    s.init()
    assert np.array_equal(s.f_old.to_numpy(), z["init_f_old"])
    done = 0
    for snap in z["snaps"]:
        snap = int(snap)
        s.run_step(snap - done)
        done = snap
        for nm in ("f_old", "f_new", "rho", "vel"):
            assert np.array_equal(getattr(s, nm).to_numpy(), z[f"s{snap}_{nm}"], equal_nan=True), (nm, snap)
        assert np.array_equal(s.get_moments_numpy(), z[f"s{snap}_moments"], equal_nan=True), snap
        assert s.get_max_velocity() == float(z[f"s{snap}_max_v"]), snap
        assert _force_close(s.get_force(), z[f"s{snap}_f_new"], mask), snap
        assert s.step_count() == snap


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("path", golden_cases(), ids=lambda p: p.split("ti_shim_")[1][:-4])
def test_fast_build_within_tolerance_of_golden(pkg, path, kernel):
    z, cfg, mask = load_golden(path)
    s = pkg.LBM2D_MRT_LES(cfg, mask_data=mask, arith="fast", kernel=kernel)
    s.init()
    last = int(z["snaps"][-1])
    s.run_step(last)
    for nm in ("rho", "f_old", "f_new"):
        assert rel_linf(getattr(s, nm).to_numpy(), z[f"s{last}_{nm}"]) <= TOL, nm
    r32, r64 = OracleLBMC(cfg, mask), OracleLBMC(cfg, mask, dtype=np.float64)  # arbiter for the noise floor
    for o in (r32, r64):
        o.init()
        o.run_step(last)
    assert np.array_equal(r32.f_old, z[f"s{last}_f_old"])
    ok, rows = fast_arith_report(s.get_moments_numpy(), s.vel.to_numpy(), r32, r64, TOL)
    print(format_report(rows))
    assert ok, [r for r in rows if r[1] > r[3]]


# ------------------------------------------------------------------ BASELINE config 1: 512x128 cylinder, 1k steps
def _config1():
    cfg = make_config(512, 128, rho_in=1.003, rho_out=1.0, nu=0.00894, cs=0.1, warmup=1000, sponge=(16, 64, 8, 8),
                      L=20.0, name="cylinder_512x128", compute_step_size=100)
    return cfg, cylinder_mask(512, 128, 128, 64, 10)


@pytest.fixture(scope="module")
def config1_oracle():
    cfg, mask = _config1()
    ref, ref64 = OracleLBMC(cfg, mask), OracleLBMC(cfg, mask, dtype=np.float64)
    for o in (ref, ref64):
        o.init()
        o.run_step(1000)
    return cfg, mask, ref, ref64


@pytest.mark.parametrize("kernel", KERNELS)
def test_config1_cylinder_1k_steps_strict_bit_exact(pkg, config1_oracle, kernel):
    cfg, mask, ref, _ = config1_oracle
    s = pkg.LBM2D_MRT_LES(cfg, mask_data=mask, arith="strict", kernel=kernel)
    s.init()
    for _ in range(10):  # the reference loop: batches of compute_step_size
        s.run_step(100)
    _assert_bit_exact(s, ref, "config1")


@pytest.mark.parametrize("kernel", KERNELS)
def test_config1_cylinder_1k_steps_fast_within_1e5(pkg, config1_oracle, kernel):
    cfg, mask, ref, ref64 = config1_oracle
    s = pkg.LBM2D_MRT_LES(cfg, mask_data=mask, arith="fast", kernel=kernel)
    s.init()
    for _ in range(10):
        s.run_step(100)
    errs = _assert_close(s, ref, ref64, TOL, "config1")
    print("config1 fast rel-Linf:", errs)


# ------------------------------------------------------------------ shapes: unaligned ny, odd nx, tiny grids
@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("nx,ny", [(64, 32), (37, 29), (50, 33), (41, 130), (23, 201), (9, 5), (3, 3), (4, 7), (130, 4),
                                   (17, 129), (25, 257), (9, 128), (10, 300)])
def test_awkward_shapes_strict_bit_exact(pkg, nx, ny, kernel):
    cfg = make_config(nx, ny, rho_in=1.02, nu=0.02, warmup=7, sponge=(min(3, nx // 3), min(5, nx // 3), 2, 2))
    mask = random_blocks_mask(nx, ny, 4, seed=nx * 1000 + ny, smin=1, smax=max(1, min(5, nx // 3, ny // 3)), keep_in=0, keep_out=0)
    ref = OracleLBMC(cfg, mask)
    s = pkg.LBM2D_MRT_LES(cfg, mask_data=mask, arith="strict", kernel=kernel)
    ref.init(), s.init()
    for n in (1, 1, 9, 30):
        ref.run_step(n), s.run_step(n)
        _assert_bit_exact(s, ref, f"{nx}x{ny} after +{n}")


@pytest.mark.parametrize("kernel", KERNELS)
def test_random_bc_types_and_masks_strict_bit_exact(pkg, kernel):
    """Property sweep: random boundary types (incl. no-op 1/3 on odd sides), values, solids on the ring."""
    rng = np.random.default_rng(2024)
    for trial in range(24):
        nx, ny = int(rng.integers(6, 40)), int(rng.integers(5, 40))
        types = [int(t) for t in rng.integers(0, 4, 4)]
        vals = [[float(v) for v in rng.uniform(-0.04, 0.04, 2)] for _ in range(4)]
        cfg = make_config(nx, ny, bc_type=types, bc_value=vals, rho_in=float(rng.uniform(0.98, 1.04)),
                          rho_out=float(rng.uniform(0.98, 1.02)), nu=float(rng.uniform(0.01, 0.1)),
                          cs=float(rng.choice([0.0, 0.1, 0.17])), warmup=int(rng.integers(0, 12)),
                          sponge=tuple(int(v) for v in rng.integers(0, 5, 4)), strength=float(rng.uniform(0, 3)))
        mask = rng.random((nx, ny)) < 0.08
        ref = OracleLBMC(cfg, mask)
        s = pkg.LBM2D_MRT_LES(cfg, mask_data=mask, arith="strict", kernel=kernel)
        ref.init(), s.init()
        ref.run_step(25), s.run_step(25)
        _assert_bit_exact(s, ref, f"trial {trial}: {nx}x{ny} types={types}")
        f = pkg.LBM2D_MRT_LES(cfg, mask_data=mask, arith="fast", kernel=kernel)
        f.init()
        f.run_step(25)
        if np.isfinite(ref.f_old).all():
            assert rel_linf(f.f_old.to_numpy(), ref.f_old) <= TOL, f"trial {trial} fast"


def test_very_long_domain_uses_the_z_dimension_of_the_grid(pkg):
    """nx = 70 000: more than 65 535 grid rows (columns + ring rows), so the row index spills into gridDim.z."""
    nx, ny = 70000, 40
    cfg = make_config(nx, ny, rho_in=1.01, nu=0.02, warmup=5, sponge=(16, 64, 4, 4))
    mask = random_blocks_mask(nx, ny, 400, seed=5, smin=2, smax=12, keep_in=0, keep_out=0)
    ref = OracleLBMC(cfg, mask)
    ref.init()
    ref.run_step(25)
    for kernel in ("register",):
        s = pkg.LBM2D_MRT_LES(cfg, mask_data=mask, arith="strict", kernel=kernel)
        s.init()
        s.run_step(25)
        _assert_bit_exact(s, ref, f"{nx}x{ny} {kernel}")
        s.close()


def test_no_mask_and_all_fluid_equivalent(pkg):
    cfg = make_config(48, 24, rho_in=1.01, warmup=5)
    a = pkg.LBM2D_MRT_LES(cfg, mask_data=None, arith="strict")
    b = pkg.LBM2D_MRT_LES(cfg, mask_data=np.zeros((48, 24), bool), arith="strict")
    a.init(), b.init()
    a.run_step(40), b.run_step(40)
    assert np.array_equal(a.f_old.to_numpy(), b.f_old.to_numpy())
    assert np.array_equal(a.get_force(), np.zeros(2, np.float32))


# ------------------------------------------------------------------ API contract
def test_api_surface_and_batching(pkg):
    cfg, mask = _config1()
    s = pkg.LBM2D_MRT_LES(cfg, mask_data=mask)
    assert (s.nx, s.ny) == (512, 128) and not hasattr(s, "u_inlet")
    assert abs(s.Re - (np.sqrt(2 / 3 * 0.003) * 20 / 0.00894)) < 1e-9 and abs(s.tau_0 - (3 * 0.00894 + 0.5)) < 1e-15
    s.init()
    assert s.get_max_velocity() == 0.0 and s.step_count() == 0
    m0 = s.get_moments_numpy()
    assert m0.shape == (512, 128, 9) and m0.dtype == np.float32
    assert np.allclose(m0[..., 0], 1.0) and np.allclose(m0[..., 1], -2.0, atol=1e-6)
    s.run_step(10)
    s.run_step(5)
    t = pkg.LBM2D_MRT_LES(cfg, mask_data=mask)
    t.init()
    t.run_step(15)
    assert np.array_equal(s.f_old.to_numpy(), t.f_old.to_numpy())
    vel, msk = s.get_physical_fields()
    assert vel.shape == (512, 128, 2) and msk.shape == (512, 128) and msk.dtype == np.float32
    assert np.array_equal(msk, mask.astype(np.float32))
    f = s.get_force()
    assert f.shape == (2,) and f.dtype == np.float32 and f"{f[0]:.2e}"
    a, b = s.get_moments_numpy(), s.get_moments_numpy()
    assert a is not b and a.ctypes.data != b.ctypes.data  # fresh caller-owned arrays
    assert s.launch_count() > 15
    s.init()  # re-init resets the state
    assert s.step_count() == 0 and s.get_max_velocity() == 0.0


def test_nan_propagates_to_the_stability_fuse(pkg):
    cfg = make_config(64, 32, rho_in=8.0, nu=0.0005, cs=0.0, warmup=0, sponge=(0, 0, 0, 0), strength=0.0)
    mask = cylinder_mask(64, 32, 20, 16, 4)
    ref = OracleLBMC(cfg, mask)
    ref.init()
    for arith in ("strict", "fast"):
        s = pkg.LBM2D_MRT_LES(cfg, mask_data=mask, arith=arith)
        s.init()
        s.run_step(400)
        assert np.isnan(s.get_max_velocity()) or s.get_max_velocity() > 0.25
    ref.run_step(400)
    assert np.isnan(ref.get_max_velocity())
    s = pkg.LBM2D_MRT_LES(cfg, mask_data=mask, arith="strict")
    s.init()
    s.run_step(400)
    assert np.isnan(s.get_max_velocity()) and np.isnan(s.get_force()).any()


@pytest.mark.parametrize("les", [True, False])
def test_blow_up_is_bit_identical_to_the_oracle_through_inf_and_nan(pkg, les):
    """A run that diverges: the populations grow through the whole float range, overflow to inf and turn into NaN cell by
    cell.  That walks every rare path of the strict kernel -- the operand guards of the inline packed division / square root
    (library code outside their box), the dense inverse transform for non-finite moments (a zero coefficient times inf is
    NaN, not 0) with its own copy of the tail -- and the result must stay the oracle's, bit for bit (NaN == NaN), at every
    stage, on one GPU kernel per step as on the graph-replayed batches."""
    nx, ny = 96, 70
    cfg = make_config(nx, ny, rho_in=8.0, nu=0.0005, cs=0.17 if les else 0.0, warmup=0, sponge=(6, 12, 4, 4), strength=0.5)
    mask = cylinder_mask(nx, ny, 30, 33, 6)
    ref = OracleLBMC(cfg, mask)
    ref.init()
    s = pkg.LBM2D_MRT_LES(cfg, mask_data=mask, arith="strict")
    s.init()
    seen_inf = seen_nan = seen_mixed = False
    for it in range(60):
        n = 1 if it % 3 == 0 else 9
        ref.run_step(n), s.run_step(n)
        f = s.f_old.to_numpy()
        assert np.array_equal(f, ref.f_old, equal_nan=True), f"f_old after {ref.frame_count if hasattr(ref, 'frame_count') else it} calls"
        assert np.array_equal(s.rho.to_numpy(), ref.rho, equal_nan=True) and np.array_equal(s.vel.to_numpy(), ref.vel, equal_nan=True)
        mv, rv = s.get_max_velocity(), ref.get_max_velocity()
        assert mv == rv or (np.isnan(mv) and np.isnan(rv))
        bad = ~np.isfinite(f)
        seen_inf |= bool(np.isinf(f).any())
        seen_nan |= bool(np.isnan(f).any())
        seen_mixed |= bool(bad.any() and not bad.all())
    assert seen_nan and seen_mixed, (seen_inf, seen_nan, seen_mixed)   # the transition was inside the compared window


# ------------------------------------------------------------------ full-size (BASELINE config 3 grid)
def test_full_size_8192x2048_vs_oracle_and_invariants(pkg):
    nx, ny = 8192, 2048
    cfg = make_config(nx, ny, rho_in=1.01, nu=0.007, cs=0.1, warmup=50, sponge=(128, 896, 128, 128), L=200.0)
    rng = np.random.default_rng(1)
    mask = np.zeros((nx, ny), bool)
    for _ in range(60):
        w, h = rng.integers(60, 400, 2)
        x, y = rng.integers(256, nx - 1024 - w), rng.integers(0, ny - h)
        mask[x:x + w, y:y + h] = True
    ref = OracleLBMC(cfg, mask)
    ref.init()
    ref.run_step(12)
    s = pkg.LBM2D_MRT_LES(cfg, mask_data=mask, arith="strict", kernel="tma")
    s.init()
    s.run_step(12)
    assert np.array_equal(s.rho.to_numpy(), ref.rho) and np.array_equal(s.vel.to_numpy(), ref.vel)
    assert s.get_max_velocity() == ref.get_max_velocity()
    r = pkg.LBM2D_MRT_LES(cfg, mask_data=mask, arith="strict", kernel="register")
    r.init()
    r.run_step(12)
    assert np.array_equal(r.f_old.to_numpy(), ref.f_old)
    del r
    f = pkg.LBM2D_MRT_LES(cfg, mask_data=mask, arith="fast")
    f.init()
    f.run_step(12)
    assert rel_linf(f.rho.to_numpy(), ref.rho) <= TOL and rel_linf(f.get_moments_numpy(), ref.get_moments_numpy()) <= TOL
    assert np.abs(f.vel.to_numpy() - ref.vel).max() <= 1e-6
    assert _force_close(f.get_force(), ref.f_new, ref.mask, rel=2e-5)
    assert _force_close(s.get_force(), ref.f_new, ref.mask)
    # size-independent properties at full size: fast and strict builds stay within tolerance over a
    # longer run, and the rest state (no pressure drop) is a fixed point
    s.run_step(188), f.run_step(188)
    assert rel_linf(f.rho.to_numpy(), s.rho.to_numpy()) <= TOL
    assert rel_linf(f.get_moments_numpy(), s.get_moments_numpy()) <= TOL
    assert np.abs(f.vel.to_numpy() - s.vel.to_numpy()).max() <= 5e-6
    cfg0 = make_config(nx, ny, rho_in=1.0, rho_out=1.0, nu=0.007, sponge=(128, 896, 128, 128))
    r = pkg.LBM2D_MRT_LES(cfg0, mask_data=mask, arith="fast")
    r.init()
    r.run_step(50)
    assert r.get_max_velocity() < 1e-6 and abs(float(r.rho.to_numpy().mean()) - 1.0) < 1e-6


@pytest.mark.parametrize("arith", ["strict", "fast"])
@pytest.mark.parametrize("nx,ny", [(8192, 2048), (2048, 8192)])
def test_early_start_is_bit_identical_to_full_serialisation(pkg, arith, nx, ny, monkeypatch):
    """The first columns of a step start on the progress counter while the previous step drains (step_kernel).
    Same state, bit for bit, as with early start off and as with PDL off, over several hundred steps in uneven
    batches (a race would show up as a difference)."""
    cfg = make_config(nx, ny, rho_in=1.01, nu=0.007, cs=0.1, warmup=50, sponge=(64, 256, 64, 64), L=200.0)
    rng = np.random.default_rng(3)
    mask = np.zeros((nx, ny), bool)
    for _ in range(40):
        w, h = rng.integers(20, 200, 2)
        x, y = rng.integers(0, nx - w), rng.integers(0, ny - h)   # some rectangles touch the first columns / the ring
        mask[x:x + w, y:y + h] = True
    outs = []
    for env in ({}, {"LBM2D_EARLY_CTAS": "0"}, {"LBM2D_NO_PDL": "1"}):
        for k in ("LBM2D_EARLY_CTAS", "LBM2D_NO_PDL"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        s = pkg.LBM2D_MRT_LES(cfg, mask_data=mask, arith=arith, kernel="register")
        s.init()
        for n in (7, 200, 1, 2, 190):
            s.run_step(n)
        outs.append((s.f_old.to_numpy(), s.rho.to_numpy(), s.get_max_velocity(), s.step_count()))
        s.close()
    for o in outs[1:]:
        assert np.array_equal(outs[0][0], o[0]) and np.array_equal(outs[0][1], o[1]) and outs[0][2:] == o[2:]


@pytest.mark.parametrize("nx,ny", [(2048, 512), (1536, 512), (1280, 768)])
def test_early_start_long_run_on_grids_of_a_few_waves(pkg, nx, ny, monkeypatch):
    """The regime where the hand-over is tightest: 2-4 waves of CTAs per step, so the previous step's first columns
    finish only shortly before its tail and the check goes both ways.  30 000 steps, production arithmetic, stable
    flow: bit-identical to plain stream-ordered launches."""
    cfg = make_config(nx, ny, rho_in=1.002, nu=0.05, cs=0.15, warmup=500, sponge=(16, 64, 8, 8))
    rng = np.random.default_rng(nx)
    mask = np.zeros((nx, ny), bool)
    for _ in range(20):
        w, h = rng.integers(8, 60, 2)
        x, y = rng.integers(nx // 8, nx // 2), rng.integers(0, ny - h)
        mask[x:x + w, y:y + h] = True
    out = []
    for no_pdl in (False, True):
        monkeypatch.delenv("LBM2D_NO_PDL", raising=False)
        if no_pdl:
            monkeypatch.setenv("LBM2D_NO_PDL", "1")
        s = pkg.LBM2D_MRT_LES(cfg, mask_data=mask)
        s.init()
        for n in (3, 997, 29000):
            s.run_step(n)
        out.append(s.f_old.to_numpy())
        s.close()
    assert np.isfinite(out[0]).all() and np.array_equal(out[0], out[1])


@pytest.mark.parametrize("arith", ["strict", "fast"])
@pytest.mark.parametrize("nx,ny,bc", [(512, 128, (0, 2, 1, 2)), (1024, 256, (0, 0, 1, 2)), (203, 130, (0, 2, 1, 0))])
def test_graph_replay_of_a_batch_is_bit_identical_to_stream_launches(pkg, arith, nx, ny, bc, monkeypatch):
    """Launch-bound grids replay a whole run_step(K) batch as one CUDA graph once the soft-start ramp is over (lbm_run).
    Same state, bit for bit, as the plain PDL launches: even and odd K (both buffer parities are captured), batches
    before / across / after the end of the ramp, a velocity-Dirichlet wall that reads the ramp, repeated replays."""
    cfg = make_config(nx, ny, rho_in=1.01, nu=0.01, cs=0.1, warmup=40, sponge=(8, 32, 4, 4))
    cfg["boundary_condition"]["type"] = list(bc)
    cfg["boundary_condition"]["value"] = [[0.03, 0.0], [0.02, 0.0], [0.0, 0.0], [0.015, 0.0]]
    mask = cylinder_mask(nx, ny, nx // 4, ny // 2 + 2, max(4, ny // 12))
    outs = []
    for no_graph in (False, True):
        monkeypatch.delenv("LBM2D_NO_GRAPH", raising=False)
        if no_graph:
            monkeypatch.setenv("LBM2D_NO_GRAPH", "1")
        s = pkg.LBM2D_MRT_LES(cfg, mask_data=mask, arith=arith, kernel="register")
        s.init()
        got = []
        # 10 + 25: ramp (stream launches); from step 39 on: graphs; the twelve sizes 8..19 overflow the handle's graph cache
        for n in (10, 25, 25, 100, 100, 33, 33, 100, 9, 1, 100) + tuple(range(8, 20)) + (100, 33):
            s.run_step(n)
            got.append((s.get_max_velocity(), s.step_count(), tuple(s.get_force())))
        outs.append((s.f_old.to_numpy(), s.rho.to_numpy(), s.vel.to_numpy(), got, s.graph_replay_count()))
        s.close()
    a, b = outs
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and a[3] == b[3]
    assert a[3][-1][1] == 536 + 162 + 133 and np.isfinite(a[0]).all()
    assert a[4] == 21 and b[4] == 0   # every batch of >= 8 steps behind the ramp was a graph replay


# ------------------------------------------------------------------ long unsteady run: mean fields (north_star: <= 1e-3)
def test_long_unsteady_run_mean_fields_within_1e3(pkg, tmp_path):
    """30 000 steps of an off-centre cylinder at Re ~ 100 (vortex shedding: the standard deviation of jx over the
    recorded frames is 40 % of its mean), production arithmetic against the strict build (which is bit-identical to
    the fp32 oracle): instantaneous fields drift apart at the 1e-4..1e-3 level, the time-averaged moments -- the
    writer's `mean_vel_field`, accumulated on the device over 201 frames -- must agree to <= 1e-3 relative L-inf."""
    dwm = importlib.import_module("01-lbm-2d_b200.device_writer")
    ops = importlib.import_module("01-lbm-2d_b200.simulation_ops")
    nx, ny = 512, 128
    mask = cylinder_mask(nx, ny, 128, 66, 10)
    res = {}
    for arith in ("strict", "fast"):
        cfg = make_config(nx, ny, rho_in=1.003, nu=0.00894, cs=0.1, warmup=1000, sponge=(16, 64, 8, 8), L=20.0,
                          compute_step_size=100, buffer=0, save_h=56)
        cfg["outputs"]["start_record_step"] = 10000
        s = pkg.LBM2D_MRT_LES(cfg, mask_data=mask, arith=arith)
        s.init()
        w = dwm.DeviceLBMCaseWriter(str(tmp_path / f"{arith}.h5"), cfg, nx, ny, mask_data=mask, solver=s)
        meta = ops.run_simulation_loop(cfg, s, None, None, None, w, max_steps=30000, progress=False)
        assert meta["status"] == "Success" and meta["final_steps"] == 30000
        res[arith] = w.finalize()
    a, b = res["fast"], res["strict"]
    assert a["turbulence"].shape[0] == 201
    unsteady = np.std(b["turbulence"][:, 3], axis=0).max() / np.abs(b["mean_vel_field"][3]).max()
    assert unsteady > 0.1, unsteady   # the case really is unsteady
    for ch in (0, 3, 5):              # rho, jx, jy: what consumers derive u, v, p from
        assert rel_linf(a["mean_vel_field"][ch], b["mean_vel_field"][ch]) <= 1e-3, ch
    assert rel_linf(a["mean_vel_field"], b["mean_vel_field"]) <= 1e-3
    assert rel_linf(a["mean_vel_sq_field"], b["mean_vel_sq_field"]) <= 2e-3


def test_inline_packed_division_equals_the_lane_wise_library_path(pkg, config1_oracle, monkeypatch):
    """LBM2D_NO_FAST_DIV=1 sends every division / square root of the strict kernel through __fdiv_rn / __fsqrt_rn (the
    code the inline FFMA2 sequences fall back to outside their guarded range): same bits, and both equal the oracle."""
    cfg, mask, ref, _ = config1_oracle
    outs = []
    for off in (False, True):
        monkeypatch.delenv("LBM2D_NO_FAST_DIV", raising=False)
        if off:
            monkeypatch.setenv("LBM2D_NO_FAST_DIV", "1")
        s = pkg.LBM2D_MRT_LES(cfg, mask_data=mask)
        s.init()
        s.run_step(1000)
        outs.append(s.f_old.to_numpy())
        s.close()
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], ref.f_old)
