"""CPU tests of the oracle itself: pins against the reference-under-shim golden vectors,
C == numpy restatement, and the analytic invariants of SURVEY.md section 8(c)."""
import numpy as np
import pytest

from helpers import cylinder_mask, golden_cases, load_golden, make_config, random_blocks_mask, rel_linf
from oracle.lbm_oracle_c import OracleLBMC
from oracle.lbm_oracle_np import E, M_NP, OracleLBM, inv_m, ramp_value

FIELDS = ("f_old", "f_new", "rho", "vel")


def _check_against_golden(o, z):
    assert np.array_equal(o.f_old, z["init_f_old"])
    done = 0
    for s in z["snaps"]:
        s = int(s)
        o.run_step(s - done)
        done = s
        for nm in FIELDS:
            assert np.array_equal(getattr(o, nm), z[f"s{s}_{nm}"], equal_nan=True), (nm, s)
        assert np.array_equal(o.get_moments_numpy(), z[f"s{s}_moments"], equal_nan=True), s
        assert np.array_equal(o.get_force(), z[f"s{s}_force"]), s
        assert o.get_max_velocity() == float(z[f"s{s}_max_v"]), s
        assert o.frame_count == int(z[f"s{s}_frame_count"]) == s


@pytest.mark.parametrize("path", golden_cases(), ids=lambda p: p.split("ti_shim_")[1][:-4])
def test_numpy_oracle_bit_exact_vs_reference_under_shim(path):
    """Golden vectors = the unmodified reference source run under tests/golden/gen/fake_taichi.py."""
    z, cfg, mask = load_golden(path)
    o = OracleLBM(cfg, mask_data=mask)
    o.init()
    _check_against_golden(o, z)


@pytest.mark.parametrize("path", golden_cases(), ids=lambda p: p.split("ti_shim_")[1][:-4])
def test_c_oracle_bit_exact_vs_reference_under_shim(path):
    z, cfg, mask = load_golden(path)
    o = OracleLBMC(cfg, mask_data=mask)
    o.init()
    _check_against_golden(o, z)


def test_golden_set_is_complete():
    names = {p.split("ti_shim_")[1][:-4] for p in golden_cases()}
    assert {"default", "dirichlet_tb", "backflow", "noop_types", "les_off_warm0", "cavity"} <= names


def test_golden_backflow_branch_is_exercised():
    z, cfg, _ = load_golden([p for p in golden_cases() if "backflow" in p][0])
    v = z["s15_vel"]
    assert (v[-1, 1:-1, 0] < 0).all() and np.array_equal(v[-1, 1:-1], v[-2, 1:-1])


def test_c_oracle_equals_numpy_oracle_larger_case():
    cfg = make_config(96, 40, rho_in=1.02, nu=0.01, warmup=30, sponge=(6, 12, 4, 4))
    mask = cylinder_mask(96, 40, 24, 20, 5) | random_blocks_mask(96, 40, 6, seed=3, keep_in=30, keep_out=14)
    a, b = OracleLBM(cfg, mask), OracleLBMC(cfg, mask)
    a.init(), b.init()
    for n in (1, 7, 60):
        a.run_step(n), b.run_step(n)
        for nm in FIELDS:
            assert np.array_equal(getattr(a, nm), getattr(b, nm)), nm
    assert np.array_equal(a.get_moments_numpy(), b.get_moments_numpy())
    assert np.array_equal(a.get_force(), b.get_force())
    assert a.get_max_velocity() == b.get_max_velocity()


def test_f64_oracle_c_vs_numpy_and_close_to_f32():
    cfg = make_config(48, 24, rho_in=1.01, nu=0.02, warmup=20)
    mask = cylinder_mask(48, 24, 14, 12, 3)
    a, b = OracleLBM(cfg, mask, dtype=np.float64), OracleLBMC(cfg, mask, dtype=np.float64)
    c = OracleLBMC(cfg, mask)
    for o in (a, b, c):
        o.init()
        o.run_step(200)
    assert np.allclose(a.f_old, b.f_old, rtol=0, atol=1e-14)
    assert rel_linf(c.rho, b.rho) < 1e-5 and rel_linf(c.vel, b.vel) < 1e-4


# ---------------------------------------------------------------- invariants, SURVEY 8(c) (i)-(v)
def test_M_times_invM_is_identity_and_exact_fractions():
    lit, ex = inv_m(False), inv_m(True)
    assert np.allclose(M_NP.astype(np.float64) @ ex.astype(np.float64), np.eye(9), atol=1e-6)
    nz = ex != 0
    assert np.array_equal(lit[nz], ex[nz])          # non-zeros are the correctly rounded fractions
    assert np.abs(lit[~nz]).max() < 1e-15           # the rest is inversion noise


def test_exact_inv_m_gives_identical_trajectories():
    cfg = make_config(40, 20, rho_in=1.02, warmup=10)
    mask = cylinder_mask(40, 20, 12, 10, 3)
    a, b = OracleLBMC(cfg, mask), OracleLBMC(cfg, mask, exact_inv_m=True)
    a.init(), b.init()
    a.run_step(150), b.run_step(150)
    assert np.array_equal(a.f_old, b.f_old) and np.array_equal(a.f_new, b.f_new)


def test_M_feq_equals_meq():
    o = OracleLBM(make_config(8, 8), dtype=np.float64)
    rng = np.random.default_rng(0)
    rho = 1 + 0.05 * rng.standard_normal(50)
    vel = 0.05 * rng.standard_normal((50, 2))
    feq = o._f_eq(rho, vel)
    m = feq @ M_NP.astype(np.float64).T
    u, v = vel[:, 0], vel[:, 1]
    u2 = u * u + v * v
    meq = np.stack([rho, rho * (-2 + 3 * u2), rho * (1 - 3 * u2), rho * u, -rho * u, rho * v, -rho * v,
                    rho * (u * u - v * v), rho * u * v], axis=1)
    assert np.allclose(m, meq, atol=1e-12)


def test_rest_state_is_a_fixed_point():
    cfg = make_config(32, 16, rho_in=1.0, rho_out=1.0)
    o = OracleLBMC(cfg, cylinder_mask(32, 16, 10, 8, 2))
    o.init()
    f0 = o.f_old.copy()
    o.run_step(50)
    assert np.abs(o.f_old - f0).max() < 2e-7
    assert o.get_max_velocity() < 1e-6


def test_collision_conserves_mass_and_momentum():
    cfg = make_config(32, 16, rho_in=1.03, warmup=5)
    o = OracleLBM(cfg, dtype=np.float64)
    o.init()
    o.run_step(30)
    nx, ny = o.nx, o.ny
    from oracle.lbm_oracle_np import E
    pulled = np.stack([o.f_old[1 - E[k, 0]:nx - 1 - E[k, 0], 1 - E[k, 1]:ny - 1 - E[k, 1], k] for k in range(9)], -1)
    o.collide_and_stream()
    post = o.f_new[1:-1, 1:-1]
    assert np.allclose(pulled.sum(-1), post.sum(-1), atol=1e-13)
    assert np.allclose(pulled @ E[:, 0].astype(float), post @ E[:, 0].astype(float), atol=1e-13)
    assert np.allclose(pulled @ E[:, 1].astype(float), post @ E[:, 1].astype(float), atol=1e-13)


def test_y_mirror_symmetry():
    nx, ny = 48, 25
    # the reference's sponge is off by one cell between bottom (j < w) and top (j > ny - w), so the
    # mirror invariant only holds with the sponge switched off
    cfg = make_config(nx, ny, rho_in=1.02, warmup=10, sponge=(4, 8, 3, 3), strength=0.0)
    o = OracleLBMC(cfg, cylinder_mask(nx, ny, 14, 12, 3), dtype=np.float64)
    o.init()
    o.run_step(120)
    assert np.allclose(o.rho, o.rho[:, ::-1], atol=1e-12)
    assert np.allclose(o.vel[..., 0], o.vel[:, ::-1, 0], atol=1e-12)
    assert np.allclose(o.vel[..., 1], -o.vel[:, ::-1, 1], atol=1e-12)


def test_ramp_values():
    assert ramp_value(1, 1000) == np.float32(1) - np.float32(np.cos(np.float64(np.float32(0.5 * 3.14159265) * np.float32(1) / np.float32(1000))))
    assert abs(float(ramp_value(500, 1000)) - (1 - np.cos(np.pi / 4))) < 1e-6
    assert ramp_value(1000, 1000) == ramp_value(1001, 1000) == ramp_value(10**6, 1000) == np.float32(1.0)
    assert ramp_value(1, 0) == np.float32(1.0)      # warmup_steps = 0 -> inf -> min(1, .) = 1


def test_missing_config_key_raises_keyerror():
    cfg = make_config(16, 8)
    del cfg["simulation"]["ghost_moments_s"]
    with pytest.raises(KeyError):
        OracleLBM(cfg)


def test_c_and_numpy_oracles_agree_on_random_small_configs():
    """Randomised cross-check of the two independent restatements (boundary types incl. no-op ones, solids on the
    ring, LES on / off, warm-up 0): every field bit-identical after 15 steps."""
    rng = np.random.default_rng(7)
    for trial in range(12):
        nx, ny = int(rng.integers(4, 24)), int(rng.integers(3, 20))
        cfg = make_config(nx, ny, bc_type=[int(t) for t in rng.integers(0, 4, 4)],
                          bc_value=[[float(v) for v in rng.uniform(-0.04, 0.04, 2)] for _ in range(4)],
                          rho_in=float(rng.uniform(0.98, 1.04)), rho_out=float(rng.uniform(0.98, 1.02)),
                          nu=float(rng.uniform(0.01, 0.1)), cs=float(rng.choice([0.0, 0.1, 0.17])),
                          warmup=int(rng.integers(0, 10)), sponge=tuple(int(v) for v in rng.integers(0, 4, 4)),
                          strength=float(rng.uniform(0, 3)))
        mask = rng.random((nx, ny)) < 0.1
        a, b = OracleLBM(cfg, mask), OracleLBMC(cfg, mask)
        a.init(), b.init()
        a.run_step(15), b.run_step(15)
        for nm in FIELDS:
            assert np.array_equal(getattr(a, nm), getattr(b, nm), equal_nan=True), (trial, nm)
        assert np.array_equal(a.get_moments_numpy(), b.get_moments_numpy(), equal_nan=True), trial


def test_bounce_back_extension_of_the_numpy_oracle():
    """`obstacle_mode="bounce_back"` is NOT reference behaviour (ref:452-455 refills); it is the checker of the CUDA
    build's optional mode.  Properties of the rule itself: OPP is the index of -e_k; with no solids it changes nothing;
    in a pressure-driven channel between solid slabs the no-slip wall sits half-way between the last fluid and the
    first solid node (y = 2.5 / ny - 3.5 for three solid rows per side), where the reference's refill gives 2.2."""
    from oracle.lbm_oracle_np import OPP
    assert all((E[OPP[k]] == -E[k]).all() for k in range(9))
    cfg = make_config(20, 12, rho_in=1.01, warmup=3)
    a, b = OracleLBM(cfg, None), OracleLBM(cfg, None, obstacle_mode="bounce_back")
    a.init(), b.init()
    a.run_step(20), b.run_step(20)
    assert np.array_equal(a.f_old, b.f_old)
    nx, ny = 24, 22
    cfg = make_config(nx, ny, rho_in=1.0006, rho_out=1.0, nu=0.1, cs=0.0, warmup=0, sponge=(0, 0, 0, 0), strength=0.0)
    mask = np.zeros((nx, ny), bool)
    mask[:, :3] = True
    mask[:, -3:] = True
    roots = {}
    for mode in ("bounce_back", "refill"):
        o = OracleLBM(cfg, mask, dtype=np.float64, obstacle_mode=mode)
        o.init()
        o.run_step(2500)
        roots[mode] = np.sort(np.roots(np.polyfit(np.arange(3, ny - 3), o.vel[nx // 2, 3:ny - 3, 0], 2)))
        if mode == "bounce_back":
            assert np.all(o.rho[mask] == 1.0) and np.all(o.vel[mask] == 0.0)
    assert np.abs(roots["bounce_back"] - [2.5, ny - 3.5]).max() < 0.03
    assert np.abs(roots["refill"] - [2.5, ny - 3.5]).max() > 0.2
