"""GPU parity on the BASELINE.json workloads themselves (benchmarks/workloads.py -- the inputs bench.py, the sweep and
SCALE run), not on look-alikes: SURVEY.md section 8(d) inputs 2-5.

  configs[1]  tube bank 2048x512, 1 000 steps, two export frames at the dataset interval of 500 (production ROI /
              fractional INTER_AREA ratio / save height 256) -- solver state and writer output vs the oracles
  configs[2]  urban 8192x2048, 1 000 steps (the run north_star names, the workload behind the headline number)
  configs[3]  random obstacles: the generator at 4096x1024 on one GPU (the full 32768x8192 grid runs in the
              >= 2 GPU slab test, tests/test_gpu_slab.py)
  configs[4]  three sweep cases end to end through batch.run_cases (replica mode, device writer, result shards)

The default (strict) arithmetic must be bit-identical to the fp32 C oracle; the fast build is reported per channel.
"""
import importlib
import json
import os

import numpy as np
import pytest

from benchmarks import workloads as W
from helpers import fast_arith_report, format_report, rel_linf
from oracle.lbm_oracle_c import OracleLBMC
from oracle.writer_oracle import WriterOracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    return importlib.import_module("01-lbm-2d_b200")


def _bit_exact_state(s, ref, tag):
    assert np.array_equal(s.f_old.to_numpy(), ref.f_old), f"f_old {tag}"
    assert np.array_equal(s.rho.to_numpy(), ref.rho), f"rho {tag}"
    assert np.array_equal(s.vel.to_numpy(), ref.vel), f"vel {tag}"
    assert np.array_equal(s.get_moments_numpy(), ref.get_moments_numpy()), f"moments {tag}"
    assert s.get_max_velocity() == ref.get_max_velocity(), f"max|u| {tag}"


def test_configs1_tube_bank_1000_steps_with_two_export_frames(pkg, tmp_path):
    dwm = importlib.import_module("01-lbm-2d_b200.device_writer")
    ops = importlib.import_module("01-lbm-2d_b200.simulation_ops")
    cfg, mask = W.tube_bank_2048x512()
    nx, ny = cfg["simulation"]["nx"], cfg["simulation"]["ny"]
    assert (nx, ny) == (2048, 512) and cfg["outputs"]["dataset"]["interval_steps"] == 500
    ref, ref64 = OracleLBMC(cfg, mask), OracleLBMC(cfg, mask, dtype=np.float64)
    wo = WriterOracle(cfg, nx, ny)
    ref.init(), ref64.init()
    for _ in range(2):
        ref.run_step(500), ref64.run_step(500)
        wo.append(ref.get_moments_numpy())
    want = wo.finalize()

    s = pkg.LBM2D_MRT_LES(cfg, mask_data=mask)   # default arithmetic = strict
    s.init()
    w = dwm.DeviceLBMCaseWriter(str(tmp_path / "tube.h5"), cfg, nx, ny, mask_data=mask, solver=s)
    assert (w.target_w, w.target_h) == (wo.target_w, wo.target_h) and w.crop_h / w.target_h == 1.5   # fractional ratio
    meta = ops.run_simulation_loop(cfg, s, None, None, None, w, max_steps=1000, progress=False)
    assert meta["status"] == "Success" and meta["final_steps"] == 1000
    _bit_exact_state(s, ref, "tube bank, 1000 steps")
    got = w.finalize()
    assert got["turbulence"].shape == (2, 9, 256, wo.target_w)
    for k in ("turbulence", "mean_vel_field", "mean_vel_sq_field", "sum_vor"):
        assert np.array_equal(got[k], want[k]), k
    for k in ("stats_min", "stats_max", "stats_mean"):
        assert np.array_equal(np.asarray(w.attrs[k]), want[k]), k
    s.close()

    f = pkg.LBM2D_MRT_LES(cfg, mask_data=mask, arith="fast")
    f.init()
    f.run_step(1000)
    assert rel_linf(f.rho.to_numpy(), ref.rho) <= 1e-5 and rel_linf(f.f_old.to_numpy(), ref.f_old) <= 1e-5
    ok, rows = fast_arith_report(f.get_moments_numpy(), f.vel.to_numpy(), ref, ref64)
    print("tube bank fast per channel:", format_report(rows))
    assert ok, rows


def test_configs2_urban_8192x2048_1000_steps_bit_exact(pkg):
    """The headline workload, the run length north_star states, the arithmetic the headline is measured with."""
    cfg, mask = W.urban()
    assert (cfg["simulation"]["nx"], cfg["simulation"]["ny"]) == (8192, 2048)
    s = pkg.LBM2D_MRT_LES(cfg, mask_data=mask)
    s.init()
    ref = OracleLBMC(cfg, mask)
    ref.init()
    for n in (500, 500):                      # the run loop's batches (compute_step_size = 500)
        s.run_step(n)
        ref.run_step(n)
        assert s.get_max_velocity() == ref.get_max_velocity()
    assert np.array_equal(s.rho.to_numpy(), ref.rho)
    assert np.array_equal(s.vel.to_numpy(), ref.vel)
    assert np.array_equal(s.f_old.to_numpy(), ref.f_old)
    m = s.get_moments_numpy()
    assert np.array_equal(m, ref.get_moments_numpy())
    del m
    # fast arithmetic on the same run: field-level tolerance (per-channel floors are measured on configs[0], [1])
    f = pkg.LBM2D_MRT_LES(cfg, mask_data=mask, arith="fast")
    f.init()
    f.run_step(1000)
    assert rel_linf(f.rho.to_numpy(), ref.rho) <= 1e-5
    assert np.abs(f.vel.to_numpy() - ref.vel).max() <= 5e-6      # absolute, lattice units (max|u| ~ 1e-2)


def test_configs3_random_obstacle_generator_4096x1024(pkg):
    cfg, mask = W.random_obstacles(nx=4096, ny=1024, n_shapes=60, seed=1234)
    assert mask.any()
    ref = OracleLBMC(cfg, mask)
    s = pkg.LBM2D_MRT_LES(cfg, mask_data=mask)
    ref.init(), s.init()
    for n in (200, 100):
        ref.run_step(n), s.run_step(n)
    _bit_exact_state(s, ref, "random 4096x1024, 300 steps")


def test_configs4_three_sweep_cases_through_batch_run_cases(pkg, tmp_path):
    """Replica mode on the GPU: rank 0 of 1, three of the 64 procedural cases, device writer, result shards; the
    exported frames and statistics of every case are bit-identical to oracle + writer oracle."""
    batch = importlib.import_module("01-lbm-2d_b200.batch")
    cases = {f"sweep_{s:02d}": W.sweep_case(s) for s in (0, 7, 41)}
    out = str(tmp_path / "sweep")
    res = batch.run_cases(cases, out, rank=0, world=1, device=0, max_steps=600, concurrency=2)
    assert sorted(res) == sorted(cases) and all(r["status"] == "Success" and r["final_steps"] == 600 for r in res.values())
    merged = batch.merge_shards(out, remove=True)
    assert {n: r["status"] for n, r in merged.items()} == {n: "Success" for n in cases}
    assert json.load(open(os.path.join(out, "sim_results.json"))).keys() == merged.keys()
    dwm = importlib.import_module("01-lbm-2d_b200.device_writer")
    for name, (cfg, mask) in cases.items():
        nx, ny = cfg["simulation"]["nx"], cfg["simulation"]["ny"]
        ref = OracleLBMC(cfg, mask)
        ref.init()
        wo = WriterOracle(cfg, nx, ny)
        for _ in range(3):                                    # compute_step_size = interval = 200
            ref.run_step(200)
            wo.append(ref.get_moments_numpy())
        want = wo.finalize()
        got = dwm.read_case(os.path.join(out, name))
        for k in ("turbulence", "mean_vel_field", "mean_vel_sq_field", "sum_vor"):
            assert np.array_equal(got[k], want[k]), (name, k)
        assert got["static_mask"].shape == (2, wo.target_h, wo.target_w)
    # resume: a second session skips all three
    again = batch.run_cases(cases, out, rank=0, world=1, device=0, max_steps=600)
    assert again == {}
