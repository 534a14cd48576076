"""Driver entry points: build() compiles every native artefact, smoke() runs one tiny case on cuda:0
and checks it against the oracle."""
from __future__ import annotations

import importlib
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _pkg():
    return importlib.import_module("01-lbm-2d_b200")


def build() -> None:
    """nvcc -gencode arch=compute_100a,code=sm_100a the CUDA library in-tree, compile the C oracle
    (the checker), import the package.  Works without a GPU (cross-compilation)."""
    pkg = _pkg()
    path = pkg.build_library(force=True)
    assert os.path.exists(path), path
    from oracle import lbm_oracle_c

    lbm_oracle_c.build(force=True)
    # oracle/_ref: the reference is pure Python + Taichi (no compilable sources) -> nothing to build
    pkg.load_library()


def smoke() -> None:
    """One small cylinder case on cuda:0 through the C ABI, strict build bit-compared and fast build
    tolerance-compared against the CPU oracle."""
    import numpy as np

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import cylinder_mask, make_config, rel_linf
    from oracle.lbm_oracle_c import OracleLBMC

    pkg = _pkg()
    nx, ny, steps = 256, 96, 200
    cfg = make_config(nx, ny, rho_in=1.01, nu=0.01, warmup=100, sponge=(8, 24, 4, 4), name="smoke")
    mask = cylinder_mask(nx, ny, 64, 48, 8)
    ref, ref64 = OracleLBMC(cfg, mask), OracleLBMC(cfg, mask, dtype=np.float64)
    for o in (ref, ref64):
        o.init()
        o.run_step(steps)
    for arith in ("strict", "fast"):
        s = pkg.LBM2D_MRT_LES(cfg, mask_data=mask, arith=arith, device=0)
        s.init()
        s.run_step(steps)
        rho, vel, mom = s.rho.to_numpy(), s.vel.to_numpy(), s.get_moments_numpy()
        force, maxv = s.get_force(), s.get_max_velocity()
        if arith == "strict":
            assert np.array_equal(rho, ref.rho) and np.array_equal(vel, ref.vel), "strict build is not bit-exact"
            assert np.array_equal(mom, ref.get_moments_numpy())
            assert maxv == ref.get_max_velocity()
        errs = (rel_linf(rho, ref.rho), rel_linf(vel, ref.vel), rel_linf(mom, ref.get_moments_numpy()))
        assert max(errs[0], errs[2]) <= 1e-5, (arith, errs)
        # u: as close to the fp64 arbiter as the reference-order fp32 arithmetic is (see tests/test_gpu_parity.py)
        assert rel_linf(vel, ref64.vel) <= 2.0 * rel_linf(ref.vel, ref64.vel) + 1e-6, (arith, errs)
        assert np.allclose(force, ref.get_force(), rtol=1e-4, atol=1e-6), (force, ref.get_force())
        print(f"smoke[{arith}]: rel-Linf rho/vel/moments = {errs[0]:.2e}/{errs[1]:.2e}/{errs[2]:.2e}, "
              f"max|u| = {maxv:.5f}, F = {force}, launches = {s.launch_count()}")
        s.close()


if __name__ == "__main__":
    build()
    if len(sys.argv) > 1 and sys.argv[1] == "smoke":
        smoke()
