#!/usr/bin/env python
"""Generate tests/golden/ref_loop_*.npz: the UNMODIFIED reference run loop
(`/root/reference/src/lbm_mrt_les/core/simulation_ops.py:60 run_simulation_loop`) driving the UNMODIFIED reference solver
(`core/LBM2D_MRT_LES.py`, under the Taichi stand-in) and the UNMODIFIED reference writer (`io/lbm_writer.py
LBMCaseWriter`, with tests/fake_h5py.py standing in for h5py and the real cv2 / scipy) -- the call sequence of
`pipeline/run_one_case.py:48-64,135`.  What is recorded: every dataset and attribute of the case file plus the
metadata dict the loop returns.

tests/test_gpu_reference_loop.py then runs THIS repo's class, loop mirror and DeviceLBMCaseWriter on the same config and
mask on the GPU and compares bit for bit (the GPU box has no /root/reference; the build container has no GPU -- the
fixture is the bridge).  tests/test_reference_loop_cpu.py re-runs the reference loop + writer here with the oracle as the
solver and checks it against the same fixture.

    python tests/golden/gen/make_ref_loop_fixture.py        # ~1 minute, needs /root/reference
"""
from __future__ import annotations

import importlib
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, TESTS)
REF_SRC = "/root/reference/src"


def install_stand_ins():
    """taichi -> fake_taichi, h5py -> in-memory stand-in, matplotlib -> empty stub (imported by utils, never called)."""
    import fake_h5py
    import fake_taichi

    fake_taichi.install()
    sys.modules["h5py"] = fake_h5py
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    return fake_h5py


def reference_modules():
    ops = importlib.import_module("lbm_mrt_les.core.simulation_ops")
    writer = importlib.import_module("lbm_mrt_les.io.lbm_writer")
    return ops, writer


def case():
    """96x44 channel with a cylinder and a block, sponges on all sides, ROI at a fractional INTER_AREA ratio (34 -> 20
    rows), data every 15 steps from step 30 on, 90 steps."""
    from helpers import cylinder_mask, make_config

    nx, ny = 96, 44
    cfg = make_config(nx, ny, rho_in=1.02, nu=0.02, cs=0.1, warmup=20, sponge=(6, 14, 3, 3), buffer=2, save_h=20,
                      compute_step_size=15, name="ref_loop_case")
    cfg["outputs"]["start_record_step"] = 30
    mask = cylinder_mask(nx, ny, 30, 22, 5)
    mask[52:58, 8:15] = True
    return cfg, mask, 90


def run_reference(solver_factory, path):
    fake_h5py = install_stand_ins()
    ops, writer_mod = reference_modules()
    cfg, mask, max_steps = case()
    nx, ny = cfg["simulation"]["nx"], cfg["simulation"]["ny"]
    solver = solver_factory(cfg, mask)
    solver.init()                                                                    # run_one_case.py:49
    writer = writer_mod.LBMCaseWriter(path, cfg, nx, ny, mask_data=mask)             # run_one_case.py:135 (sync flavour)
    meta = ops.run_simulation_loop(cfg, solver, None, None, None, writer, max_steps)  # run_one_case.py:152
    writer.close()
    f = fake_h5py.FILES[path]
    out = {f"ds_{k}": np.asarray(f[k][...]) for k in f.keys()}
    out.update({f"attr_{k}": np.asarray(v) for k, v in f.attrs.items() if k != "config_json"})
    out["meta_json"] = np.array(json.dumps(meta))
    out["config_json"] = np.array(json.dumps(cfg))
    out["mask"] = mask
    out["max_steps"] = np.array(max_steps)
    return out


def main():
    sys.path.insert(0, HERE)
    import make_ti_shim_fixtures as shim

    ref_cls = shim.load_reference_class()

    class RefSolver(ref_cls):
        def __init__(self, *a, **kw):
            super().__init__(*a, **kw)
            # same typing rule as make_ti_shim_fixtures.py: values the reference assigns to kernel locals are f32 in
            # Taichi (`rho_in = self.rho_in_target`), so `rho_in - 1.0` is an f32 subtraction
            self.rho_in_target = np.float32(self.rho_in_target)
            self.rho_out_target = np.float32(self.rho_out_target)

        # ti.atomic_max on a kernel-local scalar is the one construct the stand-in cannot emulate (a Python float is not
        # a reference); the value only feeds the stability fuse and the progress bar, never the datasets
        def get_max_velocity(self):
            v = self.vel.to_numpy()
            return float(np.sqrt(v[..., 0] ** 2 + v[..., 1] ** 2).max())

    out = run_reference(lambda cfg, mask: RefSolver(cfg, mask_data=mask), "/tmp/ref_loop_case.h5")
    dst = os.path.join(os.path.dirname(HERE), "ref_loop_case.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
