"""Worker of tests/test_gpu_slab.py: one rank of an x-slab run (torchrun, NCCL), checked on rank 0
against the monolithic CPU oracle.  Usage: torchrun --nproc-per-node N tests/slab_worker.py [kernel]"""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import cylinder_mask, force_f64, make_config, random_blocks_mask, rel_linf  # noqa: E402
from oracle.lbm_oracle_c import OracleLBMC  # noqa: E402


def main():
    kernel = sys.argv[1] if len(sys.argv) > 1 else "auto"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    slab = importlib.import_module("01-lbm-2d_b200.slab")
    nx, ny = 203, 130
    cfg = make_config(nx, ny, rho_in=1.02, nu=0.015, warmup=25, sponge=(6, 20, 4, 4))
    mask = cylinder_mask(nx, ny, 50, 60, 9) | random_blocks_mask(nx, ny, 10, seed=9, smin=2, smax=9, keep_in=0, keep_out=0)
    for x0, _ in slab.partition(nx, world)[1:]:
        mask[x0 - 1:x0 + 1, 40:48] = True   # solids straddling every interface
        mask[x0, :2] = True
    dwm = importlib.import_module("01-lbm-2d_b200.device_writer")
    from oracle.writer_oracle import WriterOracle

    cfg["domain_zones"]["buffer"] = 3
    cfg["outputs"]["dataset"]["save_resolution_height"] = 29
    for arith in ("strict", "fast"):
        s = slab.SlabLBM(cfg, mask, rank=rank, world=world, device=local, arith=arith, kernel=kernel)
        s.init()
        writer = dwm.DeviceLBMCaseWriter(os.path.join("/tmp", f"slab_case_{rank}.h5"), cfg, nx, ny, solver=s)
        for n in (1, 10, 49):
            s.run_step(n)
            writer.append_from_solver(s)
        exported = writer.finalize()
        fields = {nm: s.gather(getattr(s.solver, nm).to_numpy()) for nm in ("f_old", "f_new", "rho", "vel")}
        fields["moments"] = s.gather(s.get_moments_numpy())
        force, maxv = s.get_force(), s.get_max_velocity()
        assert s.step_count() == 60
        if rank == 0:
            ref = OracleLBMC(cfg, mask)
            ref.init()
            wo = WriterOracle(cfg, nx, ny)
            for n in (1, 10, 49):
                ref.run_step(n)
                wo.append(ref.get_moments_numpy())
            want_export = wo.finalize()
            for key in ("turbulence", "mean_vel_field", "mean_vel_sq_field", "sum_vor"):
                if arith == "strict":
                    assert np.array_equal(exported[key], want_export[key]), ("export", key)
                else:
                    assert np.abs(exported[key] - want_export[key]).max() <= 1e-5 * max(1.0, np.abs(want_export[key]).max()), ("export", key)
            if arith == "strict":
                assert np.array_equal(writer.attrs["stats_min"], want_export["stats_min"])
                assert np.array_equal(writer.attrs["stats_max"], want_export["stats_max"])
            want = dict(f_old=ref.f_old, f_new=ref.f_new, rho=ref.rho, vel=ref.vel, moments=ref.get_moments_numpy())
            F, S = force_f64(ref.f_new, ref.mask)
            if arith == "strict":
                for nm, a in want.items():
                    assert np.array_equal(fields[nm], a), (arith, nm, float(np.abs(fields[nm] - a).max()))
                assert maxv == ref.get_max_velocity(), (maxv, ref.get_max_velocity())
                assert np.abs(force - F).max() <= 2e-6 * S + 1e-9, (force, F)
            else:
                for nm in ("f_old", "rho", "moments"):
                    assert rel_linf(fields[nm], want[nm]) <= 1e-5, (arith, nm)
                assert np.abs(fields["vel"] - ref.vel).max() <= 2e-6
                assert np.abs(force - F).max() <= 2e-5 * S + 1e-9
            print(f"SLAB-OK world={world} arith={arith} kernel={kernel} halo={s.halo_path} maxv={maxv:.6f} F={force}")
        s.close()
    # ---- video-frame fields on slabs: filter / gradient reach across the slab borders (strict state from above is gone:
    # a fresh short run), and the optional bounce-back obstacle mode on slabs against the numpy oracle's rule
    from oracle import viz_oracle
    from oracle.lbm_oracle_np import OracleLBM

    if kernel != "tma":
        s = slab.SlabLBM(cfg, mask, rank=rank, world=world, device=local)
        s.init()
        s.run_step(40)
        vel = s.gather(s.vel.to_numpy())
        for sigma in (1.0, 2.5, 0.0):
            mag, vor = s.get_viz_fields(sigma)
            mag, vor = s.gather(mag), s.gather(vor)
            if rank == 0:
                want_mag, want_vor = viz_oracle.viz_fields(vel, sigma)
                assert np.array_equal(mag, want_mag) and np.array_equal(vor, want_vor), ("viz", sigma)
        s.close()
        bcfg = make_config(nx, ny, rho_in=1.02, nu=0.03, warmup=5, sponge=(6, 20, 4, 4))
        s = slab.SlabLBM(bcfg, mask, rank=rank, world=world, device=local, obstacle_mode="bounce_back")
        s.init()
        s.run_step(45)
        f_old, rho = s.gather(s.solver.f_old.to_numpy()), s.gather(s.rho.to_numpy())
        if rank == 0:
            ref = OracleLBM(bcfg, mask, obstacle_mode="bounce_back")
            ref.init()
            ref.run_step(45)
            assert np.array_equal(f_old, ref.f_old) and np.array_equal(rho, ref.rho), "bounce-back on slabs"
            print(f"SLAB-EXTRAS-OK world={world} halo={s.halo_path}")
        s.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
