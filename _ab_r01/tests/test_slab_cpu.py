"""Multi-rank path on CPU (gloo, world_size 2 and 3): the product's slab plumbing
(`partition`, `halo_plan`, `exchange_host`, `reduce_force`, `reduce_max_velocity`) moving halos between
slab instances of the numpy oracle must reproduce the monolithic oracle bit for bit.
The GPU path uses the same partition / plan and exchanges the same three populations per direction
inside lbm_run() over NCCL (tests/test_gpu_slab.py)."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import cylinder_mask, make_config, random_blocks_mask
from oracle.lbm_oracle_np import OracleLBM

slab = importlib.import_module("01-lbm-2d_b200.slab")


def test_partition_and_plan():
    assert slab.partition(10, 3) == [(0, 4), (4, 3), (7, 3)]
    assert slab.partition(8192, 4) == [(0, 2048), (2048, 2048), (4096, 2048), (6144, 2048)]
    with pytest.raises(ValueError):
        slab.partition(5, 3)
    assert slab.halo_plan(0, 3) == [(1, "E", (1, 5, 8), (3, 6, 7))]
    assert slab.halo_plan(1, 3) == [(2, "E", (1, 5, 8), (3, 6, 7)), (0, "W", (3, 6, 7), (1, 5, 8))]
    assert slab.halo_plan(2, 3) == [(1, "W", (3, 6, 7), (1, 5, 8))]


def _case():
    nx, ny = 37, 18
    cfg = make_config(nx, ny, rho_in=1.02, nu=0.02, warmup=9, sponge=(4, 7, 2, 2))
    mask = cylinder_mask(nx, ny, 11, 9, 3) | random_blocks_mask(nx, ny, 5, seed=5, smin=1, smax=4, keep_in=0, keep_out=0)
    mask[12, :3] = True   # touches the bottom wall right at a slab interface (world=3: 13 | 12 | 12)
    mask[24:26, ny - 2:] = True
    return cfg, mask


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, steps, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg, mask = _case()
    x0, n = slab.partition(cfg["simulation"]["nx"], world)[rank]
    o = OracleLBM(cfg, mask, slab=(x0, n))
    o.init()
    for _ in range(steps):
        o.run_step(1)
        slab.exchange_host(dist, rank, world, o.halo_pack, o.halo_unpack)
    own = slice(o._own0, o._own0 + n)
    force = slab.reduce_force(dist, o.get_force())
    maxv = slab.reduce_max_velocity(dist, o.get_max_velocity_owned() if hasattr(o, "get_max_velocity_owned") else
                                    float(np.sqrt((o.vel[own] ** 2).sum(-1)).max()))
    nanmax = slab.reduce_max_velocity(dist, float("nan") if rank == world - 1 else 0.1)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), f_old=o.f_old[own], f_new=o.f_new[own], rho=o.rho[own],
             vel=o.vel[own], moments=o.get_moments_numpy()[own], force=force, maxv=maxv, nanmax=nanmax)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_oracles_with_halo_exchange_equal_monolithic(world, tmp_path):
    steps = 40
    mp.spawn(_worker, args=(world, _free_port(), steps, str(tmp_path)), nprocs=world, join=True)
    cfg, mask = _case()
    ref = OracleLBM(cfg, mask)
    ref.init()
    ref.run_step(steps)
    parts = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    for nm in ("f_old", "f_new", "rho", "vel"):
        got = np.concatenate([p[nm] for p in parts], axis=0)
        assert np.array_equal(got, getattr(ref, nm)), nm
    assert np.array_equal(np.concatenate([p["moments"] for p in parts], axis=0), ref.get_moments_numpy())
    for p in parts:  # every rank holds the global reductions
        assert np.allclose(p["force"], ref.get_force(), rtol=1e-5, atol=1e-7)
        assert float(p["maxv"]) == ref.get_max_velocity()
        assert np.isnan(float(p["nanmax"]))
