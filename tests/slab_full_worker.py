"""Worker of tests/test_gpu_slab.py::test_configs3_full_grid_on_slabs: BASELINE configs[3] at its stated size
(32768x8192 random obstacles, benchmarks/workloads.py) on N x-slabs, 10 steps, against the single-GPU kernel run on rank 0
(which is bit-identical to the CPU oracle at sizes the oracle handles: tests/test_gpu_workloads.py uses the same
generator).  The 9.7 GB population field is compared through per-slab checksums of its bit patterns (sum and xor of the
uint32 words, computed slab by slab on both sides), plus max|u| and the force.
Usage: torchrun --nproc-per-node N tests/slab_full_worker.py"""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from benchmarks import workloads as W  # noqa: E402


def checksum(a):
    w = np.ascontiguousarray(a).view(np.uint32).ravel()
    return int(w.sum(dtype=np.uint64)), int(np.bitwise_xor.reduce(w))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module("01-lbm-2d_b200")
    slab = importlib.import_module("01-lbm-2d_b200.slab")
    cfg, mask = W.random_obstacles()
    nx, ny = cfg["simulation"]["nx"], cfg["simulation"]["ny"]
    assert (nx, ny) == (32768, 8192)
    s = slab.SlabLBM(cfg, mask, rank=rank, world=world, device=local)
    s.init()
    s.run_step(7)
    s.run_step(3)
    maxv, force = s.get_max_velocity(), s.get_force()
    mine = checksum(s.solver.f_old.to_numpy())
    halo = s.halo_path
    s.close()
    sums = [None] * world
    dist.all_gather_object(sums, mine)
    if rank == 0:
        m = pkg.LBM2D_MRT_LES(cfg, mask_data=mask, device=local)
        m.init()
        m.run_step(10)
        f = m.f_old.to_numpy()
        for r, (x0, n) in enumerate(slab.partition(nx, world)):
            assert checksum(f[x0:x0 + n]) == sums[r], f"slab {r} differs from the single-GPU result"
        assert m.get_max_velocity() == maxv
        fm = m.get_force()
        assert np.allclose(fm, force, rtol=1e-4, atol=2e-6), (fm, force)   # a cancellation of O(1) link terms, summed per slab
        print(f"SLAB-FULL-OK world={world} halo={halo} grid={nx}x{ny} steps=10 maxv={maxv:.6f}")
        m.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
