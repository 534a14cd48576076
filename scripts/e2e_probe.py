"""Where does a run-loop batch spend its host time on N slabs?  torchrun --nproc-per-node N scripts/e2e_probe.py"""
import importlib, os, sys, time, json
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dist.init_process_group("nccl", device_id=torch.device("cuda", local))
slab = importlib.import_module("01-lbm-2d_b200.slab"); dwm = importlib.import_module("01-lbm-2d_b200.device_writer")
cfg, mask = bench.build_workload("urban", world)
nx, ny = cfg["simulation"]["nx"], cfg["simulation"]["ny"]
s = slab.SlabLBM(cfg, mask, rank=rank, world=world, device=local); s.init(); s.run_step(200); s.synchronize()
w = dwm.DeviceLBMCaseWriter(f"/tmp/probe_{rank}.h5", cfg, nx, ny, solver=s)
for it in range(3):
    t = [time.perf_counter()]
    s.run_step(500); t.append(time.perf_counter())
    s.synchronize(); t.append(time.perf_counter())
    f = s.get_force(); t.append(time.perf_counter())
    v = s.get_max_velocity(); t.append(time.perf_counter())
    g = s.export_frame_gathered(); t.append(time.perf_counter())
    if rank == 0:
        names = ["enqueue run_step(500)", "gpu wait", "get_force", "get_max_velocity", "export_frame_gathered"]
        print(it, {n: round((b - a) * 1e3, 2) for n, a, b in zip(names, t, t[1:])}, flush=True)
w.close(); s.close(); dist.destroy_process_group()
