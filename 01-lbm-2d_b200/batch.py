"""Replica mode: a batch of independent cases, one process per GPU (BASELINE configs[4]).

The reference's `pipeline/batch_run.py:219` walks the sorted config list sequentially on one device and
records a status per case in `sim_results.json` (`io/sim_results_io.py:133-172`, a read-modify-write that is
not multi-writer safe).  Here rank r of `world` takes cases r, r+world, ... of the same sorted list, runs each
through the reference's run loop (`simulation_ops.run_simulation_loop`) with the device-side writer, and
writes its own shard `sim_results.rank{r}.json`; rank 0 merges the shards after a barrier -- no shared
file is ever written by two processes.  There is no data-path collective: the cases are independent.

    torchrun --nproc-per-node 8 01-lbm-2d_b200/batch.py --sweep 64 --out outputs/sweep
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import time


def shard(names, rank, world):
    """Cases of this rank: every world-th entry of the SORTED list (the reference's order, batch_run.py:48)."""
    return sorted(names)[rank::world]


def merge_shards(out_dir, world):
    merged = {}
    for r in range(world):
        path = os.path.join(out_dir, f"sim_results.rank{r}.json")
        if os.path.exists(path):
            with open(path) as f:
                merged.update(json.load(f))
    tmp = os.path.join(out_dir, "sim_results.json.tmp")
    with open(tmp, "w") as f:
        json.dump(dict(sorted(merged.items())), f, indent=2)
    os.replace(tmp, os.path.join(out_dir, "sim_results.json"))  # atomic, like sim_results_io.py:55-66
    return merged


def run_cases(cases, out_dir, rank=0, world=1, device=None, max_steps=None, progress=False):
    """cases: {name: (config, mask)}.  Returns this rank's {name: result}."""
    pkg = importlib.import_module("01-lbm-2d_b200")
    ops = importlib.import_module("01-lbm-2d_b200.simulation_ops")
    dwm = importlib.import_module("01-lbm-2d_b200.device_writer")
    os.makedirs(out_dir, exist_ok=True)
    results = {}
    for name in shard(list(cases), rank, world):
        cfg, mask = cases[name]
        t0 = time.perf_counter()
        try:
            solver = pkg.LBM2D_MRT_LES(cfg, mask_data=mask, device=device)
            solver.init()
            writer = dwm.DeviceLBMCaseWriter(os.path.join(out_dir, f"{name}.h5"), cfg, solver.nx, solver.ny,
                                             mask_data=mask, solver=solver)
            meta = ops.run_simulation_loop(cfg, solver, None, None, None, writer,
                                           max_steps or cfg["simulation"]["max_steps"], progress=progress)
            writer.close()
            solver.close()
        except Exception as e:  # a failed case must not take the batch down (case_executor.py:151-160)
            meta = {"status": "Error", "reason": str(e), "final_steps": 0}
        meta["wall_time_s"] = time.perf_counter() - t0
        meta["rank"] = rank
        results[name] = meta
        with open(os.path.join(out_dir, f"sim_results.rank{rank}.json"), "w") as f:
            json.dump(results, f, indent=2)
    return results


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sweep", type=int, default=64, help="number of synthetic 1024x256 cases (seeds 0..N-1)")
    ap.add_argument("--out", default="gpurun_out/sweep")
    ap.add_argument("--max-steps", type=int, default=None)
    args = ap.parse_args()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from benchmarks import workloads as W

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cases = {f"sweep_{s:02d}": W.sweep_case(s) for s in range(args.sweep)}
    t0 = time.perf_counter()
    run_cases(cases, args.out, rank, world, device=local, max_steps=args.max_steps)
    if dist is not None:
        dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        merged = merge_shards(args.out, world)
        ok = sum(1 for r in merged.values() if r["status"] == "Success")
        steps = sum(r["final_steps"] for r in merged.values())
        print(json.dumps({"metric": "cases/hour (64 x 1024x256 sweep incl. export)", "value": len(merged) / dt * 3600,
                          "n_gpus": world, "cases": len(merged), "success": ok, "total_steps": steps, "wall_s": dt}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
