"""Replica mode: a batch of independent cases, one process per GPU (BASELINE configs[4]).

The reference's `pipeline/batch_run.py:219` walks the sorted config list sequentially on one device and
records a status per case in `sim_results.json` (`io/sim_results_io.py:133-172`, a read-modify-write that is
not multi-writer safe).  Here rank r of `world` takes cases r, r+world, ... of the same sorted list, runs each
through the reference's run loop (`simulation_ops.run_simulation_loop`) with the device-side writer, and
writes its own shard `sim_results.rank{r}.json`; rank 0 merges the shards after a barrier -- no shared
file is ever written by two processes.  There is no data-path collective: the cases are independent.

Resume semantics are the reference's (`batch_run.py:78-116, 219-351`): a case recorded as `Success` or
`Failed` is skipped, one left `Running` by a crashed session is retried; `Running` is written BEFORE a
case starts; a case whose run loop does not end in `Success` is recorded as `Failed` with its reason and
its output file is removed (`case_executor.py:105-107, 151-160`); `max_success` stops a rank once the
successes of earlier sessions plus its share of the new ones reach the quota.

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


def merge_shards(out_dir, world=None, remove=False):
    """Fold every rank shard into sim_results.json (atomic replace).  `remove`: delete the shards afterwards --
    rank 0 does this at the start of a session (`consolidate`), before any rank writes a new shard."""
    merged = {}
    try:   # records of earlier sessions stay unless this session re-ran the case
        with open(os.path.join(out_dir, "sim_results.json")) as f:
            merged.update(json.load(f))
    except (OSError, ValueError):
        pass
    for path in _shard_files(out_dir):   # every shard present (an earlier session may have used more ranks)
        try:
            with open(path) as f:
                for name, rec in json.load(f).items():
                    if not (merged.get(name, {}).get("status") == "Success" and rec.get("status") != "Success"):
                        merged[name] = rec
        except (OSError, ValueError):
            continue
    tmp = os.path.join(out_dir, "sim_results.json.tmp")
    with open(tmp, "w") as f:
        json.dump(dict(sorted(merged.items())), f, indent=2)
    os.replace(tmp, os.path.join(out_dir, "sim_results.json"))  # atomic, like sim_results_io.py:55-66
    if remove:
        for path in _shard_files(out_dir):
            os.remove(path)
    return merged


def consolidate(out_dir):
    """Session start, ONE process (rank 0, before the barrier that releases the others): shards left by a
    crashed or differently sized earlier session become part of sim_results.json."""
    os.makedirs(out_dir, exist_ok=True)
    return merge_shards(out_dir, remove=True) if _shard_files(out_dir) else None


def _shard_files(out_dir):
    if not os.path.isdir(out_dir):
        return []
    return sorted(os.path.join(out_dir, f) for f in os.listdir(out_dir)
                  if f.startswith("sim_results.rank") and f.endswith(".json"))


def load_status_map(out_dir):
    """{case: status} from sim_results.json (sim_results_io.py:117-130)."""
    try:
        with open(os.path.join(out_dir, "sim_results.json")) as f:
            return {name: rec.get("status") for name, rec in json.load(f).items()}
    except (OSError, ValueError):
        return {}


def resume_plan(names, status_map):
    """(already_success, skip) as batch_run.py:78-116: Success and Failed are skipped, Running is retried."""
    skip, ok = set(), 0
    for name in names:
        st = status_map.get(name)
        if st == "Success":
            skip.add(name)
            ok += 1
        elif st == "Failed":
            skip.add(name)
    return ok, skip


def _gpu_runner(name, cfg, mask, out_dir, device, max_steps, progress):
    """One case through the reference's run loop with the device-side writer."""
    pkg = importlib.import_module("01-lbm-2d_b200")
    ops = importlib.import_module("01-lbm-2d_b200.simulation_ops")
    dwm = importlib.import_module("01-lbm-2d_b200.device_writer")
    t = [time.perf_counter()]
    solver = pkg.LBM2D_MRT_LES(cfg, mask_data=mask, device=device)
    try:
        solver.init()
        t.append(time.perf_counter())
        writer = dwm.DeviceLBMCaseWriter(os.path.join(out_dir, f"{name}.h5"), cfg, solver.nx, solver.ny,
                                         mask_data=mask, solver=solver)
        t.append(time.perf_counter())
        meta = ops.run_simulation_loop(cfg, solver, None, None, None, writer,
                                       max_steps or cfg["simulation"]["max_steps"], progress=progress)
        t.append(time.perf_counter())
        writer.close()
        t.append(time.perf_counter())
    finally:
        solver.close()
    if os.environ.get("LBM2D_CASE_TIMING"):   # where a case's wall time goes (ms): create + init | writer + static mask | run loop | file close
        print(f"[case {name}] " + " | ".join(f"{(b - a) * 1e3:.1f}" for a, b in zip(t, t[1:])), file=sys.stderr)
    return meta


def _remove_outputs(out_dir, name):
    for ext in (".h5", ".npz", ".turbulence.f32"):   # case_executor.py:_cleanup_failed_outputs
        try:
            os.remove(os.path.join(out_dir, name + ext))
        except OSError:
            pass


def run_cases(cases, out_dir, rank=0, world=1, device=None, max_steps=None, progress=False, *,
              resume=True, max_success=None, runner=None, concurrency=1):
    """cases: {name: (config, mask)}.  Returns this rank's {name: result} (skipped cases keep their old record
    in the merged file and are not in the returned dict).

    `concurrency` > 1 runs that many of the rank's cases at a time on the same GPU, one host thread and one CUDA
    stream (one solver handle) each: the sweep grids are a fraction of a wave of CTAs and a case spends most of its
    wall time on the host (mask SDF, frame stacking, file output), so the cases overlap each other's host work and
    fill the GPU's idle SMs.  Cases still START in the sorted order."""
    import threading

    runner = runner or _gpu_runner
    os.makedirs(out_dir, exist_ok=True)
    if world == 1:
        consolidate(out_dir)   # with several ranks the launcher does this once, before the start barrier (main)
    status_map = load_status_map(out_dir) if resume else {}
    already_success, skip = resume_plan(sorted(cases), status_map)
    quota = None
    if max_success is not None:   # batch_run.py:201-213, 233-241; the remaining quota is split over the ranks
        remaining = max(0, max_success - already_success)
        quota = remaining // world + (1 if rank < remaining % world else 0)
    shard_path = os.path.join(out_dir, f"sim_results.rank{rank}.json")
    results = {}
    todo = [n for n in shard(list(cases), rank, world) if n not in skip]
    state = {"next": 0, "success": 0, "in_flight": 0}
    cond = threading.Condition()

    def flush():   # called with the lock held
        tmp = shard_path + ".tmp"
        with open(tmp, "w") as f:
            json.dump(results, f, indent=2)
        os.replace(tmp, shard_path)

    def claim():
        """Next case of this rank, or None when the list or the success quota is exhausted."""
        with cond:
            while True:
                if state["next"] >= len(todo) or (quota is not None and state["success"] >= quota):
                    return None
                if quota is not None and state["success"] + state["in_flight"] >= quota:
                    cond.wait()   # the cases in flight may still fail: wait for one of them before starting more
                    continue
                name = todo[state["next"]]
                state["next"] += 1
                state["in_flight"] += 1
                results[name] = {"status": "Running", "rank": rank}   # crash-safe pre-write, batch_run.py:253-258
                flush()
                return name

    def worker():
        while (name := claim()) is not None:
            cfg, mask = cases[name]
            t0 = time.perf_counter()
            ok = False
            try:
                meta = dict(runner(name, cfg, mask, out_dir, device, max_steps, progress))
                if meta.get("status") != "Success":   # case_executor.py:105-107
                    raise RuntimeError(f"Simulation failed: {meta.get('reason', meta.get('status'))}")
                ok = True
            except Exception as e:  # a failed case must not take the batch down (case_executor.py:151-160)
                _remove_outputs(out_dir, name)
                meta = {"status": "Failed", "reason": str(e), "final_steps": 0}
            meta["wall_time_s"] = round(time.perf_counter() - t0, 2)
            meta["rank"] = rank
            with cond:
                results[name] = meta
                state["in_flight"] -= 1
                state["success"] += int(ok)
                flush()
                cond.notify_all()

    threads = [threading.Thread(target=worker, name=f"lbm-case-{i}") for i in range(max(1, int(concurrency)) - 1)]
    # A batch of a sweep case is ~1 ms of GPU work and a handful of short Python steps between blocking C calls; with
    # the interpreter's default 5 ms switch interval a case thread coming back from such a call can wait that long for
    # the lock while another thread runs Python code (frame bookkeeping, file metadata) -- several batches' worth.
    old_interval = sys.getswitchinterval()
    if threads:
        sys.setswitchinterval(min(old_interval, 2e-4))
    try:
        for t in threads:
            t.start()
        worker()
        for t in threads:
            t.join()
    finally:
        sys.setswitchinterval(old_interval)
    with cond:
        flush()
    return results


def run_sweep(n_cases=64, out_dir="gpurun_out/sweep", max_steps=None, max_success=None, resume=False, concurrency=2,
              emit=print):
    """BASELINE configs[4]: `n_cases` procedural 1024x256 cases (seeds 0 .. n-1), one process per GPU (torchrun), every
    case through the run loop with the device-side writer.  Rank 0 emits ONE JSON line: cases/hour over all ranks,
    process start-up (CUDA context, imports) excluded and reported.  Also what `bench.py --workload sweep` runs."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from benchmarks import workloads as W

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cases = {f"sweep_{s:02d}": W.sweep_case(s) for s in range(n_cases)}
    # process start-up (CUDA context, module load, cv2 / scipy imports) is paid once per worker, not per case:
    # keep it out of the cases/hour window and report it separately
    t_start = time.perf_counter()
    import cv2  # noqa: F401
    import scipy.ndimage  # noqa: F401
    first = next(iter(cases.values()))
    warm = importlib.import_module("01-lbm-2d_b200").LBM2D_MRT_LES(first[0], mask_data=first[1], device=local)
    warm.init()
    warm.run_step(2)
    warm.get_max_velocity()
    # a GPU that has been idle takes a second or two to leave its low-power clocks; a sweep runs for minutes to hours, so
    # the cases/hour window starts on a GPU that is already busy (like process start-up, reported as excluded time)
    t_busy = time.perf_counter()
    while time.perf_counter() - t_busy < float(os.environ.get("LBM2D_SWEEP_WARM_S", "2.0")):
        warm.run_step(2000)
        warm.get_max_velocity()
    warm.close()
    startup_s = time.perf_counter() - t_start
    if rank == 0:
        if not resume:   # a benchmark run starts from nothing
            import shutil

            shutil.rmtree(out_dir, ignore_errors=True)
        elif world > 1:
            consolidate(out_dir)
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    run_cases(cases, out_dir, rank, world, device=local, max_steps=max_steps, resume=resume, max_success=max_success,
              concurrency=concurrency)
    if dist is not None:
        dist.barrier()
    dt = time.perf_counter() - t0
    line = None
    if rank == 0:
        before = load_status_map(out_dir)
        merged = merge_shards(out_dir, world, remove=True)
        ran_now = [n for n in merged if before.get(n) not in ("Success", "Failed")]   # skipped cases cost no time
        ok = sum(1 for r in merged.values() if r["status"] == "Success")
        steps = sum(merged[n].get("final_steps", 0) for n in ran_now)
        cfg0 = first[0]
        line = {"metric": "cases/hour (procedural 1024x256 sweep incl. export, BASELINE configs[4])", "value": len(ran_now) / dt * 3600,
                "unit": "cases/hour", "n_gpus": world, "higher_is_better": True, "scaling": "weak" if n_cases % max(world, 1) == 0 else "strong",
                "dtype": "f32", "data": "synthetic", "concurrency": concurrency, "cases": len(ran_now), "cases_recorded": len(merged),
                "success": ok, "total_steps": steps, "steps_per_case": steps // max(1, len(ran_now)), "wall_s": dt,
                "startup_s_excluded": startup_s, "mlups_aggregate": steps * cfg0["simulation"]["nx"] * cfg0["simulation"]["ny"] / dt / 1e6,
                "config": {"workload": f"{n_cases} procedural masks 1024x256 (circles / rotated squares / triangles), replica mode: one "
                                       f"process per GPU, {concurrency} cases in flight per GPU, run loop + device-side writer + file output",
                           "arith": "strict"}}
        emit(json.dumps(line))
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sweep", type=int, default=64, help="number of synthetic 1024x256 cases (seeds 0..N-1)")
    ap.add_argument("--out", default="gpurun_out/sweep")
    ap.add_argument("--max-steps", type=int, default=None)
    ap.add_argument("--max-success", type=int, default=None)
    ap.add_argument("--no-resume", action="store_true")
    ap.add_argument("--concurrency", type=int, default=1, help="cases in flight per GPU (threads, one stream each)")
    args = ap.parse_args()
    run_sweep(args.sweep, args.out, args.max_steps, args.max_success, not args.no_resume, args.concurrency)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
