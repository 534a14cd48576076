#!/usr/bin/env python
"""bench.py -- MLUPS of the fused D2Q9 MRT-LES step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one lattice-Boltzmann time step (one fused pass over the whole grid).
N = 1 workload: BASELINE.json configs[2], the 8192x2048 urban LES case -- the configuration the
metric's "% of HBM roofline" is quoted on: its 1.22 GB double-buffered
state is ~10x the 126 MB L2, so consecutive steps cannot be served from cache (configs[0..1] are
L2-resident and are parity-test cases).  N > 1: the same 8192x2048 slab per GPU, stacked along x
into one (8192 N)x2048 domain, one halo column exchanged per step -> "scaling": "weak".

Prints ONE JSON line (rank 0).  `value`: whole-job MLUPS, state resident in HBM, timed with CUDA
events on the solver's stream around exactly K steps.  `e2e`: the same metric through the
reference-facing Python API the way the reference's run loop drives it (batches of
compute_step_size steps, get_force + get_max_velocity after each, a moments frame to host numpy
at the dataset interval).  `roofline`: 72 algorithmic bytes per cell update over
the measured average step-kernel time, against MEASURED_PEAKS.json.  `cpu_baseline`: the C/OpenMP
port of the reference's three-pass step (oracle/) on this host's cores, bounded sample.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
ALGO_BYTES_PER_CELL = 72.0  # 9 fp32 reads + 9 fp32 writes
METRIC = "MLUPS (fused D2Q9 MRT-LES step)"


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def mark(self, which):
        """wall-clock bounds of the timed region (samples are filtered to it)"""
        if which == 0:
            self.t0 = time.time()
        else:
            self.t1 = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")][1:]))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], None, set()
        rows = [r for (t, r) in self.rows if self.t0 is None or (self.t0 - 0.02 <= t <= (self.t1 or t) + 0.05)]
        where = "timed region"
        if len(rows) < 3:  # the timed region was shorter than a few nvidia-smi sampling periods
            rows, where = [r for (_, r) in self.rows], "warm-up + timed region"
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "sampled_over": where}


def build_workload(name, n_gpus):
    from benchmarks import workloads as W

    if name == "random":  # BASELINE configs[3]: fixed 32768x8192 grid -> strong scaling over the slabs
        return W.random_obstacles()
    if name == "urban" and n_gpus > 1:
        return W.urban(nx=8192 * n_gpus, ny=2048, seed=1, n_rects=100 * n_gpus, max_attempts=400 * n_gpus)
    return W.WORKLOADS[name]()


def time_cpu_port(cfg, mask, budget_s, threads=None, steps=None):
    """The reference's three-pass step as restated in oracle/lbm_oracle.c, all host threads."""
    from oracle import lbm_oracle_c

    cores = threads or (os.cpu_count() or 1)
    lbm_oracle_c.set_threads(cores)
    o = lbm_oracle_c.OracleLBMC(cfg, mask)
    o.init()
    t0 = time.perf_counter()
    o.run_step(1)
    t1 = time.perf_counter() - t0
    n = steps if steps is not None else int(max(2, min(200, budget_s / max(t1, 1e-4))))
    t0 = time.perf_counter()
    o.run_step(n)
    dt = time.perf_counter() - t0
    cells = cfg["simulation"]["nx"] * cfg["simulation"]["ny"]
    return cells * n / dt / 1e6, cores, n, dt


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  Taichi is not installable
    in this image (no wheel, no network), so this is the oracle port (kind = "port"), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg, mask = build_workload(args.workload, 1)
    nx, ny = cfg["simulation"]["nx"], cfg["simulation"]["ny"]
    nx_full = nx
    sample = f"{args.warmup}+{args.steps} full-grid steps of {nx}x{ny}"
    from oracle import lbm_oracle_c

    cores = os.cpu_count() or 1
    lbm_oracle_c.set_threads(cores)
    o = lbm_oracle_c.OracleLBMC(cfg, mask)
    o.init()
    t0 = time.perf_counter()
    o.run_step(1)
    t1 = time.perf_counter() - t0
    if t1 * (args.steps + args.warmup) > 120.0:  # keep the run within a few minutes: crop the slab in x
        frac = max(1, int(120.0 / (t1 * (args.steps + args.warmup)) * nx) // 64 * 64)
        nx_s = max(256, frac)
        cfg = json.loads(json.dumps(cfg))
        cfg["simulation"]["nx"] = nx_s
        mask = np.ascontiguousarray(mask[:nx_s])
        sample = f"{args.warmup}+{args.steps} steps of the first {nx_s} of {nx} columns ({nx_s}x{ny})"
        o = lbm_oracle_c.OracleLBMC(cfg, mask)
        o.init()
        nx = nx_s
    o.run_step(max(0, args.warmup - 1))
    t0 = time.perf_counter()
    o.run_step(args.steps)
    dt = time.perf_counter() - t0
    mlups = nx * ny * args.steps / dt / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": mlups, "unit": "MLUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": f"{args.workload} {nx_full}x{ny} (BASELINE configs[2] per GPU), D2Q9 MRT-LES Cs=0.1, bc [0,2,1,2]",
            "grid": [nx, ny], "parallelism": f"host CPU, {cores} OpenMP threads (the reference's three-pass step, C port)",
            "sample": sample, "arith": "strict fp32 (reference evaluation order)",
        },
        "cpu_baseline": {"value": mlups, "unit": "MLUPS", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": mlups, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_ours(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pkg = importlib.import_module("01-lbm-2d_b200")

    cfg, mask = build_workload(args.workload, world)
    if args.grid:
        from benchmarks import workloads as W

        gx, gy = (int(v) for v in args.grid.lower().split("x"))
        cfg, mask = W.urban(nx=gx, ny=gy, seed=1, x_lo=gx // 32, x_hi_margin=gx // 8, n_rects=max(4, gx * gy // 170000))
    if os.environ.get("BENCH_BC"):  # experiments: boundary types, e.g. BENCH_BC=0313
        cfg["boundary_condition"]["type"] = [int(c) for c in os.environ["BENCH_BC"]]
    nx, ny = cfg["simulation"]["nx"], cfg["simulation"]["ny"]
    if world > 1:
        from importlib import import_module

        slabs = import_module("01-lbm-2d_b200.slab")
        solver = slabs.SlabLBM(cfg, mask, rank=rank, world=world, device=local_rank)
    else:
        solver = pkg.LBM2D_MRT_LES(cfg, mask_data=mask, arith=args.arith, kernel=args.kernel, device=local_rank)
    solver.init()
    view = solver.device_view()
    stream = torch.cuda.ExternalStream(view.stream, device=torch.device("cuda", local_rank))

    def barrier():
        solver.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: exactly K steps between two events on the solver's stream
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)  # let nvidia-smi come up so that it samples the timed region
    solver.run_step(max(3, args.warmup))
    barrier()
    l0 = solver.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark(0)
    e0.record(stream)
    solver.run_step(args.steps)
    e1.record(stream)
    barrier()
    sampler.mark(1)
    ms = e0.elapsed_time(e1)
    launches = solver.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    mlups = nx * ny * args.steps / (ms * 1e-3) / 1e6

    # ---- end to end through the reference-facing API: the reference's run loop (simulation_ops.py:87-209) --
    # batches of compute_step_size steps, get_force + get_max_velocity (stability fuse) after each, and at the
    # dataset interval an export frame to host memory.  Primary number: the repo's writer path, where the
    # frame is cropped / INTER_AREA-resized / accumulated on the device (DeviceLBMCaseWriter) and only the
    # (9, H, W) frame crosses PCIe.  Secondary: the reference's unmodified writer contract, a full (nx, ny, 9)
    # host array per export (get_moments_numpy).
    css = cfg["simulation"]["compute_step_size"]
    interval = cfg["outputs"]["dataset"]["interval_steps"]
    n_batches = 1 if args.quick else 2
    ops = importlib.import_module("01-lbm-2d_b200.simulation_ops")
    e2e = {}
    for label in (("device_writer", "full_frame") if not args.quick else ("device_writer",)):
        writer = None
        if label == "device_writer":
            dwm = importlib.import_module("01-lbm-2d_b200.device_writer")
            writer = dwm.DeviceLBMCaseWriter(os.path.join(ROOT, "gpurun_out", "bench_case.h5"), cfg, nx, ny, solver=solver)
            lo, hi = solver.export_columns if hasattr(solver, "export_columns") else (0, writer.target_w)
            frame_bytes = 9 * writer.target_w * writer.target_h * 4   # whole job; this rank holds columns [lo, hi)
        else:
            class _FullFrame:  # what the reference's AsyncLBMCaseWriter receives
                n = 0

                def append(self, m):
                    self.n += m.nbytes

            writer = _FullFrame()
            frame_bytes = nx * ny * 9 * 4   # whole job, nx*ny*9*4/world per rank
        barrier()
        t0 = time.perf_counter()
        meta = ops.run_simulation_loop(cfg, solver, None, None, None, writer, max_steps=n_batches * css, progress=False)
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        done = meta["final_steps"]
        n_frames = done // interval
        e2e[label] = {"value": nx * ny * done / dt / 1e6, "unit": "MLUPS", "h2d_bytes_per_step": 0,
                      "d2h_bytes_per_step": (n_batches * 12 + n_frames * frame_bytes) / max(1, done),
                      "status": meta["status"],
                      "what": f"run_simulation_loop: {n_batches} x [run_step({css}) + get_force + get_max_velocity] + "
                              f"{n_frames} export frame(s) of {frame_bytes} B to host ({label})"}
    e2e_main = e2e.get("device_writer") or e2e.get("full_frame")

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    peak, peak_src = measured_hbm_peak()
    cells_per_gpu = nx * ny / world
    avg_kernel_s = ms * 1e-3 / args.steps
    achieved = ALGO_BYTES_PER_CELL * cells_per_gpu / avg_kernel_s / 1e9
    cpu_mlups, cores, cpu_n, cpu_dt = (time_cpu_port(*build_workload(args.workload, 1), budget_s=15.0)
                                       if world == 1 and not args.quick else (None, None, None, None))
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture of this grid
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            tj = json.load(f)
        if tj["grid"] == [int(nx // world), int(ny)] and args.arith == "fast" and args.kernel == "auto":
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": mlups, "unit": "MLUPS", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": f"{args.workload} {nx}x{ny} (BASELINE configs[2] per GPU), D2Q9 MRT-LES Cs=0.1, bc [0,2,1,2]",
            "grid": [nx, ny], "parallelism": "single GPU" if world == 1 else f"x-slabs x{world}, 1 halo column / step",
            "l2_policy": "working set 1.22 GB per GPU >> 126 MB L2: inputs larger than L2, no flush needed",
            "arith": args.arith, "kernel": args.kernel, "solid_fraction": float(mask.mean()),
        },
        "e2e": e2e_main,
        "e2e_reference_writer_path": e2e.get("full_frame") if e2e_main is not e2e.get("full_frame") else None,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "kernel": "lbm::step_kernel<STRICT=false, EMIT=false, V=2> (K-1 of K launches; the K-th is the EMIT variant)",
                     "algorithmic_bytes_per_launch": ALGO_BYTES_PER_CELL * cells_per_gpu},
    }
    if cpu_mlups is not None:
        line["cpu_baseline"] = {"value": cpu_mlups, "unit": "MLUPS", "cores": cores, "kind": "port",
                                "sample": f"{cpu_n} full-grid steps of {args.workload} ({cpu_dt:.1f} s)"}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="urban", choices=["urban", "cylinder", "tube_bank", "random"])
    ap.add_argument("--grid", default=None, help="NXxNY: urban-style obstacles on a custom grid (experiments)")
    ap.add_argument("--quick", action="store_true", help="timed region only (profiling runs): no e2e / cpu_baseline legs")
    ap.add_argument("--arith", default="fast", choices=["fast", "strict"])
    ap.add_argument("--kernel", default="auto", choices=["auto", "register", "tma", "register2", "register1", "async"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
