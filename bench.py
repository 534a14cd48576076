#!/usr/bin/env python
"""bench.py -- MLUPS of the fused D2Q9 MRT-LES step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--arith strict|fast]

One "step" = one lattice-Boltzmann time step (one fused pass over the whole grid).
N = 1 workload: BASELINE.json configs[2], the 8192x2048 urban LES case -- the configuration the
metric's "% of HBM roofline" is quoted on: its 1.22 GB double-buffered
state is ~10x the 126 MB L2, so consecutive steps cannot be served from cache (configs[0..1] are
L2-resident and are parity-test cases).  N > 1: the same 8192x2048 slab per GPU, stacked along x
into one (8192 N)x2048 domain, one halo column exchanged per step -> "scaling": "weak".

Prints ONE JSON line (rank 0).
  value      whole-job MLUPS, state resident in HBM: the MEDIAN of `windows.n` back-to-back windows of exactly K steps
             each, timed with CUDA events on the solver's stream (max over ranks per window), after an internal
             warm-up of max(W, 200) steps.  All windows are in the line; a single 4 ms window (K = 20) is at the
             mercy of one host hiccup, the median of 25 is not.
  arithmetic the DEFAULT, bit-exact ("strict") arithmetic: `parity_mode` says what that means and `alt_arith` holds
             the same measurement for the optional fast arithmetic (tolerance-level parity).
  e2e        the same metric through the reference-facing Python API the way the reference's run loop drives it
             (batches of compute_step_size steps, get_force + get_max_velocity after each, a moments frame to host
             at the dataset interval).
  roofline   72 algorithmic bytes per cell update over the measured average step time, against
             MEASURED_PEAKS.json.
  cpu_baseline  the C/OpenMP port of the reference's three-pass step (oracle/) on this host's cores, bounded sample.
  slab_parity   (N > 1) slabs vs the single-GPU kernel on a small case, checked before anything is timed.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
ALGO_BYTES_PER_CELL = 72.0  # 9 fp32 reads + 9 fp32 writes
METRIC = "MLUPS (fused D2Q9 MRT-LES step)"
BASELINE_CONFIG = {"cylinder": "configs[0]", "tube_bank": "configs[1]", "urban": "configs[2]", "random": "configs[3]",
                   "sweep_case": "configs[4], one case"}
N_WINDOWS = 25
MIN_WARMUP_STEPS = 200


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons sampled in-process through NVML (a thread polling every ~2 ms) DURING the timed region."""

    def __init__(self, index=0):
        self.index, self.rows, self.on, self.thread, self.err = index, [], False, None, None
        self.t0 = self.t1 = None
        try:
            import pynvml

            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, repr(e)

    def _poll(self):
        nv = self.nv
        while self.on:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((time.time(), float(mhz), int(reasons)))
            except Exception as e:  # pragma: no cover
                self.err = repr(e)
                break
            time.sleep(0.002)

    def start(self):
        if self.nv is None:
            return
        self.on = True
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def mark(self, which):
        if which == 0:
            self.t0 = time.time()
        else:
            self.t1 = time.time()

    def stop(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"NVML unavailable: {self.err}"], "samples": 0}
        self.on = False
        self.thread.join(timeout=1)
        nv = self.nv
        rows = [r for r in self.rows if self.t0 is not None and self.t0 <= r[0] <= (self.t1 or r[0])]
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        seen = sorted(n for n, bit in names.items() if any(r[2] & bit for r in rows))
        return {"sm_mhz": statistics.median(r[1] for r in rows) if rows else None, "sm_max_mhz": self.max_mhz,
                "reasons": seen, "samples": len(rows), "sampled_over": "timed region (in-process NVML, 2 ms period)"}


def build_workload(name, n_gpus):
    from benchmarks import workloads as W

    if name == "random":  # BASELINE configs[3]: fixed 32768x8192 grid -> strong scaling over the slabs
        return W.random_obstacles()
    if name == "sweep_case":  # one of the 64 procedural 1024x256 cases of BASELINE configs[4] (step timing)
        return W.sweep_case(0)
    if name == "urban" and n_gpus > 1:
        # weak scaling: every GPU gets THE configs[2] block pattern (the 8192x2048 mask tiled along x), so the work per
        # GPU -- solid fraction included -- is exactly that of the N = 1 run; inlet on the first slab, outlet on the last
        cfg, mask = W.urban()
        cfg = json.loads(json.dumps(cfg))
        cfg["simulation"]["nx"] = 8192 * n_gpus
        cfg["simulation"]["name"] = f"urban_{8192 * n_gpus}x2048"
        return cfg, np.ascontiguousarray(np.tile(mask, (n_gpus, 1)))
    return W.WORKLOADS[name]()


def workload_label(name, nx, ny, world):
    """`config.workload` of both arms (the reference arm times one GPU's share of it on the host)."""
    return (f"{name} {nx}x{ny} (BASELINE {BASELINE_CONFIG.get(name, 'configs[2]')}" +
            (" per GPU, its mask tiled along x" if world > 1 else "") + "), D2Q9 MRT-LES Cs=0.1, bc [0,2,1,2]")


def time_cpu_port(cfg, mask, budget_s, threads=None, steps=None):
    """The reference's three-pass step as restated in oracle/lbm_oracle.c, all host threads (cpu_baseline leg)."""
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
    sample = f"{args.warmup}+{args.steps} full-grid steps of {nx}x{ny}" + (f" (one GPU's share of the {args.gpus}-GPU workload)" if args.gpus > 1 else "")
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
            "workload": workload_label(args.workload, nx_full * max(1, args.gpus), ny, args.gpus),
            "grid": [nx, ny], "parallelism": f"host CPU, {cores} OpenMP threads (the reference's three-pass step, C port)",
            "sample": sample, "arith": "strict fp32 (reference evaluation order)",
        },
        "cpu_baseline": {"value": mlups, "unit": "MLUPS", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": mlups, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def slab_parity_check(pkg, rank, world, local_rank, dist):
    """Slabs vs the single-GPU kernel (itself bit-identical to the CPU oracle: tests/test_gpu_parity.py) on the
    203x130 case of tests/slab_worker.py, 60 steps, strict arithmetic, before anything is timed."""
    import torch

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import cylinder_mask, make_config, random_blocks_mask

    slabs = importlib.import_module("01-lbm-2d_b200.slab")
    nx, ny = 203, 130
    cfg = make_config(nx, ny, rho_in=1.02, nu=0.015, warmup=25, sponge=(6, 20, 4, 4))
    mask = cylinder_mask(nx, ny, 50, 60, 9) | random_blocks_mask(nx, ny, 10, seed=9, smin=2, smax=9, keep_in=0, keep_out=0)
    for x0, _ in slabs.partition(nx, world)[1:]:
        mask[x0 - 1:x0 + 1, 40:48] = True   # solids straddling every interface
    s = slabs.SlabLBM(cfg, mask, rank=rank, world=world, device=local_rank, arith="strict")
    s.init()
    for n in (1, 10, 49):
        s.run_step(n)
    f_old = s.gather(s.solver.f_old.to_numpy())
    maxv = s.get_max_velocity()
    s.close()
    ok = 1
    if rank == 0:
        m = pkg.LBM2D_MRT_LES(cfg, mask_data=mask, arith="strict", device=local_rank)
        m.init()
        m.run_step(60)
        ok = int(np.array_equal(m.f_old.to_numpy(), f_old) and m.get_max_velocity() == maxv)
        m.close()
    t = torch.tensor([ok], device="cuda")
    dist.broadcast(t, src=0)
    if int(t.item()) != 1:
        raise SystemExit("slab parity check FAILED: slabs differ from the single-GPU kernel; refusing to time a wrong result")
    return {"result": "bit-exact (strict)", "check": f"x-slabs x{world} vs single-GPU kernel, 203x130, 60 steps, f and max|u|; "
            "the single-GPU strict kernel is bit-identical to the CPU oracle (tests/test_gpu_parity.py)"}


def timed_windows(solver, stream, steps, n_windows, barrier, world, dist):
    """n_windows back-to-back windows of exactly `steps` steps, CUDA events on the solver's stream.
    Returns (per-window ms: max over ranks, per-rank medians)."""
    import torch

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n_windows + 1)]
    barrier()
    ev[0].record(stream)
    for i in range(n_windows):
        solver.run_step(steps)
        ev[i + 1].record(stream)
    barrier()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(n_windows)]
    per_rank = None
    if world > 1:
        t = torch.tensor(ms, device="cuda", dtype=torch.float64)
        allr = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allr, t)
        per_rank = [float(r.median().item()) for r in allr]
        ms = [float(v) for v in torch.stack(allr).max(dim=0).values.tolist()]
    return ms, per_rank


def run_ours(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pkg = importlib.import_module("01-lbm-2d_b200")
    slab_parity = slab_parity_check(pkg, rank, world, local_rank, dist) if world > 1 else None

    cfg, mask = build_workload(args.workload, world)
    if args.grid:
        from benchmarks import workloads as W

        gx, gy = (int(v) for v in args.grid.lower().split("x"))
        cfg, mask = W.urban(nx=gx, ny=gy, seed=1, x_lo=gx // 32, x_hi_margin=gx // 8, n_rects=max(4, gx * gy // 170000))
    if os.environ.get("BENCH_BC"):  # experiments: boundary types, e.g. BENCH_BC=0313
        cfg["boundary_condition"]["type"] = [int(c) for c in os.environ["BENCH_BC"]]
    nx, ny = cfg["simulation"]["nx"], cfg["simulation"]["ny"]

    def make_solver(arith):
        if world > 1:
            slabs = importlib.import_module("01-lbm-2d_b200.slab")
            return slabs.SlabLBM(cfg, mask, rank=rank, world=world, device=local_rank, arith=arith, kernel=args.kernel)
        return pkg.LBM2D_MRT_LES(cfg, mask_data=mask, arith=arith, kernel=args.kernel, device=local_rank)

    def make_barrier(solver):
        def barrier():
            solver.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
        return barrier

    def measure(arith, sampler=None):
        solver = make_solver(arith)
        solver.init()
        stream = torch.cuda.ExternalStream(solver.device_view().stream, device=torch.device("cuda", local_rank))
        barrier = make_barrier(solver)
        warm = max(MIN_WARMUP_STEPS, args.warmup, 3)
        solver.run_step(warm)
        barrier()
        l0 = solver.launch_count()
        if sampler:
            sampler.mark(0)
        ms, per_rank = timed_windows(solver, stream, args.steps, args.windows, barrier, world, dist)
        if sampler:
            sampler.mark(1)
        launches = (solver.launch_count() - l0) / args.windows
        return solver, barrier, warm, ms, per_rank, launches

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    solver, barrier, warm, win_ms, per_rank, launches = measure(args.arith, sampler)
    clocks = sampler.stop() if rank == 0 else None
    ms = statistics.median(win_ms)
    mlups = nx * ny * args.steps / (ms * 1e-3) / 1e6
    peak, peak_src = measured_hbm_peak()
    cells_per_gpu = nx * ny / world

    def roof(ms_window):
        return ALGO_BYTES_PER_CELL * cells_per_gpu / (ms_window * 1e-3 / args.steps) / 1e9

    # ---- end to end through the reference-facing API: the reference's run loop (simulation_ops.py:87-209) --
    # batches of compute_step_size steps, get_force + get_max_velocity (stability fuse) after each, and at the
    # dataset interval an export frame to host memory.  Primary number: the repo's writer path, where the
    # frame is cropped / INTER_AREA-resized / accumulated on the device (DeviceLBMCaseWriter) and only the
    # (9, H, W) frame crosses PCIe.  Secondary: the reference's unmodified writer contract, a full (nx, ny, 9)
    # host array per export (get_moments_numpy -> a fresh caller-owned array backed by a pinned pool buffer).
    css = cfg["simulation"]["compute_step_size"]
    interval = cfg["outputs"]["dataset"]["interval_steps"]
    n_batches = 1 if args.quick else 2
    ops = importlib.import_module("01-lbm-2d_b200.simulation_ops")
    e2e = {}
    for label in (("device_writer", "full_frame") if not args.quick else ("device_writer",)):
        if label == "device_writer":
            dwm = importlib.import_module("01-lbm-2d_b200.device_writer")
            writer = dwm.DeviceLBMCaseWriter(os.path.join(tempfile.gettempdir(), f"lbm_bench_case_r{rank}.h5"), cfg, nx, ny, solver=solver)
            frame_bytes = 9 * writer.target_w * writer.target_h * 4   # whole job; each rank holds a column range
        else:
            class _FullFrame:  # what the reference's AsyncLBMCaseWriter receives: it keeps the array until written
                n = 0

                def append(self, m):
                    self.n += m.nbytes

            writer = _FullFrame()
            frame_bytes = nx * ny * 9 * 4   # whole job, nx*ny*9*4/world per rank
            solver.get_moments_numpy()      # first use allocates the pinned pool buffer (once per process)
        if label == "device_writer":   # untimed warm-up of this path (first export: NCCL channels, allocations, file open)
            warm_w = dwm.DeviceLBMCaseWriter(os.path.join(tempfile.gettempdir(), f"lbm_bench_warm_r{rank}.h5"), cfg, nx, ny, solver=solver)
            ops.run_simulation_loop(cfg, solver, None, None, None, warm_w, max_steps=css, progress=False)
            warm_w.close()
            writer.attach(solver)      # reset the device-side statistics for the timed case
        barrier()
        t0 = time.perf_counter()
        meta = ops.run_simulation_loop(cfg, solver, None, None, None, writer, max_steps=n_batches * css, progress=False)
        barrier()
        dt = time.perf_counter() - t0
        if label == "device_writer":
            writer.close()
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
                              f"{n_frames} export frame(s) of {frame_bytes} B to host ({label}); a simulation has no per-step "
                              "input, so h2d is 0; solver construction (16.8 MB mask upload, link / bit-plane build) is "
                              "once per case and outside the timed region"}
    e2e_main = e2e.get("device_writer") or e2e.get("full_frame")
    halo_path = getattr(solver, "halo_path", None)
    halo_path = {"peer": "peer-memory stores from the step kernel (NVLink, CUDA IPC), one launch per step",
                 "nccl": "grouped ncclSend/ncclRecv, edge columns on a side stream"}.get(halo_path, halo_path)
    solver.close()

    # ---- the other arithmetic, same measurement (no e2e legs)
    alt = None
    if not args.quick and not args.no_alt:
        other = "fast" if args.arith == "strict" else "strict"
        s2, _, _, ms2, _, _ = measure(other)
        s2.close()
        m2 = statistics.median(ms2)
        alt = {"arith": other, "mlups": nx * ny * args.steps / (m2 * 1e-3) / 1e6, "ms_per_step": m2 / args.steps,
               "frac": roof(m2) / peak,
               "parity": "bit-identical to the fp32 oracle" if other == "strict" else
               "rho, f <= 1e-5; moments and u per channel <= max(1e-5, 3 x fp32 noise floor) (tests/helpers.py::fast_arith_report)"}

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    cpu_mlups, cores, cpu_n, cpu_dt = (time_cpu_port(*build_workload(args.workload, 1), budget_s=15.0)
                                       if world == 1 and not args.quick else (None, None, None, None))
    strict = args.arith == "strict"
    line = {
        "metric": METRIC, "value": mlups, "unit": "MLUPS", "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": workload_label(args.workload, nx, ny, world),
            "grid": [nx, ny], "parallelism": "single GPU" if world == 1 else f"x-slabs x{world}, 1 halo column / step, halo path: {halo_path}",
            "l2_policy": "working set 1.22 GB per GPU >> 126 MB L2: inputs larger than L2, no flush needed",
            "arith": args.arith, "kernel": args.kernel, "solid_fraction": float(mask.mean()),
        },
        "windows": {"n": args.windows, "steps_each": args.steps, "statistic": "median", "ms": [round(v, 5) for v in win_ms],
                    "min_ms": min(win_ms), "max_ms": max(win_ms), "per_rank_median_ms": per_rank},
        "parity_mode": {"arith": args.arith, "mlups": mlups, "frac": roof(ms) / peak,
                        "parity": "bit-identical to the fp32 oracle (rho, u, f, 9 moments, max|u|): tests/test_gpu_parity.py, "
                                  "tests/test_gpu_workloads.py run THIS workload for 1 000 steps" if strict else
                                  "tolerance-level (see alt_arith for the bit-exact arithmetic)"},
        "alt_arith": alt,
        "e2e": e2e_main,
        "e2e_reference_writer_path": e2e.get("full_frame") if e2e_main is not e2e.get("full_frame") else None,
        "gpu_launches": int(round(launches)),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": roof(ms), "peak": peak, "unit": "GB/s", "frac": roof(ms) / peak,
                     "traffic": None,
                     "traffic_note": "not measured in-process; ncu --set full capture of this command: profiles/r02_step_kernel_ncu.md",
                     "peak_source": peak_src,
                     "kernel": f"lbm::step_kernel<STRICT={str(strict).lower()}, EMIT=false> (K-1 of K launches per window; the K-th is the EMIT variant)",
                     "algorithmic_bytes_per_launch": ALGO_BYTES_PER_CELL * cells_per_gpu},
    }
    if slab_parity:
        line["slab_parity"] = slab_parity
    if cpu_mlups is not None:
        line["cpu_baseline"] = {"value": cpu_mlups, "unit": "MLUPS", "cores": cores, "kind": "port",
                                "sample": f"{cpu_n} full-grid steps of {args.workload} ({cpu_dt:.1f} s)"}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--windows", type=int, default=N_WINDOWS, help="back-to-back timed windows of --steps steps each")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="urban", choices=["urban", "cylinder", "tube_bank", "random", "sweep_case", "sweep"])
    ap.add_argument("--grid", default=None, help="NXxNY: urban-style obstacles on a custom grid (experiments)")
    ap.add_argument("--quick", action="store_true", help="timed region only (profiling runs): no e2e full-frame / alt / cpu legs")
    ap.add_argument("--no-alt", action="store_true", help="skip the measurement of the other arithmetic")
    ap.add_argument("--concurrency", type=int, default=2, help="--workload sweep: cases in flight per GPU")
    ap.add_argument("--arith", default="strict", choices=["fast", "strict"])
    ap.add_argument("--kernel", default="auto", choices=["auto", "register", "tma"])
    args = ap.parse_args()
    if args.workload == "sweep":   # BASELINE configs[4]: cases/hour in replica mode (no slabs, no collective on the data path)
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "the sweep metric (cases/hour incl. export) has no CPU arm here; "
                              "use the default workload for the reference arm"}))
            return
        batch = importlib.import_module("01-lbm-2d_b200.batch")
        batch.run_sweep(n_cases=32 * max(1, int(os.environ.get("WORLD_SIZE", "1"))), out_dir=os.path.join(tempfile.gettempdir(), "lbm_sweep"),
                        concurrency=args.concurrency)
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            import torch.distributed as dist

            dist.destroy_process_group()
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
