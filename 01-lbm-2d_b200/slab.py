"""x-slab domain decomposition of one large grid over the GPUs of a node.

The reference is single-device; this is what the B200 build adds.  One process per GPU
(`torchrun`), rank r owns global columns `partition(nx, world)[r]` plus one halo column towards each
neighbour.  Every step each interface moves the three populations that stream across it, in both
directions (f1,f5,f8 eastward; f3,f6,f7 westward; `ny` floats each) -- inside `lbm_run`, by grouped
`ncclSend/ncclRecv` on the solver's stream (NVLink), so a batch of N steps is still ONE host call.
Boundary ownership: inlet on rank 0, outlet on the last rank, top/bottom rows on every rank.  The
per-batch scalars (force, max|u|) are reduced here with torch.distributed.

Halo transport: on NVLink / NVSwitch nodes the slabs map each other's population buffers through CUDA IPC
(`halo="auto"` -> "peer") and the step kernel itself stores the outgoing populations into the neighbour's halo
column and synchronises through per-step counters -- ONE launch per step, no NCCL call on the step path
(include/lbm2d.h, lbm_peer_connect).  `halo="nccl"` keeps the grouped ncclSend/ncclRecv exchange (also the
automatic fallback when peer access is missing).

`torch.distributed` is plumbing only (rendezvous, id / handle exchange, two scalar reductions per batch).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _capi
from .solver import LBM2D_MRT_LES

EAST_GOING = (1, 5, 8)  # e_x = +1: pulled from column i-1, so they travel west -> east
WEST_GOING = (3, 6, 7)


def partition(nx: int, world: int):
    """Contiguous column ranges [(x0, n), ...] west to east; the remainder goes to the first ranks."""
    if world < 1 or nx < 2 * world:
        raise ValueError(f"cannot split nx={nx} into {world} slabs of at least 2 columns")
    base, rem = divmod(nx, world)
    out, x0 = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((x0, n))
        x0 += n
    return out


def halo_plan(rank: int, world: int):
    """[(neighbour_rank, side, planes_sent, planes_received)] for this rank."""
    plan = []
    if rank + 1 < world:
        plan.append((rank + 1, "E", EAST_GOING, WEST_GOING))
    if rank > 0:
        plan.append((rank - 1, "W", WEST_GOING, EAST_GOING))
    return plan


def exchange_host(dist, rank, world, pack, unpack):
    """Host-side halo exchange (numpy through torch.distributed point-to-point; gloo on CPU).
    `pack(side) -> (ny, 3) array`, `unpack(side, array)`.  Used by the CPU tests of the decomposition;
    the GPU path exchanges inside lbm_run()."""
    import torch

    reqs, bufs = [], []
    for nb, side, _, _ in halo_plan(rank, world):
        out = torch.from_numpy(np.ascontiguousarray(pack(side)))
        inc = torch.empty_like(out)
        reqs.append(dist.isend(out, dst=nb))
        reqs.append(dist.irecv(inc, src=nb))
        bufs.append((side, inc, out))
    for r in reqs:
        r.wait()
    for side, inc, _ in bufs:
        unpack(side, inc.numpy())


def reduce_force(dist, local_force, device=None):
    """Sum of the per-slab momentum-exchange forces (each solid cell is owned by exactly one rank)."""
    import torch

    t = torch.tensor(np.asarray(local_force, np.float64), device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy().astype(np.float32)


def reduce_max_velocity(dist, local_max, device=None):
    """Global max|u|; NaN on any rank gives NaN (the stability fuse relies on it)."""
    import torch

    v = float(local_max)
    t = torch.tensor([0.0 if np.isnan(v) else v, 1.0 if np.isnan(v) else 0.0], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t = t.cpu()
    return float("nan") if t[1].item() > 0 else float(t[0].item())


class _DeviceArray:
    """A raw device pointer as `__cuda_array_interface__`, so that torch.distributed can send it (no copy)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2, "strides": None}


class SlabLBM:
    """`LBM2D_MRT_LES` for one slab of a decomposed domain: same methods; `get_force` / `get_max_velocity`
    return GLOBAL values on every rank, field getters return this rank's owned columns
    (`gather_*` assemble the global array on rank 0)."""

    def __init__(self, config, mask_data=None, *, rank, world, device=None, arith="strict", kernel="auto", dist=None,
                 halo="auto", obstacle_mode="refill"):
        if dist is None:
            import torch.distributed as dist
        self.dist, self.rank, self.world = dist, rank, world
        self._pin_rings, self._gather_states = {}, {}
        nx = config["simulation"]["nx"]
        self.slabs = partition(nx, world)
        self.x0, self.nx_owned = self.slabs[rank]
        self.solver = LBM2D_MRT_LES(config, mask_data, arith=arith, kernel=kernel, device=device,
                                    slab=(self.x0, self.nx_owned) if world > 1 else None, obstacle_mode=obstacle_mode)
        self._device = None
        if world > 1:
            import torch

            self._device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else None
            lib = self.solver._lib
            ident = np.zeros(_capi.COMM_ID_BYTES, np.uint8)
            if rank == 0:
                _capi.check(lib.lbm_comm_unique_id(ident.ctypes.data_as(C.c_void_p)))
            t = torch.from_numpy(ident).to(self._device) if self._device is not None else torch.from_numpy(ident)
            dist.broadcast(t, src=0)
            ident = t.cpu().numpy()
            _capi.check(lib.lbm_comm_connect(self.solver._h, rank, world, ident.ctypes.data_as(C.c_void_p)))
            self.halo_path = "nccl"
            halo = os.environ.get("LBM2D_HALO", halo)   # experiments: force the NCCL exchange
            if halo != "nccl" and self._device is not None and kernel in ("auto", "register"):
                self._connect_peers(lib, torch)
        for name in ("nx", "ny", "Re", "name", "nu", "tau_0", "characteristic_length", "rho_in_target",
                     "rho_out_target", "C_smag", "warmup_steps"):
            setattr(self, name, getattr(self.solver, name))
        self.vel, self.rho, self.mask = self.solver.vel, self.solver.rho, self.solver.mask

    def _connect_peers(self, lib, torch):
        """Peer-memory halo path: exchange CUDA IPC handles of the population buffers, map the neighbours' (see
        include/lbm2d.h).  All ranks take the same path: if any rank cannot map a neighbour, all keep NCCL."""
        blob = np.zeros(_capi.PEER_HANDLE_BYTES, np.uint8)
        _capi.check(lib.lbm_peer_export(self.solver._h, blob.ctypes.data_as(C.c_void_p)))
        mine = torch.from_numpy(blob).to(self._device)
        allb = [torch.empty_like(mine) for _ in range(self.world)]
        self.dist.all_gather(allb, mine)
        west = np.ascontiguousarray(allb[self.rank - 1].cpu().numpy()) if self.rank > 0 else None
        east = np.ascontiguousarray(allb[self.rank + 1].cpu().numpy()) if self.rank + 1 < self.world else None
        can = torch.ones(1, device=self._device)
        try:   # probe first: peer access must exist towards both neighbours on every rank
            for nb in (west, east):
                if nb is not None:
                    dev = int(np.frombuffer(nb[192:204].tobytes(), np.int32)[2])
                    if dev != torch.cuda.current_device() and not torch.cuda.can_device_access_peer(torch.cuda.current_device(), dev):
                        can.zero_()
        except Exception:
            can.zero_()
        self.dist.all_reduce(can, op=self.dist.ReduceOp.MIN)
        if can.item() < 1:
            return
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
        _capi.check(lib.lbm_peer_connect(self.solver._h, p(west), p(east)))
        self._peer_keepalive = (west, east)
        self.halo_path = "peer"

    def init(self):
        self.solver.init()
        if self.world > 1:   # nobody steps (and pushes halo columns) before every rank's buffers hold the initial state
            self.solver.synchronize()
            self.dist.barrier()

    def run_step(self, steps=1):
        self.solver.run_step(steps)

    def get_force(self):
        f = self.solver.get_force()
        return f if self.world == 1 else reduce_force(self.dist, f, self._device)

    def get_max_velocity(self):
        v = self.solver.get_max_velocity()
        return v if self.world == 1 else reduce_max_velocity(self.dist, v, self._device)

    def get_moments_numpy(self):
        return self.solver.get_moments_numpy()

    def get_physical_fields(self):
        return self.solver.get_physical_fields()

    def get_viz_fields(self, sigma=None):
        """(|u|, vorticity) of this rank's owned columns; the filter and the gradients reach into the neighbours'
        columns (exchanged inside the call), so `gather()` of the parts equals the single-GPU fields bit for bit."""
        return self.solver.get_viz_fields(sigma)

    # ---- on-device export reduction: ROI / target are global, every rank holds a column range of the frame ----
    def export_configure(self, x0, x1, y0, y1, target_w, target_h):
        self.solver.export_configure(x0, x1, y0, y1, target_w, target_h)
        self.export_columns = self.solver.export_columns

    def export_frame(self, want_frame=True):
        return self.solver.export_frame(want_frame)

    def export_stats(self):
        return self.solver.export_stats()

    def export_frame_gathered(self):
        """One export frame assembled on rank 0 (None elsewhere).  Over NCCL the ranks' column ranges never touch their
        own hosts: the frame stays on the GPU (`lbm_export_frame_device`), is gathered GPU to GPU into a buffer that
        was sized once, and only rank 0 copies the whole frame to (page-locked) host memory."""
        if self.world == 1 or self._device is None:
            return self.gather_columns(self.solver.export_frame())
        import torch

        ptr, shape = self.solver.export_frame_device()
        st = self._gather_state(shape, torch)
        if shape[-1]:
            local = torch.as_tensor(_DeviceArray(ptr, shape), device=self._device)
            st["pad"][..., :shape[-1]].copy_(local)
            torch.cuda.current_stream().synchronize()   # the handle's frame buffer is free for the next export
        self.dist.gather(st["pad"], st["parts"], dst=0)
        if self.rank != 0:
            return None
        torch.cat([p[..., :w] for p, w in zip(st["parts"], st["widths"])], dim=-1, out=st["full"])
        return self._to_pinned(st["full"], torch)

    def _gather_state(self, shape, torch):
        key = tuple(shape[:-1])
        st = self._gather_states.get(key)
        if st is None:   # once per export geometry: widths of all ranks, the padded send buffer, rank 0's receive buffers
            w = torch.zeros(self.world, dtype=torch.int64, device=self._device)
            w[self.rank] = shape[-1]
            self.dist.all_reduce(w)
            widths = [int(v) for v in w.tolist()]
            wmax = max(max(widths), 1)
            pad = torch.zeros(key + (wmax,), dtype=torch.float32, device=self._device)
            parts = [torch.empty_like(pad) for _ in range(self.world)] if self.rank == 0 else None
            full = torch.empty(key + (sum(widths),), dtype=torch.float32, device=self._device) if self.rank == 0 else None
            st = self._gather_states[key] = {"widths": widths, "pad": pad, "parts": parts, "full": full}
        return st

    def _to_pinned(self, full, torch):
        """One D2H copy into page-locked memory (a pageable destination runs at ~4 GB/s: 25 ms for the 86 MB frame of 8
        slabs).  The frame is handed to the writer thread, whose queue holds at most 5: a ring of 7 buffers is never
        overwritten while in use (5 queued + 1 being written + the one being filled)."""
        key = (tuple(full.shape), full.dtype)
        ring = self._pin_rings.get(key)
        if ring is None:   # page-locking is slow (~20 ms per buffer): all seven at the first frame, i.e. during start-up
            ring = self._pin_rings[key] = {"bufs": [torch.empty(full.shape, dtype=full.dtype, pin_memory=True) for _ in range(7)], "next": 0}
        host = ring["bufs"][ring["next"] % 7]
        ring["next"] += 1
        host.copy_(full)
        return host.numpy()

    def gather_columns(self, local):
        """Concatenate per-rank arrays along their LAST axis (output columns) on rank 0 (None elsewhere).
        Over NCCL the arrays travel as tensors padded to the widest rank (one collective, no pickling)."""
        if self.world == 1:
            return local
        if self._device is None:   # gloo / CPU tests
            parts = [None] * self.world if self.rank == 0 else None
            self.dist.gather_object(local, parts, dst=0)
            return np.concatenate(parts, axis=-1) if self.rank == 0 else None
        import torch

        local = np.ascontiguousarray(local)
        widths = torch.zeros(self.world, dtype=torch.int64, device=self._device)
        widths[self.rank] = local.shape[-1]
        self.dist.all_reduce(widths)
        widths = [int(w) for w in widths.tolist()]
        wmax = max(max(widths), 1)
        pad = torch.zeros(local.shape[:-1] + (wmax,), dtype=torch.from_numpy(local).dtype, device=self._device)
        if local.shape[-1]:
            pad[..., :local.shape[-1]] = torch.from_numpy(local).to(self._device)
        parts = [torch.empty_like(pad) for _ in range(self.world)] if self.rank == 0 else None
        self.dist.gather(pad, parts, dst=0)
        if self.rank != 0:
            return None
        return self._to_pinned(torch.cat([p[..., :w] for p, w in zip(parts, widths)], dim=-1), torch)

    def gather(self, local):
        """Concatenate per-rank owned-column arrays along x on rank 0 (None elsewhere)."""
        if self.world == 1:
            return local
        parts = [None] * self.world if self.rank == 0 else None
        self.dist.gather_object(local, parts, dst=0)
        return np.concatenate(parts, axis=0) if self.rank == 0 else None

    def synchronize(self):
        self.solver.synchronize()

    def step_count(self):
        return self.solver.step_count()

    def launch_count(self):
        return self.solver.launch_count()

    def device_view(self):
        return self.solver.device_view()

    def close(self):
        if self.world > 1 and getattr(self, "halo_path", "nccl") == "peer":
            self.solver.synchronize()   # a neighbour may still be pushing into this rank's halo columns
            self.dist.barrier()
        self.solver.close()
