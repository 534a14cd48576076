"""NumPy restatement of the reference D2Q9 MRT-LES solver  --  TEST INFRASTRUCTURE ONLY.

This file is the parity *checker*, not the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product path (``01-lbm-2d_b200``) never does.

It follows ``/root/reference/src/lbm_mrt_les/core/LBM2D_MRT_LES.py`` statement by statement
(citations as ``ref:LINE``), vectorised over the grid but with the reference's per-cell
evaluation order (left-to-right sums, no FMA, IEEE round-to-nearest), in fp32 by default
(the reference's ``ti.f32``) or fp64 (the arbiter).

Typing rules mirrored from Taichi 1.7.4 (default_fp = f32):
  * python-scope constant sub-expressions (``self.tau_0**2``, ``self.Cs_sq_factor``,
    ``2.0/3.0`` ...) fold in float64 and are rounded to f32 where they meet a f32 value;
  * values assigned to kernel locals (``rho_in = self.rho_in_target``) become f32 first;
  * ``int / int`` inside a kernel is a true division of both operands cast to f32;
  * ``tm.cos`` is taken as the correctly rounded f32 of the double cosine (the reference's
    libdevice / libm cosine differs from that by at most an ulp; the ramp is a per-step scalar).

Parity pinning: the reference ships no tests or golden vectors and Taichi
is not installable here, so this restatement is pinned by ``tests/golden/ti_shim_*.npz``:
outputs of the *unmodified reference source* executed under a minimal pure-Python Taichi
stand-in (``tests/golden/gen/``).  See DESIGN.md "Oracle pinning".
"""
from __future__ import annotations

import math

import numpy as np

# D2Q9 direction table, ref:146-158
E = np.array(
    [[0, 0], [1, 0], [0, 1], [-1, 0], [0, -1], [1, 1], [-1, 1], [-1, -1], [1, -1]],
    dtype=np.int64,
)
OPP = (0, 3, 4, 1, 2, 7, 8, 5, 6)  # index of -e_k (the inv_k column of ref:597-607)
# Lallemand-Luo transform, ref:167-180
M_NP = np.array(
    [
        [1, 1, 1, 1, 1, 1, 1, 1, 1],
        [-4, -1, -1, -1, -1, 2, 2, 2, 2],
        [4, -2, -2, -2, -2, 1, 1, 1, 1],
        [0, 1, 0, -1, 0, 1, -1, -1, 1],
        [0, -2, 0, 2, 0, 1, -1, -1, 1],
        [0, 0, 1, 0, -1, 1, 1, -1, -1],
        [0, 0, -2, 0, 2, 1, 1, -1, -1],
        [0, 1, -1, 1, -1, 0, 0, 0, 0],
        [0, 0, 0, 0, 0, 1, -1, 1, -1],
    ],
    dtype=np.float32,
)
# momentum-exchange lookup, ref:597-607: (dir_x, dir_y, inv_k, force_x, force_y)
FORCE_LUT = (
    (0, 0, 0, 0, 0),
    (1, 0, 3, -1, 0),
    (0, 1, 4, 0, -1),
    (-1, 0, 1, 1, 0),
    (0, -1, 2, 0, 1),
    (1, 1, 7, -1, -1),
    (-1, 1, 8, 1, -1),
    (-1, -1, 5, 1, 1),
    (1, -1, 6, -1, 1),
)


def inv_m(exact: bool = False) -> np.ndarray:
    """ref:182 ``np.linalg.inv(M_np).astype(np.float32)``.

    ``exact=True`` returns the correctly rounded fractions M^T / ||row||^2 with true zeros
    (what the CUDA kernels use); the literal version carries ~1e-17 noise where the exact
    entry is zero, which never changes an fp32 rounding in practice (tested).
    """
    if not exact:
        return np.linalg.inv(M_NP).astype(np.float32)
    m64 = M_NP.astype(np.float64)
    norms = (m64 * m64).sum(axis=1)
    return (m64.T / norms[None, :]).astype(np.float32)


def ramp_value(frame_count: int, warmup_steps, F=np.float32):
    """ref:442-443 -- cosine soft start, evaluated once per step."""
    with np.errstate(divide="ignore", invalid="ignore"):
        progress = F(frame_count) / F(warmup_steps)
    progress = F(min(F(1.0), progress)) if not np.isnan(progress) else progress
    arg = F(0.5 * 3.14159265) * progress
    c = F(math.cos(float(arg)))
    return F(F(1.0) - c)


class OracleLBM:
    """Same surface as the reference class (ref:10), state held in numpy arrays."""

    def __init__(self, config, mask_data=None, dtype=np.float32, exact_inv_m=False, slab=None, obstacle_mode="refill"):
        """`obstacle_mode`: "refill" is the reference (ref:452-455).  "bounce_back" is NOT reference behaviour: it
        is the optional half-way bounce-back mode the B200 build offers next to it -- a fluid
        cell whose upstream neighbour i - e_k is solid takes its own post-collision f_opp(k) of the previous
        step instead of pulling, and solid cells are frozen at rest (rho = 1, u = 0) -- restated here so that
        the CUDA path has a checker for that mode too.

        `slab=(x0, nx_owned)`: hold only global columns [x0, x0+nx_owned) plus one halo column on every
        side that is not a domain boundary; `halo_pack` / `halo_unpack` move the populations that
        cross an interface.  A set of slab oracles exchanging halos every step equals the monolithic oracle
        bit for bit (tests/test_slab_cpu.py) -- that is the property the multi-GPU path relies on."""
        self.config = config
        if obstacle_mode not in ("refill", "bounce_back"):
            raise ValueError(f"obstacle_mode {obstacle_mode!r}")
        self.obstacle_mode = obstacle_mode
        self.F = F = np.dtype(dtype).type
        sim = config["simulation"]  # ref:33-44 (strict indexing: KeyError on missing keys)
        self.name = sim["name"]
        self.nx = int(sim["nx"])
        self.ny = int(sim["ny"])
        self.steps_per_frame = sim["compute_step_size"]
        self.warmup_steps = sim["warmup_steps"]
        self.nu = sim["nu"]
        self.tau_0 = 3.0 * self.nu + 0.5
        self.characteristic_length = sim["characteristic_length"]
        self.rho_in_target = sim["rho_in"]
        self.rho_out_target = sim["rho_out"]
        delta_rho = self.rho_in_target - self.rho_out_target  # ref:58-64
        u_char = math.sqrt(2.0 / 3.0 * delta_rho) if delta_rho > 1e-9 else 0.01
        self.Re = (u_char * self.characteristic_length) / self.nu if self.nu > 0 else float("inf")
        self.C_smag = sim["smagorinsky_constant"]  # ref:78-82
        self.Cs_sq_factor = 18.0 * (self.C_smag**2)
        self.S_other = sim["ghost_moments_s"]
        self.viz_sigma = config["outputs"]["gui"]["gaussian_sigma"]
        zones = config["domain_zones"]  # ref:89-94
        self.sponge_w_in = max(1, zones["sponge_in"])
        self.sponge_w_out = max(1, zones["sponge_out"])
        self.sponge_w_top = max(1, zones["sponge_top"])
        self.sponge_w_bot = max(1, zones["sponge_bot"])
        self.sponge_strength = zones["sponge_strength"]

        # slab geometry: local column il <-> global column il + self._lo
        x0, nx_owned = (0, self.nx) if slab is None else (int(slab[0]), int(slab[1]))
        self.west_ring, self.east_ring = x0 == 0, x0 + nx_owned == self.nx
        self._own0 = 0 if self.west_ring else 1
        self._nx_owned = nx_owned
        self._lo = x0 - self._own0
        self.nx_local = nx_owned + (0 if self.west_ring else 1) + (0 if self.east_ring else 1)
        nxg = self.nx
        nx, ny = self.nx_local, self.ny  # ref:97-128
        self.rho = np.zeros((nx, ny), F)
        self.vel = np.zeros((nx, ny, 2), F)
        self.f_old = np.zeros((nx, ny, 9), F)
        self.f_new = np.zeros((nx, ny, 9), F)
        if mask_data is not None:
            self.mask = np.asarray(mask_data).astype(np.float32).reshape(nxg, ny)[self._lo:self._lo + nx].copy()
        else:
            self.mask = np.zeros((nx, ny), np.float32)
        bc = config["boundary_condition"]
        self.bc_type = np.array(bc["type"], dtype=np.int32)
        self.bc_value = np.array(bc["value"], dtype=np.float32).astype(F)
        self.frame_count = 0
        self.force_sum = np.zeros(2, F)
        self.moments_field = np.zeros((nx, ny, 9), F)

        self.w = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4, dtype=np.float32).astype(F)
        self.M = M_NP.astype(F)
        self.invM = inv_m(exact_inv_m).astype(F)
        if F is np.float64 and not exact_inv_m:
            # the arbiter uses the double-precision inverse, not the f32-rounded one
            self.invM = np.linalg.inv(M_NP.astype(np.float64))
            self.w = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4, dtype=np.float64)
        s = F(self.S_other)
        self.S_base = np.array([0, s, s, 0, s, 0, s, 0, 0], dtype=F)  # ref:191-201

        # per-column / per-row sponge damping, ref:364-378
        self.damp_x = np.zeros(nx, F)
        self.damp_y = np.zeros(ny, F)
        st = F(self.sponge_strength)
        for il in range(nx):
            i = il + self._lo  # the sponge uses the GLOBAL x
            if i > (nxg - self.sponge_w_out):
                c = F(i - (nxg - self.sponge_w_out)) / F(self.sponge_w_out)
                self.damp_x[il] = st * (c * c)
            elif i < self.sponge_w_in:
                c = F(self.sponge_w_in - i) / F(self.sponge_w_in)
                self.damp_x[il] = st * (c * c)
        for j in range(ny):
            if j < self.sponge_w_bot:
                c = F(self.sponge_w_bot - j) / F(self.sponge_w_bot)
                self.damp_y[j] = st * (c * c)
            elif j > (ny - self.sponge_w_top):
                c = F(j - (ny - self.sponge_w_top)) / F(self.sponge_w_top)
                self.damp_y[j] = st * (c * c)

    # ------------------------------------------------------------------ helpers
    def _f_eq(self, rho, vel):
        """ref:214-218 on arrays: rho (...), vel (..., 2) -> (..., 9)."""
        F = self.F
        ux, uy = vel[..., 0], vel[..., 1]
        uv = ux * ux + uy * uy
        out = np.empty(rho.shape + (9,), F)
        for k in range(9):
            eu = F(E[k, 0]) * ux + F(E[k, 1]) * uy
            out[..., k] = (self.w[k] * rho) * (F(1) + F(3) * eu + F(4.5) * eu * eu - F(1.5) * uv)
        return out

    # ------------------------------------------------------------------ init
    def init(self):
        """ref:235-241"""
        self.vel[...] = 0
        self.rho[...] = 1
        self.frame_count = 0
        feq = self._f_eq(self.rho, self.vel)
        self.f_old[...] = feq
        self.f_new[...] = feq

    # ------------------------------------------------------------------ step
    def collide_and_stream(self):
        """ref:243-420 (interior cells, solids included)."""
        F = self.F
        nx, ny = self.nx_local, self.ny
        fo = self.f_old
        f = [fo[1 - E[k, 0] : nx - 1 - E[k, 0], 1 - E[k, 1] : ny - 1 - E[k, 1], k] for k in range(9)]
        if self.obstacle_mode == "bounce_back":   # extension, see __init__
            solid = self.mask == 1.0
            own_fluid = ~solid[1 : nx - 1, 1 : ny - 1]
            for k in range(1, 9):
                nb_solid = solid[1 - E[k, 0] : nx - 1 - E[k, 0], 1 - E[k, 1] : ny - 1 - E[k, 1]]
                f[k] = np.where(nb_solid & own_fluid, fo[1 : nx - 1, 1 : ny - 1, OPP[k]], f[k])
        with np.errstate(all="ignore"):
            m = []
            for r in range(9):
                val = np.zeros_like(f[0])
                for c in range(9):
                    val = val + self.M[r, c] * f[c]
                m.append(val)
            rho = m[0]
            pos = rho > 0
            safe = np.where(pos, rho, F(1))
            u = np.where(pos, m[3] / safe, F(0)).astype(F)
            v = np.where(pos, m[5] / safe, F(0)).astype(F)
            u2 = u * u + v * v  # ref:221-233
            meq = [
                rho,
                rho * (F(-2.0) + F(3.0) * u2),
                rho * (F(1.0) - F(3.0) * u2),
                rho * u,
                -rho * u,
                rho * v,
                -rho * v,
                rho * (u * u - v * v),
                rho * u * v,
            ]
            neq7 = m[7] - meq[7]  # ref:334-351
            neq8 = m[8] - meq[8]
            norm = np.sqrt(F(2.0) * neq7 * neq7 + F(2.0) * neq8 * neq8)
            tau0 = F(self.tau_0)
            if self.C_smag > 0.001:
                term = F(self.tau_0**2) + (F(self.Cs_sq_factor) * norm) / rho
                tau_eddy = F(0.5) * (np.sqrt(term) - tau0)
                tau_eff = tau0 + tau_eddy
            else:
                tau_eff = np.full_like(rho, tau0)
            damp = np.maximum(self.damp_x[1 : nx - 1, None], self.damp_y[None, 1 : ny - 1])
            tau_eff = tau_eff + damp  # ref:380
            s_eff = F(1.0) / tau_eff  # ref:398-405
            S = [self.S_base[k] for k in range(7)] + [s_eff, s_eff]
            m_star = [m[k] - S[k] * (m[k] - meq[k]) for k in range(9)]
            for r in range(9):  # ref:413-420
                val = np.zeros_like(f[0])
                for c in range(9):
                    val = val + self.invM[r, c] * m_star[c]
                self.f_new[1 : nx - 1, 1 : ny - 1, r] = val

    def update_macro_var(self):
        """ref:422-436"""
        F = self.F
        nx, ny = self.nx_local, self.ny
        fn = self.f_new[1 : nx - 1, 1 : ny - 1]
        self.f_old[1 : nx - 1, 1 : ny - 1] = fn
        rho = np.zeros(fn.shape[:2], F)
        vx = np.zeros_like(rho)
        vy = np.zeros_like(rho)
        with np.errstate(all="ignore"):
            for k in range(9):
                rho = rho + fn[..., k]
                vx = vx + F(E[k, 0]) * fn[..., k]
                vy = vy + F(E[k, 1]) * fn[..., k]
            pos = rho > 0
            safe = np.where(pos, rho, F(1))
            self.rho[1 : nx - 1, 1 : ny - 1] = rho
            self.vel[1 : nx - 1, 1 : ny - 1, 0] = np.where(pos, vx / safe, F(0))
            self.vel[1 : nx - 1, 1 : ny - 1, 1] = np.where(pos, vy / safe, F(0))

    def _apply_bc_core(self, dr, ibc, jbc, inb, jnb, ramp):
        """ref:457-550 for arrays of ring cells (ibc, jbc) with neighbours (inb, jnb)."""
        F = self.F
        t = int(self.bc_type[dr])
        ibc = np.asarray(ibc)
        jbc = np.asarray(jbc)
        inb = np.asarray(inb)
        jnb = np.asarray(jnb)
        with np.errstate(all="ignore"):
            if t == 0:
                west = (ibc + self._lo) == 0
                if west.any():  # ref:461-486
                    ib, jb, in_, jn = ibc[west], jbc[west], inb[west], jnb[west]
                    rho_in = F(self.rho_in_target)
                    rho_c = F(1.0) + (rho_in - F(1.0)) * ramp
                    fn = self.f_old[in_, jn]
                    f0, f2, f3, f4, f6, f7 = (fn[:, k] for k in (0, 2, 3, 4, 6, 7))
                    ux = F(1.0) - (f0 + f2 + f4 + F(2.0) * (f3 + f6 + f7)) / rho_c
                    f1 = f3 + F(2.0 / 3.0) * rho_c * ux
                    f5 = f7 - F(0.5) * (f2 - f4) + F(1.0 / 6.0) * rho_c * ux
                    f8 = f6 + F(0.5) * (f2 - f4) + F(1.0 / 6.0) * rho_c * ux
                    self.rho[ib, jb] = rho_c
                    self.vel[ib, jb, 0] = ux
                    self.vel[ib, jb, 1] = F(0)
                    feq = self._f_eq(self.rho[ib, jb], self.vel[ib, jb])
                    feq[:, 1] = f1
                    feq[:, 5] = f5
                    feq[:, 8] = f8
                    self.f_old[ib, jb] = feq
                rest = ~west
                if rest.any():  # ref:487-492
                    ib, jb, in_, jn = ibc[rest], jbc[rest], inb[rest], jnb[rest]
                    self.vel[ib, jb, 0] = self.bc_value[dr, 0] * ramp
                    self.vel[ib, jb, 1] = self.bc_value[dr, 1] * ramp
                    self.rho[ib, jb] = self.rho[in_, jn]
                    self.f_old[ib, jb] = (
                        self._f_eq(self.rho[ib, jb], self.vel[ib, jb])
                        - self._f_eq(self.rho[in_, jn], self.vel[in_, jn])
                        + self.f_old[in_, jn]
                    )
            elif t == 1:
                east = (ibc + self._lo) == self.nx - 1
                if east.any():  # ref:495-527
                    ib, jb, in_, jn = ibc[east], jbc[east], inb[east], jnb[east]
                    rho_out = F(self.rho_out_target)
                    fn = self.f_old[in_, jn]
                    f0, f1, f2, f4, f5, f8 = (fn[:, k] for k in (0, 1, 2, 4, 5, 8))
                    ux = F(-1.0) + (f0 + f2 + f4 + F(2.0) * (f1 + f5 + f8)) / rho_out
                    back = ux < 0.0
                    # backflow guard, ref:508-516
                    b = np.nonzero(back)[0]
                    if b.size:
                        self.vel[ib[b], jb[b]] = self.vel[in_[b], jn[b]]
                        self.rho[ib[b], jb[b]] = rho_out
                        self.f_old[ib[b], jb[b]] = (
                            self._f_eq(self.rho[ib[b], jb[b]], self.vel[ib[b], jb[b]])
                            - self._f_eq(self.rho[in_[b], jn[b]], self.vel[in_[b], jn[b]])
                            + self.f_old[in_[b], jn[b]]
                        )
                    g = np.nonzero(~back)[0]  # ref:517-527 (NaN ux lands here, as in the reference)
                    if g.size:
                        uxg = ux[g]
                        f3 = f1[g] - F(2.0 / 3.0) * rho_out * uxg
                        f6 = f8[g] - F(0.5) * (f2[g] - f4[g]) - F(1.0 / 6.0) * rho_out * uxg
                        f7 = f5[g] + F(0.5) * (f2[g] - f4[g]) - F(1.0 / 6.0) * rho_out * uxg
                        self.rho[ib[g], jb[g]] = rho_out
                        self.vel[ib[g], jb[g], 0] = uxg
                        self.vel[ib[g], jb[g], 1] = F(0)
                        feq = self._f_eq(self.rho[ib[g], jb[g]], self.vel[ib[g], jb[g]])
                        feq[:, 3] = f3
                        feq[:, 6] = f6
                        feq[:, 7] = f7
                        self.f_old[ib[g], jb[g]] = feq
                # type 1 on any other cell: no-op
            elif t == 2:  # ref:529-550
                vert = ibc == inb  # top / bottom wall
                self.vel[ibc, jbc, 0] = np.where(vert, self.vel[inb, jnb, 0], F(0))
                self.vel[ibc, jbc, 1] = np.where(vert, F(0), self.vel[inb, jnb, 1])
                self.rho[ibc, jbc] = self.rho[inb, jnb]
                self.f_old[ibc, jbc] = (
                    self._f_eq(self.rho[ibc, jbc], self.vel[ibc, jbc])
                    - self._f_eq(self.rho[inb, jnb], self.vel[inb, jnb])
                    + self.f_old[inb, jnb]
                )
            # any other type: no-op

    def apply_bc(self):
        """ref:438-455"""
        F = self.F
        nx, ny = self.nx_local, self.ny
        self.frame_count += 1
        ramp = ramp_value(self.frame_count, self.warmup_steps, F)
        j = np.arange(1, ny - 1)
        z = np.zeros_like(j)
        if self.west_ring:
            self._apply_bc_core(0, z, j, z + 1, j, ramp)
        if self.east_ring:
            self._apply_bc_core(2, z + (nx - 1), j, z + (nx - 2), j, ramp)
        i = np.arange(self._own0, self._own0 + self._nx_owned)  # owned columns (all of them when not a slab)
        z = np.zeros_like(i)
        self._apply_bc_core(1, i, z + (ny - 1), i, z + (ny - 2), ramp)
        self._apply_bc_core(3, i, z, i, z + 1, ramp)
        solid = self.mask == 1.0
        solid[: self._own0] = False                       # halo columns belong to the neighbour
        solid[self._own0 + self._nx_owned:] = False
        self.vel[solid] = 0
        if self.obstacle_mode == "bounce_back":   # frozen at rest
            self.rho[solid] = 1
        self.f_old[solid] = self._f_eq(self.rho[solid], self.vel[solid])

    # ------------------------------------------------------------------ slab halos
    EAST_GOING = (1, 5, 8)   # populations with e_x = +1: pulled from column i-1
    WEST_GOING = (3, 6, 7)

    def halo_pack(self, side):
        """Populations of the first ('W') / last ('E') owned column that stream into the neighbour."""
        if side == "E":
            return self.f_old[self._own0 + self._nx_owned - 1][:, list(self.EAST_GOING)].copy()
        return self.f_old[self._own0][:, list(self.WEST_GOING)].copy()

    def halo_unpack(self, side, data):
        """Fill the 'W' / 'E' halo column with what the neighbour packed for us.  The halo's f_new (read
        by the force of an owned solid next to it) equals f_old at interior rows, as on the owning rank."""
        col = self.nx_local - 1 if side == "E" else 0
        planes = list(self.WEST_GOING if side == "E" else self.EAST_GOING)
        self.f_old[col][:, planes] = data
        fn = self.f_new[col]
        fn[1:-1, planes] = data[1:-1]

    def run_step(self, steps=1):
        """ref:552-573"""
        for _ in range(steps):
            self.collide_and_stream()
            self.update_macro_var()
            self.apply_bc()

    # ------------------------------------------------------------------ diagnostics
    def get_force(self):
        """ref:588-646; terms added sequentially in (i, j, k) order."""
        F = self.F
        nx, ny = self.nx_local, self.ny
        terms_x, terms_y = [], []
        solid_idx = np.argwhere(self.mask == 1)
        for i, j in solid_idx:
            if not (self._own0 <= i < self._own0 + self._nx_owned):
                continue
            for k in range(9):
                dx, dy, inv_k, fx, fy = FORCE_LUT[k]
                ni, nj = i + dx, j + dy
                if 0 <= ni + self._lo < self.nx and 0 <= nj < ny and self.mask[ni, nj] == 0:
                    fv = F(2.0) * self.f_new[ni, nj, inv_k]
                    terms_x.append(fv * F(fx))
                    terms_y.append(fv * F(fy))
        out = np.zeros(2, F)
        with np.errstate(all="ignore"):
            if terms_x:
                out[0] = np.cumsum(np.array(terms_x, F), dtype=F)[-1]
                out[1] = np.cumsum(np.array(terms_y, F), dtype=F)[-1]
        self.force_sum[:] = out
        return out.copy()

    def get_max_velocity(self):
        """ref:648-660 -- max ||vel|| over all cells, starting from 0."""
        with np.errstate(all="ignore"):
            mag = np.sqrt(self.vel[..., 0] * self.vel[..., 0] + self.vel[..., 1] * self.vel[..., 1])
        if np.isnan(mag).any():
            return float("nan")
        return float(max(self.F(0), mag.max()))

    def compute_moments_for_output(self):
        """ref:667-737 (hand-expanded rows, from f_new at ALL cells)."""
        F = self.F
        f = [self.f_new[..., k] for k in range(9)]
        with np.errstate(all="ignore"):
            rho = f[0]
            for k in range(1, 9):
                rho = rho + f[k]
            s14 = f[1] + f[2] + f[3] + f[4]
            s58 = f[5] + f[6] + f[7] + f[8]
            e = F(-4.0) * f[0] - s14 + F(2.0) * s58
            eps = F(4.0) * f[0] - F(2.0) * s14 + s58
            jx = f[1] - f[3] + f[5] - f[6] - f[7] + f[8]
            qx = F(-2.0) * f[1] + F(2.0) * f[3] + f[5] - f[6] - f[7] + f[8]
            jy = f[2] - f[4] + f[5] + f[6] - f[7] - f[8]
            qy = F(-2.0) * f[2] + F(2.0) * f[4] + f[5] + f[6] - f[7] - f[8]
            pxx = f[1] - f[2] + f[3] - f[4]
            pxy = f[5] - f[6] + f[7] - f[8]
        for k, a in enumerate((rho, e, eps, jx, qx, jy, qy, pxx, pxy)):
            self.moments_field[..., k] = a

    def get_moments_numpy(self):
        self.compute_moments_for_output()
        return self.moments_field.copy()

    def get_physical_fields(self):
        return self.vel.copy(), self.mask.copy()
