// Kernels of the fused D2Q9 MRT-LES step (sm_100a).  See DESIGN.md for the data layout.
//
// Layout in HBM: 9 SoA planes per buffer, each (nx_local, pitch) with y fastest and
// pitch = round_up(ny, 32) floats, so every column starts on a 128-byte line and a warp's
// vector accesses are full, aligned lines.  Two buffers (src/dst) swap roles every step.
// Variants: step_kernel (register, default; this file), step_tma_kernel (lbm2d_tma.cuh),
// step_async_kernel (lbm2d_async.cuh).
#pragma once
#include "lbm2d_device.cuh"

namespace lbm {

// Tuning knobs (measured on B200, 8192x2048: profiles/r01_tuning_sweep.md).  Small CTAs (64-128 threads) win:
// the warps of a CTA move through load / math / store in lock-step, so many small CTAs per SM keep the
// memory pipeline evenly fed.
#ifndef LBM_WPB
#define LBM_WPB 4
#endif
#ifndef LBM_MINB2
#define LBM_MINB2 10
#endif
#ifndef LBM_MINB1
#define LBM_MINB1 12
#endif
#ifndef LBM_STCS
#define LBM_STCS 0
#endif
constexpr int kWarpsPerBlock = LBM_WPB;
static_assert(LBM_WPB >= 2, "the top / bottom ring row of a group needs two warps");
constexpr int kRingGroup = 32;   // interior columns per top/bottom ring row of the grid (one lane per column)
constexpr int kThreads = kWarpsPerBlock * 32;

struct StepArgs {
    const float *__restrict__ src;  // 9 planes
    float *__restrict__ dst;        // 9 planes
    const uint8_t *__restrict__ code;  // cell code, bit0 = solid
    const uint8_t *__restrict__ links8;  // bounce-back mode only: bit k-1 set = the upstream neighbour i - e_k of this FLUID cell is solid
    const uint32_t *__restrict__ code_bits;  // the same bit, 32 cells per word (plane order): what the interior warps read
    const float *__restrict__ damp_x;  // [nx_local]   ref:364-370 (indexed by local column, holds the global value)
    const float *__restrict__ damp_y;  // [pitch]      ref:372-378
    const float *__restrict__ ramp_tab;  // [warmup+1]  ref:442-443
    const int *ctr_in;              // frame_count before this step
    int *ctr_out;                   // frame_count after this step (other parity slot)
    float *rho, *ux, *uy;           // macroscopic planes (written by EMIT steps)
    unsigned *maxv_bits;            // max(ux^2+uy^2) as ordered uint; [1] = NaN flag
    long long plane;                // floats per plane = nx_local * pitch
    int nx_local, ny, pitch, nseg;
    int x_off;                      // global x of local column 0
    int west_ring, east_ring;       // local column 0 / nx_local-1 is the domain boundary (else a halo)
    int warmup;
    int il0, il_step, il_count;     // columns of this launch: il0 + blockIdx.y * il_step, blockIdx.y < il_count
    int bump_ctr;                   // this launch advances frame_count (exactly one launch per step does)
    int n_ring;                     // ring cells handled by this launch's ring warps
    int ring_row0, ring_rows;       // grid rows [ring_row0, ring_row0 + ring_rows): W/E ring warps; the others: see step_kernel
    // Early start (see step_kernel): rows [0, early_rows) may begin on the progress counter instead of the full
    // completion of the previous step; rows [0, low_rows) of every step add 1 per CTA to it when done.
    int early_rows, low_rows;
    unsigned long long *progress;
    unsigned long long progress_expected;   // counter value once the previous step's rows [0, low_rows) are complete
    const RingCtx *ring;            // rare-path context in global memory (dst-specific)
    Physics phys;
};

__device__ __forceinline__ float vmag2_strict(float ux, float uy) {
    return __fadd_rn(__fmul_rn(ux, ux), __fmul_rn(uy, uy));
}

// aligned V-wide global accesses (V = 1, 2, 4 floats)
// Population loads bypass L1 (ld.global.cg): every value is read exactly once per step; measured 1 % faster.
#ifndef LBM_LDCS
#define LBM_LDCS 2
#endif
#if LBM_LDCS == 2
#define LBM_LD(ptr) __ldcg(ptr)
#elif LBM_LDCS
#define LBM_LD(ptr) __ldcs(ptr)
#else
#define LBM_LD(ptr) __ldg(ptr)
#endif
template <int V>
__device__ __forceinline__ void ldf(const float *p, float (&o)[V]) {   // populations: read exactly once per step
    if (V == 4) { const float4 t = LBM_LD(reinterpret_cast<const float4 *>(p)); o[0] = t.x; o[1 % V] = t.y; o[2 % V] = t.z; o[3 % V] = t.w; }
    else if (V == 2) { const float2 t = LBM_LD(reinterpret_cast<const float2 *>(p)); o[0] = t.x; o[1 % V] = t.y; }
    else o[0] = LBM_LD(p);
}
template <int V>
__device__ __forceinline__ void ldv(const float *p, float (&o)[V]) {
    if (V == 4) { const float4 t = __ldg(reinterpret_cast<const float4 *>(p)); o[0] = t.x; o[1 % V] = t.y; o[2 % V] = t.z; o[3 % V] = t.w; }
    else if (V == 2) { const float2 t = __ldg(reinterpret_cast<const float2 *>(p)); o[0] = t.x; o[1 % V] = t.y; }
    else o[0] = __ldg(p);
}
template <int V>
__device__ __forceinline__ void stv(float *p, const float (&o)[V]) {
#if LBM_STCS
    if (V == 4) __stcs(reinterpret_cast<float4 *>(p), make_float4(o[0], o[1 % V], o[2 % V], o[3 % V]));
    else if (V == 2) __stcs(reinterpret_cast<float2 *>(p), make_float2(o[0], o[1 % V]));
    else __stcs(p, o[0]);
#else
    if (V == 4) *reinterpret_cast<float4 *>(p) = make_float4(o[0], o[1 % V], o[2 % V], o[3 % V]);
    else if (V == 2) *reinterpret_cast<float2 *>(p) = make_float2(o[0], o[1 % V]);
    else *p = o[0];
#endif
}
template <int V>
__device__ __forceinline__ void ldcode(const uint8_t *p, unsigned char (&o)[V]) {
    if (V == 4) { const uchar4 t = __ldg(reinterpret_cast<const uchar4 *>(p)); o[0] = t.x; o[1 % V] = t.y; o[2 % V] = t.z; o[3 % V] = t.w; }
    else if (V == 2) { const uchar2 t = __ldg(reinterpret_cast<const uchar2 *>(p)); o[0] = t.x; o[1 % V] = t.y; }
    else o[0] = __ldg(p);
}

// Half-way bounce-back (optional obstacle mode, not the reference's): a population whose upstream neighbour is
// solid is replaced by the cell's own post-collision population of the opposite direction from the
// previous step, f_k(x, t+1) = f*_opp(k)(x, t); the source buffer holds exactly those values.
__device__ constexpr int kOpp[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
__device__ __forceinline__ void bounce_back(const StepArgs &a, unsigned links, long long o, float (&fin)[9]) {
#pragma unroll
    for (int k = 1; k < 9; ++k)
        if ((links >> (k - 1)) & 1u) fin[k] = LBM_LD(a.src + kOpp[k] * a.plane + o);
}

// ---------------------------------------------------------------------------------------------
// Ring warps.  Every boundary-ring cell is a function of ONE adjacent interior cell's fresh, un-refilled
// state (corners chain through the W/E cell).  Instead of making the interior thread that
// owns that neighbour produce it (a serial, divergent detour for one lane of a streaming warp), extra
// warps of the SAME launch take one ring cell per lane: they re-derive the owner's collision from the
// source buffer (a handful of scalar loads; the ring is O(perimeter)) and run the reference's
// apply_bc_core on it, 32 ring cells in parallel.  They depend on nothing the interior warps write.
//
// Enumeration of the ring cells of a launch covering columns {il0 + c * il_step, c < il_count}:
//   [0, n)        top row    (il_c, ny-1)   <- owner (il_c, ny-2)      dr = 1      ref:449
//   [n, 2n)       bottom row (il_c, 0)      <- owner (il_c, 1)         dr = 3      ref:450
//   then, if the launch holds column 1 and that side is a domain boundary:
//   W column (0, j), j = 1..ny-2 <- owner (1, j), dr = 0 (ref:446); corners (0, ny-1), (0, 0) <- W cell <- owner
//   and likewise E column / corners for column nx_local-2 (dr = 2, ref:447).
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline int ring_cell_count(int il0, int il_step, int il_count, int nx_local, int ny, int west_ring, int east_ring) {
    const bool w = west_ring && il0 == 1;
    const bool e = east_ring && (il0 + (il_count - 1) * il_step == nx_local - 2);
    return 2 * il_count + (w ? ny : 0) + (e ? ny : 0);   // (ny - 2) column cells + 2 corners per side
}

template <bool STRICT, bool EMIT, bool BB>
__device__ __forceinline__ void ring_cell(const StepArgs &a, int idx, int ramp_fc, float &vmax, int &vnan) {
    const int ny = a.ny, pitch = a.pitch, n = a.il_count;
    const long long plane = a.plane;
    // decode: ring cell (ilr, jr), its owner (ilo, jo), boundary side dr; corners chain W/E -> top/bottom
    int ilr, jr, ilo, jo, dr, corner_dr = -1;
    if (idx < 2 * n) {
        const bool top = idx < n;
        ilo = ilr = a.il0 + (top ? idx : idx - n) * a.il_step;
        jo = top ? ny - 2 : 1;
        jr = top ? ny - 1 : 0;
        dr = top ? 1 : 3;
    } else {
        int q = idx - 2 * n;
        const bool has_w = a.west_ring && a.il0 == 1;
        const bool west = has_w && q < ny;
        if (!west && has_w) q -= ny;
        ilo = west ? 1 : a.nx_local - 2;
        ilr = west ? 0 : a.nx_local - 1;
        dr = west ? 0 : 2;
        if (q < ny - 2) { jo = jr = q + 1; }
        else if (q == ny - 2) { jo = ny - 2; jr = ny - 1; corner_dr = 1; }   // top corner through (ilr, ny-2)
        else { jo = 1; jr = 0; corner_dr = 3; }                               // bottom corner through (ilr, 1)
    }
    // owner's pull + collision + macroscopic values (same arithmetic as the interior warps)
    float fin[9], g[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) fin[k] = LBM_LD(a.src + k * plane + (long long)(ilo - kEx[k]) * pitch + (jo - kEy[k]));
    if (BB) {
        const long long oo = (long long)ilo * pitch + jo;
        const unsigned links = __ldg(a.links8 + oo);
        if (links) bounce_back(a, links, oo, fin);
    }
    const float damp = fmaxf(__ldg(a.damp_x + ilo), __ldg(a.damp_y + jo));
    if (STRICT) collide_strict(a.phys, fin, damp, g);
    else collide_fast(a.phys, fin, damp, g);
    Cell me, r;
#pragma unroll
    for (int k = 0; k < 9; ++k) me.f[k] = g[k];
    macro_from_f<STRICT>(g, me.rho, me.ux, me.uy);
    const float ramp = __ldg(a.ramp_tab + min(ramp_fc, a.warmup));
    const int igo = a.x_off + ilo, igr = a.x_off + ilr;
    cell_rest(r);
    bc_core(a.phys, dr, igr, igo, me, r, ramp);
    if (corner_dr >= 0) {   // ref:448-450 run over i = 0 and nx-1 too: the corner reads the W/E cell just produced
        Cell cr;
        cell_rest(cr);
        bc_core(a.phys, corner_dr, igr, igr, r, cr, ramp);
        r = cr;
    }
    const long long o = (long long)ilr * pitch + jr;
    if (__ldg(a.code + o) & 1) {   // ref:452-455 also resets solid ring cells
        if (BB) cell_rest(r);      // bounce-back mode: solids are frozen at rest
        else refill(r);
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) a.dst[k * plane + o] = r.f[k];
    if (EMIT) {
        a.rho[o] = r.rho;
        a.ux[o] = r.ux;
        a.uy[o] = r.uy;
        const float m2 = __fadd_rn(__fmul_rn(r.ux, r.ux), __fmul_rn(r.uy, r.uy));
        vnan |= (m2 != m2);
        vmax = fmaxf(vmax, m2);
    }
}

// One fused pass: pull-stream, MRT-LES collision, sponge, macroscopic update, boundary ring,
// obstacle refill (ref:552-573 = collide_and_stream + update_macro_var + apply_bc), f_src -> f_dst.
//
// "Register" variant.  Interior warps: one warp = one (32 V)-cell segment of one interior column, one
// thread = V consecutive cells in y; every access is aligned and fully coalesced, the +-1 shift of the
// pull in y comes from the neighbouring lane by warp shuffle with one extra scalar load at each end of
// the segment, all issued before first use; no boundary code at all.  Ring warps (their own grid rows:
// one row behind every 32 columns for the top / bottom cells, one block of rows for the W / E columns):
// one ring cell per lane, see above.  BB: optional half-way bounce-back obstacle mode (not the reference's).
template <bool STRICT, bool EMIT, int V, bool BB = false>
__global__ void __launch_bounds__(kThreads, (V == 4 ? 10 : (V == 2 ? LBM_MINB2 : LBM_MINB1))) step_kernel(const StepArgs a) {
    // grid: x = blocks of segments down a column, y (+ z beyond 65535) = rows (columns and ring rows, see below)
    const int row = blockIdx.y + blockIdx.z * 65535;
    // Programmatic dependent launch: this grid is scheduled while the previous step drains, and waits here
    // until that grid's writes are complete and visible (a no-op for ordinary launches).  Early start: the
    // CTAs of the first columns -- the ones that get the SM slots freed during the previous step's tail --
    // need only the previous step's first columns and ring, which finished ~190 us ago; they check a
    // progress counter (one acquire load, no polling) and fall back to the full wait if it is not there yet.
    if (row < a.early_rows) {
        unsigned long long seen;
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(a.progress) : "memory");
        if (seen < a.progress_expected) asm volatile("griddepcontrol.wait;" ::: "memory");
    } else {
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    asm volatile("griddepcontrol.launch_dependents;");
    const int lane = threadIdx.x & 31;
    const int seg = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    // Row numbering: the W/E ring block sits at [ring_row0, ring_row0 + ring_rows); the other rows count
    // groups of 33 -- 32 interior columns followed by ONE row for the top and bottom ring cells of those columns.
    // Those cells are 4-byte accesses at a stride of one column; done right behind their columns they hit the
    // lines the interior warps are reading (source) and merge in L2 with the sectors they are writing
    // (destination), instead of costing a DRAM read-modify-write each (7.6 us of a 193 us step otherwise).
    const int ring_rel = row - a.ring_row0;
    const bool we_row = ring_rel >= 0 && ring_rel < a.ring_rows;
    const int vrow = ring_rel < 0 ? row : row - a.ring_rows;
    const int grp = vrow / (kRingGroup + 1), grp_r = vrow - grp * (kRingGroup + 1);
    const bool tb_row = grp_r == kRingGroup;
    const int col = grp * kRingGroup + grp_r;
    if (a.bump_ctr && blockIdx.x == 0 && row == 0 && threadIdx.x == 0) *a.ctr_out = __ldcg(a.ctr_in) + 1;  // ref:440
    float vmax = 0.0f;  // max |u|^2 over the cells written by this thread (EMIT only)
    int vnan = 0;
    if (we_row) {
        // ------------------------------- ring warps: W / E columns and corners -----------------
        const int idx = 2 * a.il_count + (ring_rel * (int)gridDim.x * kWarpsPerBlock + seg) * 32 + lane;
        if (idx < a.n_ring) ring_cell<STRICT, EMIT, BB>(a, idx, __ldcg(a.ctr_in) + 1, vmax, vnan);
    } else if (tb_row) {
        // ------------------------------- ring warps: top (warp 0) / bottom (warp 1) of one group
        const int c = grp * kRingGroup + lane;
        if (blockIdx.x == 0 && threadIdx.x < 64 && c < a.il_count)
            ring_cell<STRICT, EMIT, BB>(a, (threadIdx.x >> 5) * a.il_count + c, __ldcg(a.ctr_in) + 1, vmax, vnan);
    } else if (col < a.il_count && seg < a.nseg) {
        // ------------------------------- interior warps --------------------------------------
        const int il = a.il0 + col * a.il_step;                              // local column
        const int j0 = seg * (32 * V) + lane * V;
        const bool lane_on = j0 < a.pitch;
        const int ny = a.ny, pitch = a.pitch;
        const long long plane = a.plane;

        // pull (ref:254-257): fin[c][k] = f_k(i - e_kx, j0 + c - e_ky); every load issued before the first use
        float v[9][V];
        float edge[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const float *colp = a.src + k * plane + (long long)(il - kEx[k]) * pitch;
#pragma unroll
            for (int c = 0; c < V; ++c) v[k][c] = 0.f;
            edge[k] = 0.f;
            if (lane_on) ldf<V>(colp + j0, v[k]);
            if (kEy[k] == 1 && lane == 0 && lane_on && j0 > 0) edge[k] = LBM_LD(colp + j0 - 1);
            if (kEy[k] == -1 && lane == 31 && j0 + V < ny) edge[k] = LBM_LD(colp + j0 + V);
        }
        const bool live = lane_on && j0 < ny;  // padding lanes only feed the shuffles / the EMIT reduction
        float dx = 0.f;
        float dy[V];
        unsigned char code[V], links[V];
#pragma unroll
        for (int c = 0; c < V; ++c) { dy[c] = 0.f; code[c] = 0; links[c] = 0; }
        if (live) {
            dx = __ldg(a.damp_x + il);
            ldv<V>(a.damp_y + j0, dy);
            // solid bits of the warp's 32 V cells = V consecutive words; a lane's V cells sit in one of them
            const uint32_t w = __ldg(a.code_bits + (((long long)il * pitch + j0) >> 5));
#pragma unroll
            for (int c = 0; c < V; ++c) code[c] = (w >> ((j0 & 31) + c)) & 1u;
            if (BB) ldcode<V>(a.links8 + (long long)il * pitch + j0, links);
        }
        float fin[V][9];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            if (kEy[k] == 0) {
#pragma unroll
                for (int c = 0; c < V; ++c) fin[c][k] = v[k][c];
            } else if (kEy[k] == 1) {  // needs j-1: last element of the lane below
                float below = __shfl_up_sync(0xffffffffu, v[k][V - 1], 1);
                if (lane == 0) below = edge[k];
                fin[0][k] = below;
#pragma unroll
                for (int c = 1; c < V; ++c) fin[c][k] = v[k][c - 1];
            } else {                   // needs j+1: first element of the lane above
                float above = __shfl_down_sync(0xffffffffu, v[k][0], 1);
                if (lane == 31) above = edge[k];
#pragma unroll
                for (int c = 0; c < V - 1; ++c) fin[c][k] = v[k][c + 1];
                fin[V - 1][k] = above;
            }
        }
        if (live) {
            // collide (ref:266-420); rho / u (ref:425-436) only where consumed: EMIT steps and the obstacle refill
            float g[V][9], rho[V], ux[V], uy[V];
            bool interior[V], all_interior = true, any_interior = false;
#pragma unroll
            for (int c = 0; c < V; ++c) {
                const int j = j0 + c;
                interior[c] = (j >= 1) && (j <= ny - 2);
                all_interior &= interior[c];
                any_interior |= interior[c];
                const float damp = fmaxf(dx, dy[c]);
                if (BB && links[c]) bounce_back(a, links[c], (long long)il * pitch + j0 + c, fin[c]);
#ifdef LBM_NOMATH   // experiment: pure streaming bound of this access pattern
#pragma unroll
                for (int k = 0; k < 9; ++k) g[c][k] = fin[c][k] + damp;
#else
                if (STRICT) collide_strict(a.phys, fin[c], damp, g[c]);
                else collide_fast(a.phys, fin[c], damp, g[c]);
#endif
                rho[c] = ux[c] = uy[c] = 0.0f;
                if (EMIT || (code[c] & 1)) macro_from_f<STRICT>(g[c], rho[c], ux[c], uy[c]);
                if (code[c] & 1) {  // obstacle refill, ref:452-455 (bounce-back mode: frozen at rest)
                    ux[c] = 0.0f; uy[c] = 0.0f;
                    if (BB) rho[c] = 1.0f;
#pragma unroll
                    for (int k = 0; k < 9; ++k) g[c][k] = __fmul_rn(kW[k], rho[c]);
                }
            }
            const long long o = (long long)il * pitch + j0;
            if (all_interior) {            // the common case: one wide store per plane
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    float t[V];
#pragma unroll
                    for (int c = 0; c < V; ++c) t[c] = g[c][k];
                    stv<V>(a.dst + k * plane + o, t);
                }
                if (EMIT) {
                    stv<V>(a.rho + o, rho);
                    stv<V>(a.ux + o, ux);
                    stv<V>(a.uy + o, uy);
                }
            } else if (any_interior) {     // the vector shares a ring cell (ring warps write it) or padding: cell by cell
#pragma unroll
                for (int c = 0; c < V; ++c) {
                    if (!interior[c]) continue;
#pragma unroll
                    for (int k = 0; k < 9; ++k) a.dst[k * plane + o + c] = g[c][k];
                    if (EMIT) { a.rho[o + c] = rho[c]; a.ux[o + c] = ux[c]; a.uy[o + c] = uy[c]; }
                }
            }
            if (EMIT) {
#pragma unroll
                for (int c = 0; c < V; ++c) {
                    if (!interior[c]) continue;
                    const float m2 = vmag2_strict(ux[c], uy[c]);
                    vnan |= (m2 != m2);
                    vmax = fmaxf(vmax, m2);
                }
            }
        }
    }

    if (EMIT) {  // whole warp: max over lanes, one atomic per warp and only if it raises the running max
        for (int s = 16; s > 0; s >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, s));
        const bool any_nan = __any_sync(0xffffffffu, vnan != 0);
        if (lane == 0) {
            const unsigned bits = __float_as_uint(vmax);
            if (bits > *a.maxv_bits) atomicMax(a.maxv_bits, bits);
            if (any_nan) a.maxv_bits[1] = 1u;
        }
    }
    if (row < a.low_rows) {   // release: this CTA's part of the low rows is complete and visible device-wide
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(a.progress, 1ULL);
        }
    }
}

// init(), ref:235-241: both buffers = w_k, rho = 1, u = 0; padding cells = 0.
__global__ void init_kernel(float *f0, float *f1, float *rho, float *ux, float *uy, long long plane, int ny, int pitch) {
    const long long n = plane;
    for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < n; o += (long long)gridDim.x * blockDim.x) {
        const bool real = (int)(o % pitch) < ny;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const float v = real ? kW[k] : 0.0f;
            f0[k * plane + o] = v;
            f1[k * plane + o] = v;
        }
        rho[o] = real ? 1.0f : 0.0f;
        ux[o] = 0.0f;
        uy[o] = 0.0f;
    }
}

// ---- export / parity kernels (not on the per-step path) ---------------------------------------
struct ExportArgs {
    const float *cur;   // state after the last step (= the reference's f_old)
    const float *prev;  // state before the last step (still intact in the other buffer)
    const uint8_t *code;
    const float *damp_x, *damp_y;
    long long plane;
    int nx_local, ny, pitch;
    int il0, il1;       // local columns to export [il0, il1)
    int x_off, nx_global;
    int have_prev;      // 0 right after init(): f_new == f_old == w everywhere
    int strict;
    Physics phys;
};

// The reference's f_new (ref:104): post-collision values at interior cells -- equal to `cur` on fluid
// cells, re-derived from `prev` on solid cells (cur holds their refill there) -- and the initial
// equilibrium on the boundary ring, which the reference never updates in f_new.
__device__ __forceinline__ void load_f_new(const ExportArgs &a, int il, int j, float (&f)[9]) {
    const int ig = a.x_off + il;
    const bool ring = (ig == 0) || (ig == a.nx_global - 1) || (j == 0) || (j == a.ny - 1);
    const long long o = (long long)il * a.pitch + j;
    if (ring) {
#pragma unroll
        for (int k = 0; k < 9; ++k) f[k] = kW[k];
    } else if (!a.have_prev || !(a.code[o] & 1)) {
#pragma unroll
        for (int k = 0; k < 9; ++k) f[k] = a.cur[k * a.plane + o];
    } else {
        float fin[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) fin[k] = a.prev[k * a.plane + (long long)(il - kEx[k]) * a.pitch + (j - kEy[k])];
        const float damp = fmaxf(a.damp_x[il], a.damp_y[j]);
        if (a.strict) collide_strict(a.phys, fin, damp, f);
        else collide_fast(a.phys, fin, damp, f);
    }
}

// mode 0: moments of f_new (ref:667-737);  1: f_old as AoS;  2: f_new as AoS.   out: (ncols, ny, 9)
__global__ void export9_kernel(const ExportArgs a, int mode, float *__restrict__ out) {
    const int j = blockIdx.y * blockDim.x + threadIdx.x;   // grid: x = column (no 65 535 limit), y = blocks down the column
    const int il = a.il0 + blockIdx.x;
    if (j >= a.ny) return;
    float f[9], o9[9];
    if (mode == 1) {
        const long long o = (long long)il * a.pitch + j;
#pragma unroll
        for (int k = 0; k < 9; ++k) o9[k] = a.cur[k * a.plane + o];
    } else {
        load_f_new(a, il, j, f);
        if (mode == 0) moments_strict(f, o9);
        else {
#pragma unroll
            for (int k = 0; k < 9; ++k) o9[k] = f[k];
        }
    }
    float *dst = out + ((long long)(il - a.il0) * a.ny + j) * 9;
#pragma unroll
    for (int k = 0; k < 9; ++k) dst[k] = o9[k];
}

// (ncols, ny, nch) interleaved from up to 2 pitched planes (vel: ux,uy; rho: one plane).
__global__ void pack_planes_kernel(const float *p0, const float *p1, int nch, int il0, int ny, int pitch, float *__restrict__ out) {
    const int j = blockIdx.y * blockDim.x + threadIdx.x;
    const int il = il0 + blockIdx.x;
    if (j >= ny) return;
    const long long o = (long long)il * pitch + j;
    float *dst = out + ((long long)blockIdx.x * ny + j) * nch;
    dst[0] = p0[o];
    if (nch == 2) dst[1] = p1[o];
}

__global__ void mask_to_float_kernel(const uint8_t *code, int il0, int ny, int pitch, float *__restrict__ out) {
    const int j = blockIdx.y * blockDim.x + threadIdx.x;
    const int il = il0 + blockIdx.x;
    if (j >= ny) return;
    out[(long long)blockIdx.x * ny + j] = (code[(long long)il * pitch + j] & 1) ? 1.0f : 0.0f;
}

// Momentum exchange over the precomputed solid-fluid links, ref:588-641.
// link = (offset of the fluid neighbour in a plane) , packed (inv_k | ring<<4 | (fx+1)<<5 | (fy+1)<<7)
struct Link {
    int offset;
    int packed;
};
__global__ void force_kernel(const float *cur, long long plane, const Link *links, int n_links, double *partial /*[grid][2]*/) {
    double fx = 0.0, fy = 0.0;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_links; t += gridDim.x * blockDim.x) {
        const Link l = links[t];
        const int inv_k = l.packed & 15;
        const bool ring = (l.packed >> 4) & 1;
        const int sx = ((l.packed >> 5) & 3) - 1, sy = ((l.packed >> 7) & 3) - 1;
        const float fv = 2.0f * (ring ? kW[inv_k] : cur[inv_k * plane + l.offset]);  // ring: f_new keeps its init value
        fx += (double)(fv * (float)sx);
        fy += (double)(fv * (float)sy);
    }
    __shared__ double sx_[32], sy_[32];
    for (int s = 16; s > 0; s >>= 1) {
        fx += __shfl_xor_sync(0xffffffffu, fx, s);
        fy += __shfl_xor_sync(0xffffffffu, fy, s);
    }
    const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if ((threadIdx.x & 31) == 0) { sx_[w] = fx; sy_[w] = fy; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ax = 0.0, ay = 0.0;
        for (int i = 0; i < nw; ++i) { ax += sx_[i]; ay += sy_[i]; }
        partial[blockIdx.x * 2] = ax;
        partial[blockIdx.x * 2 + 1] = ay;
    }
}
__global__ void force_final_kernel(const double *partial, int n, float *out2) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double ax = 0.0, ay = 0.0;
        for (int i = 0; i < n; ++i) { ax += partial[2 * i]; ay += partial[2 * i + 1]; }
        out2[0] = (float)ax;
        out2[1] = (float)ay;
    }
}

}  // namespace lbm
