// Kernels of the fused D2Q9 MRT-LES step (sm_100a).  See DESIGN.md for the data layout.
//
// Layout in HBM: 9 SoA planes per buffer, each (nx_local, pitch) with y fastest and
// pitch = round_up(ny, 64) floats, so every column starts on a 128-byte line and a warp's
// vector accesses are full, aligned lines.  Two buffers (src/dst) swap roles every step.
// Variants: step_kernel (register, default; this file), step_tma_kernel (lbm2d_tma.cuh).
#pragma once
#include "lbm2d_device.cuh"

namespace lbm {

// Tuning knobs (measured on B200, 8192x2048: profiles/r01_tuning_sweep.md).  Small CTAs (64-128 threads) win:
// the warps of a CTA move through load / math / store in lock-step, so many small CTAs per SM keep the
// memory pipeline evenly fed.  Two cells per thread (64-bit accesses) beat one and four (same file).
#ifndef LBM_WPB
#define LBM_WPB 4
#endif
#ifndef LBM_MINB
#define LBM_MINB 10          // fast arithmetic: 48 registers, 40 resident warps per SM (profiles/r01_tuning_sweep.md)
#endif
#ifndef LBM_MINB_STRICT
#define LBM_MINB_STRICT 9    // strict arithmetic: 56 registers, 36 warps: -5 % step time (profiles/r02_tuning.md)
#endif
constexpr int kWarpsPerBlock = LBM_WPB;
static_assert(LBM_WPB >= 2, "the top / bottom ring row of a group needs two warps");
constexpr int kRingGroup = 32;   // interior columns per top/bottom ring row of the grid (one lane per column)
constexpr int kThreads = kWarpsPerBlock * 32;
constexpr int kCellsPerThread = 2;                       // one 64-bit access per plane, one packed fp32 pair per value
constexpr int kSegCells = 32 * kCellsPerThread;          // cells per warp; the plane pitch is a multiple of this

struct StepArgs {
    const float *__restrict__ src;  // 9 planes
    float *__restrict__ dst;        // 9 planes
    // plane k of the source buffer shifted by the pull: srcp[k] + il * pitch + j is f_k(il - e_kx, j), and plane k of
    // the destination.  Kernel parameters live in the constant bank, so an access is ONE instruction (32-bit cell
    // offset * 4 + 64-bit constant) instead of a 64-bit multiply-add chain per plane.
    const float *srcp[9];
    float *dstp[9];
    const uint8_t *__restrict__ code;  // cell code, bit0 = solid
    const uint8_t *__restrict__ links8;  // bounce-back mode only: bit k-1 set = the upstream neighbour i - e_k of this FLUID cell is solid
    const uint32_t *__restrict__ code_bits;  // the same bit, 32 cells per word (plane order): what the interior warps read
    const float *__restrict__ damp_x;  // [nx_local]   ref:364-370 (indexed by local column, holds the global value)
    const float *__restrict__ damp_y;  // [pitch]      ref:372-378
    float *rho, *ux, *uy;           // macroscopic planes (written by EMIT steps)
    unsigned *maxv_bits;            // max(ux^2+uy^2) as ordered uint; [1] = NaN flag
    long long plane;                // floats per plane = nx_local * pitch
    int nx_local, ny, pitch, nseg;
    int x_off;                      // global x of local column 0
    int west_ring, east_ring;       // local column 0 / nx_local-1 is the domain boundary (else a halo)
    float ramp;                     // soft-start factor of this step, ref:442-443 (host table, see lbm2d_capi.cu)
    int il0, il_step, il_count;     // columns of this launch: il0 + blockIdx.y * il_step, blockIdx.y < il_count
    int n_ring;                     // ring cells handled by this launch's ring warps
    int ring_row0, ring_rows;       // grid rows [ring_row0, ring_row0 + ring_rows): W/E ring warps; the others: see step_kernel
    // Early start (see step_kernel): rows [0, early_rows) may begin on the progress counter instead of the full
    // completion of the previous step; rows [0, low_rows) of every step add 1 per CTA to it when done.
    int early_rows, low_rows;
    unsigned long long *progress;
    unsigned long long progress_expected;   // counter value once the previous step's rows [0, low_rows) are complete
    const RingCtx *ring;            // rare-path context in global memory (dst-specific; TMA variant)
    // Column order of the grid: col < col_split -> il = 1 + col; col == col_split -> il = nx_local - 2 (the east edge
    // column of a slab, see below); col > col_split -> il = col.  Single GPU / NCCL launches: il0 + col * il_step
    // (col_split < 0).
    int col_split;
    // x-slabs over peer memory (lbm_peer_connect): the CTAs of an edge column -- the first / last owned column next to a
    // halo -- also store the three populations that stream across the interface straight into the neighbour's halo
    // column (NVLink peer stores), then add 1 to the neighbour's inbox counter (system-scope release); before they
    // touch their own halo they wait until their inbox shows that the neighbour's edge CTAs of the PREVIOUS step are
    // done (which also means the neighbour no longer reads the halo column this step overwrites).  side 0 = west.
    int edge_il[2];                       // local column of the edge on that side, or -1 (domain boundary / no peer mode)
    int edge_row[2];                      // the grid row (blockIdx.y + 65535 blockIdx.z) of that column, or -1
    float *peer_dst[2][3];                // neighbour's halo column in ITS destination buffer, planes kHaloPlane[side][.]
    unsigned long long *peer_inbox[2];    // the neighbour's counter this rank adds to
    const unsigned long long *inbox;      // this rank's counters [2]: written by the west / east neighbour
    unsigned long long inbox_expected;    // value once the neighbour's edge CTAs of the previous step have all signalled
    unsigned *peer_error;                 // set if a wait timed out (the step then continues on stale data; host reports it)
    Physics phys;
};

// populations that stream across an interface: side 0 (to the west neighbour) e_x = -1, side 1 (to the east) e_x = +1
__device__ constexpr int kHaloPlane[2][3] = {{3, 6, 7}, {1, 5, 8}};

__device__ __forceinline__ int col_to_il(const StepArgs &a, int col) {
    if (a.col_split < 0) return a.il0 + col * a.il_step;
    return col < a.col_split ? 1 + col : (col == a.col_split ? a.nx_local - 2 : col);
}

// Wait (one thread) until this rank's inbox from `side` reaches `expected`; bounded: ~2 s, then flag an error.
__device__ __forceinline__ void inbox_wait(const StepArgs &a, int side) {
    unsigned long long seen, t0 = 0;
    for (unsigned spins = 0;; ++spins) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(a.inbox + side) : "memory");
        if (seen >= a.inbox_expected) return;
        if ((spins & 1023u) == 1023u) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 2000000000ULL) {
                atomicExch(a.peer_error, 1u);
                return;
            }
        }
        __nanosleep(64);
    }
}

__device__ __forceinline__ float vmag2_strict(float ux, float uy) {
    return __fadd_rn(__fmul_rn(ux, ux), __fmul_rn(uy, uy));
}

// Population loads bypass L1 (ld.global.cg): every value is read exactly once per step; measured 1 % faster.
#define LBM_LD(ptr) __ldcg(ptr)

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

template <bool STRICT, bool EMIT, bool BB, bool PEER>
__device__ __forceinline__ void ring_cell(const StepArgs &a, int idx, float ramp, float &vmax, int &vnan) {
    const int ny = a.ny, pitch = a.pitch, n = a.il_count;
    const long long plane = a.plane;
    // decode: ring cell (ilr, jr), its owner (ilo, jo), boundary side dr; corners chain W/E -> top/bottom
    int ilr, jr, ilo, jo, dr, corner_dr = -1;
    if (idx < 2 * n) {
        const bool top = idx < n;
        ilo = ilr = PEER ? col_to_il(a, top ? idx : idx - n) : a.il0 + (top ? idx : idx - n) * a.il_step;
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
    if (PEER)
#pragma unroll
    for (int side = 0; side < 2; ++side)   // top / bottom cell of an edge column: also part of the neighbour's halo column
        if (ilr == a.edge_il[side]) {
#pragma unroll
            for (int q = 0; q < 3; ++q) a.peer_dst[side][q][jr] = r.f[kHaloPlane[side][q]];
        }
    if (EMIT) {
        a.rho[o] = r.rho;
        a.ux[o] = r.ux;
        a.uy[o] = r.uy;
        const float m2 = __fadd_rn(__fmul_rn(r.ux, r.ux), __fmul_rn(r.uy, r.uy));
        vnan |= (m2 != m2);
        vmax = fmaxf(vmax, m2);
    }
}

// Stores of one thread's two cells (j0, j0 + 1 of local column t / pitch): the nine populations, on EMIT steps rho / u and
// the max|u|^2 bookkeeping, on a slab edge column (peer-memory path) also the neighbour's halo column.
template <bool EMIT, bool PEER>
__device__ __forceinline__ void store_cells(const StepArgs &a, int t, int j0, const float (&g)[2][9], const float (&rho)[2],
                                            const float (&ux)[2], const float (&uy)[2], bool edge_w, bool edge_e, float &vmax,
                                            int &vnan) {
    const int ny = a.ny;
    (void)ny;
    const bool lo_int = j0 >= 1 && j0 <= ny - 2, hi_int = j0 + 1 <= ny - 2;   // interior cells (ring cells: ring warps)
    if (lo_int && hi_int) {            // the common case: one 64-bit store per plane
#pragma unroll
        for (int k = 0; k < 9; ++k) *reinterpret_cast<float2 *>(a.dstp[k] + t) = make_float2(g[0][k], g[1][k]);
        if (EMIT) {
            *reinterpret_cast<float2 *>(a.rho + t) = make_float2(rho[0], rho[1]);
            *reinterpret_cast<float2 *>(a.ux + t) = make_float2(ux[0], ux[1]);
            *reinterpret_cast<float2 *>(a.uy + t) = make_float2(uy[0], uy[1]);
        }
    } else if (lo_int || hi_int) {     // the pair shares a ring cell (ring warps write it) or padding: cell by cell
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            if (!(c == 0 ? lo_int : hi_int)) continue;
#pragma unroll
            for (int k = 0; k < 9; ++k) a.dstp[k][t + c] = g[c][k];
            if (EMIT) { a.rho[t + c] = rho[c]; a.ux[t + c] = ux[c]; a.uy[t + c] = uy[c]; }
        }
    }
    if (PEER && (edge_w || edge_e)) {   // the neighbour's halo column: the same cells, the three planes it will pull
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            if (!(side == 0 ? edge_w : edge_e)) continue;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int k = kHaloPlane[side][q];
                float *p = a.peer_dst[side][q] + j0;
                if (lo_int && hi_int) *reinterpret_cast<float2 *>(p) = make_float2(g[0][k], g[1][k]);
                else {
                    if (lo_int) p[0] = g[0][k];
                    if (hi_int) p[1] = g[1][k];
                }
            }
        }
    }
    if (EMIT) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            if (!(c == 0 ? lo_int : hi_int)) continue;
            const float m2 = vmag2_strict(ux[c], uy[c]);
            vnan |= (m2 != m2);
            vmax = fmaxf(vmax, m2);
        }
    }
}

// The general tail of an interior pair: rho / u where consumed (EMIT steps, solids), obstacle refill, stores.
template <bool STRICT, bool EMIT, bool BB, bool PEER>
__device__ __forceinline__ void finish_cells(const StepArgs &a, int t, int j0, unsigned code2, float (&g)[2][9], bool edge_w,
                                             bool edge_e, float &vmax, int &vnan) {
    float rho[2] = {0.f, 0.f}, ux[2] = {0.f, 0.f}, uy[2] = {0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const bool solid = (code2 >> c) & 1u;
        if (EMIT || solid) macro_from_f<STRICT>(g[c], rho[c], ux[c], uy[c]);
        if (solid) {  // obstacle refill, ref:452-455 (bounce-back mode: frozen at rest)
            ux[c] = 0.0f; uy[c] = 0.0f;
            if (BB) rho[c] = 1.0f;
#pragma unroll
            for (int k = 0; k < 9; ++k) g[c][k] = __fmul_rn(kW[k], rho[c]);
        }
    }
    store_cells<EMIT, PEER>(a, t, j0, g, rho, ux, uy, edge_w, edge_e, vmax, vnan);
}

// Interior warp: one 64-cell segment of one interior column (see step_kernel).  `edge_w` / `edge_e`: the column is a slab
// edge on the peer-memory path (warp-uniform) and its outgoing populations also go to the neighbour's halo column.
template <bool STRICT, bool EMIT, bool BB, bool PEER>
__device__ __forceinline__ void interior_warp(const StepArgs &a, int il, int seg, int lane, bool edge_w, bool edge_e,
                                              float &vmax, int &vnan) {
    const int j0 = seg * kSegCells + lane * 2;                          // < pitch: the pitch is a multiple of 64
    const int ny = a.ny;
    const int t = il * a.pitch + j0;                                     // cell offset inside a plane (< 2^31, checked at create)

    // pull (ref:254-257): fin[k] = f_k(i - e_kx, j - e_ky) for j = j0, j0 + 1; every load issued before the first use
    float2 v[9];
    float edge[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float *p = a.srcp[k] + t;
        v[k] = LBM_LD(reinterpret_cast<const float2 *>(p));
        edge[k] = 0.f;
        // segment ends: the value of the neighbouring segment (seg 0 / the last one: a ring or padding cell's input, unused)
        if (kEy[k] == 1 && lane == 0 && seg > 0) edge[k] = LBM_LD(p - 1);
        if (kEy[k] == -1 && lane == 31) edge[k] = LBM_LD(p + 2);   // at most one float past the row: inside the allocation
    }
    // No `live` branch: lanes in the padding of the last segment (ny % 64 != 0) run the same code on the rest-state
    // values the padding holds (init_kernel) and store nothing -- a branch here makes ptxas sink the three loads that
    // feed no shuffle below the shuffles, i.e. behind a full memory latency (4.5 us of a 186 us step).
    const float dx = __ldg(a.damp_x + il);
    const float2 dy = __ldg(reinterpret_cast<const float2 *>(a.damp_y + j0));
    // solid bits of the warp's 64 cells = 2 consecutive words; a lane's 2 cells sit in one of them
    const unsigned code2 = (__ldg(a.code_bits + (t >> 5)) >> (j0 & 31)) & 3u;
    unsigned char links[2] = {0, 0};
    if (BB) {
        const uchar2 l2 = __ldg(reinterpret_cast<const uchar2 *>(a.links8 + t));
        links[0] = l2.x;
        links[1] = l2.y;
    }
    float fin[2][9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        if (kEy[k] == 0) {
            fin[0][k] = v[k].x;
            fin[1][k] = v[k].y;
        } else if (kEy[k] == 1) {  // needs j-1: last element of the lane below
            float below = __shfl_up_sync(0xffffffffu, v[k].y, 1);
            if (lane == 0) below = edge[k];
            fin[0][k] = below;
            fin[1][k] = STRICT ? opaque_copy(v[k].x) : v[k].x;   // see opaque_copy(): the pair (below, v.x) in natural order
        } else {                   // needs j+1: first element of the lane above
            float above = __shfl_down_sync(0xffffffffu, v[k].x, 1);
            if (lane == 31) above = edge[k];
            fin[0][k] = STRICT ? opaque_copy(v[k].y) : v[k].y;
            fin[1][k] = above;
        }
    }
    {
        if (BB) {
#pragma unroll
            for (int c = 0; c < 2; ++c)
                if (links[c]) bounce_back(a, links[c], (long long)t + c, fin[c]);
        }
        // collide (ref:266-420)
        float g[2][9];
        if (STRICT) {
            f32x2 f2[9], g2[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) f2[k] = pack2(fin[0][k], fin[1][k]);
            // A non-finite relaxed moment (a run that is blowing up) takes the dense inverse AND its own copy of the tail:
            // merging the two results in front of a shared tail cost six register moves on the hot path.
            f32x2 ms2[9];
            if (!collide_strict_front<Lane2>(a.phys, f2, pack2(fmaxf(dx, dy.x), fmaxf(dx, dy.y)), ms2)) {
                Lane2::inverse_dense(ms2, g2);
#pragma unroll
                for (int k = 0; k < 9; ++k) unpack2(g2[k], g[0][k], g[1][k]);
                finish_cells<STRICT, EMIT, BB, PEER>(a, t, j0, code2, g, edge_w, edge_e, vmax, vnan);
                return;
            }
            collide_strict_back<Lane2>(ms2, g2);
#pragma unroll
            for (int k = 0; k < 9; ++k) unpack2(g2[k], g[0][k], g[1][k]);
        } else {
            collide_fast(a.phys, fin[0], fmaxf(dx, dy.x), g[0]);
            collide_fast(a.phys, fin[1], fmaxf(dx, dy.y), g[1]);
        }
        // rho / u (ref:425-436) only where consumed: EMIT steps and the obstacle refill.  The common case -- no solid
        // cell in the pair, no EMIT -- stores straight from the collision's registers; the other path has its own copy of
        // the stores (a shared tail would cost a register move per value at the merge: 14 MOVs on the hot path).
        float rho[2] = {0.f, 0.f}, ux[2] = {0.f, 0.f}, uy[2] = {0.f, 0.f};
        if (!EMIT && code2 == 0) {
            store_cells<EMIT, PEER>(a, t, j0, g, rho, ux, uy, edge_w, edge_e, vmax, vnan);
        } else {
            finish_cells<STRICT, EMIT, BB, PEER>(a, t, j0, code2, g, edge_w, edge_e, vmax, vnan);
        }
    }
}

// One fused pass: pull-stream, MRT-LES collision, sponge, macroscopic update, boundary ring,
// obstacle refill (ref:552-573 = collide_and_stream + update_macro_var + apply_bc), f_src -> f_dst.
//
// Interior warps: one warp = one 64-cell segment of one interior column, one thread = 2 consecutive cells in y;
// every access is an aligned, fully coalesced 64-bit access, the +-1 shift of the pull in y comes from the
// neighbouring lane by warp shuffle with one extra scalar load at each end of the segment, all issued before
// first use; no boundary code at all.  The two cells of a thread travel through the strict collision as the two
// halves of packed fp32 pairs (Lane2).  Ring warps (their own grid rows: one row behind every 32 columns for the
// top / bottom cells, one block of rows for the W / E columns): one ring cell per lane, see above.
// BB: optional half-way bounce-back obstacle mode (not the reference's).
// PEER: x-slab launch on the peer-memory halo path (edge-column waits / remote stores / counters compiled in).
template <bool STRICT, bool EMIT, bool BB = false, bool PEER = false>
__global__ void __launch_bounds__(kThreads, STRICT ? LBM_MINB_STRICT : LBM_MINB) step_kernel(const StepArgs a) {
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
    // frame_count (ref:440): the ramp value comes from the host as a kernel argument and nothing on the device reads a
    // step counter -- with early start a CTA of step n+1 may run before the last ring warp of step n.  The diagnostic
    // device copy of the counter is set once per lbm_run (set_counter_kernel), not by every CTA's prologue.

    float vmax = 0.0f;  // max |u|^2 over the cells written by this thread (EMIT only)
    int vnan = 0;
    // slab edge column (CTA-uniform; both false off the peer-memory slab path): wait for the neighbour's previous step
    // The two edge columns sit in grid rows the host knows (edge_row[side], -1 = none): two uniform compares per CTA
    // decide; every other column maps with one compare (col_to_il's middle case IS the east edge row).
    const bool edge_w = PEER && row == a.edge_row[0], edge_e = PEER && row == a.edge_row[1];
    const int il_cta = !PEER ? a.il0 + col * a.il_step
                             : ((edge_w || edge_e) ? (edge_e ? a.edge_il[1] : a.edge_il[0]) : (col < a.col_split ? col + 1 : col));
    if (PEER && (edge_w || edge_e)) {
        if (threadIdx.x == 0) {
            if (edge_w) inbox_wait(a, 0);
            if (edge_e) inbox_wait(a, 1);
        }
        __syncthreads();
    }
    if (we_row || tb_row) {
        // ------------------------------- ring warps (ONE call site of ring_cell: it is ~1 500 instructions) ----------
        int ring_idx = -1;
        bool edge_here[2] = {false, false};
        if (we_row) {   // W / E columns and corners
            const int idx = 2 * a.il_count + (ring_rel * (int)gridDim.x * kWarpsPerBlock + seg) * 32 + lane;
            if (idx < a.n_ring) ring_idx = idx;
        } else if (blockIdx.x == 0) {   // top (warp 0) / bottom (warp 1) cells of one group of 32 columns
            const int c = grp * kRingGroup + lane;
            if (threadIdx.x < 64 && c < a.il_count) ring_idx = (threadIdx.x >> 5) * a.il_count + c;
            if (PEER) {   // does this group hold a slab edge column?  (CTA-uniform)  Its ring cells read the halo and go to the neighbour too.
                const int c0 = grp * kRingGroup, c1 = min(c0 + kRingGroup, a.il_count);
#pragma unroll
                for (int side = 0; side < 2; ++side)
                    if (a.edge_il[side] >= 0)
                        for (int cc = c0; cc < c1; ++cc) edge_here[side] |= col_to_il(a, cc) == a.edge_il[side];
                if (edge_here[0] || edge_here[1]) {
                    if (threadIdx.x == 0) {
                        if (edge_here[0]) inbox_wait(a, 0);
                        if (edge_here[1]) inbox_wait(a, 1);
                    }
                    __syncthreads();
                }
            }
        }
        if (ring_idx >= 0) ring_cell<STRICT, EMIT, BB, PEER>(a, ring_idx, a.ramp, vmax, vnan);
        if (PEER && (edge_here[0] || edge_here[1])) {
            __syncthreads();
            if (threadIdx.x == 0) {
                __threadfence_system();
                if (edge_here[0]) atomicAdd_system(a.peer_inbox[0], 1ULL);
                if (edge_here[1]) atomicAdd_system(a.peer_inbox[1], 1ULL);
            }
        }
    } else if (col < a.il_count && seg < a.nseg) {   // nseg = ceil(ny / 64): the segment starts inside the column
        // ------------------------------- interior warps --------------------------------------
        // slab edge columns (2 of thousands) take their own copy of the interior code with the neighbour stores compiled
        // in; every other column runs exactly the single-GPU code (sharing one copy cost 3.4 % of the step on ALL columns)
        interior_warp<STRICT, EMIT, BB, false>(a, il_cta, seg, lane, false, false, vmax, vnan);
        // A slab edge column also goes to the neighbour's halo column: every thread forwards the three populations of
        // ITS OWN two cells, re-read from the destination buffer (a thread sees its own stores).  A second copy of the
        // interior code with the remote stores inlined measured 1.1 % slower on EVERY column (instruction cache).
        if (PEER && (edge_w || edge_e)) {
            const int j0 = seg * kSegCells + lane * 2, t = il_cta * a.pitch + j0;
            const bool lo_int = j0 >= 1 && j0 <= a.ny - 2, hi_int = j0 + 1 <= a.ny - 2;
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                if (!(side == 0 ? edge_w : edge_e)) continue;
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const float *src = a.dstp[kHaloPlane[side][q]] + t;
                    float *dst = a.peer_dst[side][q] + j0;
                    if (lo_int && hi_int) *reinterpret_cast<float2 *>(dst) = __ldcg(reinterpret_cast<const float2 *>(src));
                    else {
                        if (lo_int) dst[0] = __ldcg(src);
                        if (hi_int) dst[1] = __ldcg(src + 1);
                    }
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
    if (PEER && (edge_w || edge_e)) {   // edge-column CTA: its part of the neighbour's halo column is complete
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            if (edge_w) atomicAdd_system(a.peer_inbox[0], 1ULL);
            if (edge_e) atomicAdd_system(a.peer_inbox[1], 1ULL);
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

// End of a batch on the peer-memory slab path: later work on the stream (force, exports) reads the halo columns of the
// final step, which the neighbours write.
__global__ void halo_wait_kernel(const StepArgs a) {
    if (threadIdx.x == 0) {
        if (a.edge_il[0] >= 0) inbox_wait(a, 0);
        if (a.edge_il[1] >= 0) inbox_wait(a, 1);
    }
}

// The diagnostic device copy of frame_count (lbm_step_count): set once behind the steps of an lbm_run call -- the step kernels
// carry no step index (which is also what lets a batch be replayed as a CUDA graph).
__global__ void set_counter_kernel(int *ctr, int value) { *ctr = value; }

// Self test of Lane2's inline division / square root (lbm_selftest_arith): random operands from the guarded box
// against __fdiv_rn / __fsqrt_rn, bit for bit.  out[0..2] = mismatches of a / b (b in [1/8, 8]), 1 / t (t in
// [2^-10, 2^40]) and sqrt(x).
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long &x) {
    unsigned long long z = (x += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float float_from(unsigned mant, int exp2, bool negative) {   // 1.mant * 2^exp2
    return __uint_as_float((negative ? 0x80000000u : 0u) | ((unsigned)(exp2 + 127) << 23) | (mant & 0x7fffffu));
}
__global__ void selftest_arith_kernel(long long n_per_thread, unsigned long long seed, unsigned long long *out) {
    unsigned long long st = seed + 0x632be59bd9b4e019ULL * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1);
    unsigned long long bad[3] = {0, 0, 0};
    for (long long it = 0; it < n_per_thread; ++it) {
        float a[2], b[2], t[2], x[2];
        for (int l = 0; l < 2; ++l) {
            const unsigned long long r0 = splitmix64(st), r1 = splitmix64(st), r2 = splitmix64(st);
            // mantissas: random, or one of the hard patterns (all zeros / all ones / single bits)
            auto mant = [](unsigned long long r) {
                const unsigned m = (unsigned)(r >> 8) & 0x7fffffu, sel = (unsigned)r & 15u;
                return sel == 0 ? 0u : sel == 1 ? 0x7fffffu : sel == 2 ? 0x7ffffeu : sel == 3 ? 1u : sel == 4 ? (1u << ((r >> 40) % 23)) : m;
            };
            b[l] = float_from(mant(r0), (int)((r0 >> 44) % 6) - 3, false);                // [1/8, 8)
            if (b[l] > 8.0f) b[l] = 8.0f;
            const unsigned za = (unsigned)(r1 >> 60);
            a[l] = za == 0 ? 0.0f : float_from(mant(r1), (int)((r1 >> 44) % 160) - 100, (r1 >> 59) & 1);   // +0 or 2^-100 .. 2^60
            t[l] = float_from(mant(r2), (int)((r2 >> 44) % 50) - 10, false);              // [2^-10, 2^40)
            x[l] = float_from(mant(r0 ^ r2), (int)((r1 >> 32) % 200) - 100, false);       // [2^-100, 2^100)
        }
        const f32x2 A = pack2(a[0], a[1]), B = pack2(b[0], b[1]), T = pack2(t[0], t[1]), X = pack2(x[0], x[1]);
        float q0, q1;
        unpack2(Lane2::div_refined(A, B, Lane2::rcp_refined(B)), q0, q1);
        bad[0] += (__float_as_uint(q0) != __float_as_uint(__fdiv_rn(a[0], b[0]))) + (__float_as_uint(q1) != __float_as_uint(__fdiv_rn(a[1], b[1])));
        unpack2(Lane2::div_refined(Lane2::bc(1.0f), T, Lane2::rcp_refined(T)), q0, q1);
        bad[1] += (__float_as_uint(q0) != __float_as_uint(__fdiv_rn(1.0f, t[0]))) + (__float_as_uint(q1) != __float_as_uint(__fdiv_rn(1.0f, t[1])));
        unpack2(Lane2::sqrt_inrange(X), q0, q1);
        bad[2] += (__float_as_uint(q0) != __float_as_uint(__fsqrt_rn(x[0]))) + (__float_as_uint(q1) != __float_as_uint(__fsqrt_rn(x[1])));
    }
    for (int k = 0; k < 3; ++k)
        if (bad[k]) atomicAdd(out + k, bad[k]);
}

// init(), ref:235-241: both buffers = w_k (padding cells included), rho = 1, u = 0.
__global__ void init_kernel(float *f0, float *f1, float *rho, float *ux, float *uy, long long plane, int ny, int pitch) {
    const long long n = plane;
    for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < n; o += (long long)gridDim.x * blockDim.x) {
        const bool real = (int)(o % pitch) < ny;
#pragma unroll
        for (int k = 0; k < 9; ++k) {   // the padding cells keep this rest state for ever (never written again)
            f0[k * plane + o] = kW[k];
            f1[k * plane + o] = kW[k];
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
