// "Async" variant of the fused step: warp-private cp.async rings.
//
// Same work split as the register variant (one warp = one 64-cell segment of a column, two cells per
// thread), but each warp is persistent and keeps the NEXT segments' populations in flight in a private
// shared-memory ring filled by cp.async (LDGSTS): the loads of segment i+1, i+2 overlap the math and
// the stores of segment i without costing registers or occupancy, and need no block-level barrier.
// Rows are copied in aligned 16-byte chunks with a 4-float apron for the planes that shift in y, so
// the +-1 shift of the pull becomes a shared-memory offset (no shuffles, no edge loads).
#pragma once
#include "lbm2d_kernels.cuh"

namespace lbm {

constexpr int kASeg = 64;                        // cells per segment (2 per lane)
constexpr int kARowH = kASeg + 8;                // apron rows
// float offsets of the 9 rows inside a stage (e_ky != 0 for k = 2, 4, 5, 6, 7, 8), then damp_y, then codes
__device__ constexpr int kARowOff[9] = {0, 64, 128, 200, 264, 336, 408, 480, 552};
constexpr int kADampOff = 624, kACodeOff = 688, kAStageFloats = 704;
#ifndef LBM_ASTAGES
#define LBM_ASTAGES 3
#endif
#ifndef LBM_AWPB
#define LBM_AWPB 2
#endif
#ifndef LBM_AMINB
#define LBM_AMINB 12
#endif
constexpr int kAStages = LBM_ASTAGES;
constexpr int kAWarps = LBM_AWPB;

__device__ __forceinline__ void cp16(float *dst, const void *src, bool valid, bool l1) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int n = valid ? 16 : 0;                // src-size 0: nothing is read, the 16 bytes are zero-filled
    if (l1) asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
    else asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// queue the copies of one segment (local column il, first row j0) into `st`
__device__ __forceinline__ void async_issue(const StepArgs &a, float *st, int il, int j0, int lane) {
    const int pitch = a.pitch;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float *col = a.src + k * a.plane + (long long)(il - kEx[k]) * pitch;
        if (kEy[k] == 0) {
            const int off = j0 + 4 * lane;
            if (lane < 16) cp16(st + kARowOff[k] + 4 * lane, col + min(off, pitch - 4), off + 3 < pitch, false);
        } else {
            const int off = j0 - 4 + 4 * lane;
            if (lane < 18) cp16(st + kARowOff[k] + 4 * lane, col + min(max(off, 0), pitch - 4), off >= 0 && off + 3 < pitch, false);
        }
    }
    const int off = j0 + 4 * lane;
    if (lane < 16) cp16(st + kADampOff + 4 * lane, a.damp_y + min(off, pitch - 4), off + 3 < pitch, true);
    if (lane < 4) cp16(st + kACodeOff + 4 * lane, a.code + (long long)il * pitch + min(j0 + 16 * lane, pitch - 16), j0 + 16 * lane + 15 < pitch, true);
}

template <bool STRICT, bool EMIT>
__global__ void __launch_bounds__(32 * kAWarps, LBM_AMINB) step_async_kernel(const StepArgs a) {
    extern __shared__ __align__(16) float smem_a[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float *ring = smem_a + wib * (kAStages * kAStageFloats);
    const int wid = blockIdx.x * kAWarps + wib, nwarps = gridDim.x * kAWarps;
    if (a.bump_ctr && blockIdx.x == 0 && threadIdx.x == 0) *a.ctr_out = *a.ctr_in + 1;  // ref:440
    const int nseg = (a.pitch + kASeg - 1) / kASeg;
    const int n_items = a.il_count * nseg;
    const int ny = a.ny, pitch = a.pitch;
    const long long plane = a.plane;
    float vmax = 0.0f;
    int vnan = 0;
    const float ramp = __ldg(a.ramp_tab + min(*a.ctr_in + 1, a.warmup));

    // prologue: fill the first kAStages-1 stages
#pragma unroll
    for (int s = 0; s < kAStages - 1; ++s) {
        const int item = wid + s * nwarps;
        if (item < n_items) async_issue(a, ring + s * kAStageFloats, a.il0 + (item / nseg) * a.il_step, (item % nseg) * kASeg, lane);
        cp_commit();
    }
    int stage = 0;
    for (int item = wid; item < n_items; item += nwarps) {
        {   // keep the ring full: the stage consumed in the previous iteration gets the segment kAStages-1 ahead
            const int nxt = item + (kAStages - 1) * nwarps;
            const int ns = (stage + kAStages - 1) % kAStages;
            if (nxt < n_items) async_issue(a, ring + ns * kAStageFloats, a.il0 + (nxt / nseg) * a.il_step, (nxt % nseg) * kASeg, lane);
            cp_commit();
        }
        cp_wait<kAStages - 1>();
        __syncwarp();
        const float *st = ring + stage * kAStageFloats;
        const int il = a.il0 + (item / nseg) * a.il_step;
        const int j0 = (item % nseg) * kASeg + 2 * lane;

        float fin[2][9];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            if (kEy[k] == 0) {
                const float2 t = *reinterpret_cast<const float2 *>(st + kARowOff[k] + 2 * lane);
                fin[0][k] = t.x; fin[1][k] = t.y;
            } else {
                fin[0][k] = st[kARowOff[k] + 4 + 2 * lane - kEy[k]];
                fin[1][k] = st[kARowOff[k] + 5 + 2 * lane - kEy[k]];
            }
        }
        const float2 dy2 = *reinterpret_cast<const float2 *>(st + kADampOff + 2 * lane);
        const uchar2 code2 = *reinterpret_cast<const uchar2 *>(reinterpret_cast<const unsigned char *>(st + kACodeOff) + 2 * lane);
        __syncwarp();   // every lane has read its stage before the next iteration refills it
        stage = (stage + 1) % kAStages;

        if (j0 < ny) {
            const float dy[2] = {dy2.x, dy2.y};
            const unsigned char code[2] = {code2.x, code2.y};
            const float dx = __ldg(a.damp_x + il);
            const bool edge_col = (il == 1 && a.west_ring) || (il == a.nx_local - 2 && a.east_ring);
            const bool touches_ring = (j0 <= 1) || (j0 + 2 >= ny - 1) || edge_col;
            float g[2][9], rho[2], ux[2], uy[2];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const float damp = fmaxf(dx, dy[c]);
                if (STRICT) collide_strict(a.phys, fin[c], damp, g[c]);
                else collide_fast(a.phys, fin[c], damp, g[c]);
                rho[c] = ux[c] = uy[c] = 0.0f;
                if (EMIT || touches_ring || (code[c] & 1)) macro_from_f<STRICT>(g[c], rho[c], ux[c], uy[c]);
            }
            Cell own[2];
            if (touches_ring) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
#pragma unroll
                    for (int k = 0; k < 9; ++k) own[c].f[k] = g[c][k];
                    own[c].rho = rho[c]; own[c].ux = ux[c]; own[c].uy = uy[c];
                }
            }
            const bool any_interior = (j0 + 1 >= 1 && j0 <= ny - 2);
            if (any_interior) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int j = j0 + c;
                    const bool interior = (j >= 1) && (j <= ny - 2);
                    if (interior && (code[c] & 1)) {
                        ux[c] = 0.0f; uy[c] = 0.0f;
#pragma unroll
                        for (int k = 0; k < 9; ++k) g[c][k] = __fmul_rn(kW[k], rho[c]);
                    }
                    if (!interior) {
                        rho[c] = 0.0f; ux[c] = 0.0f; uy[c] = 0.0f;
#pragma unroll
                        for (int k = 0; k < 9; ++k) g[c][k] = 0.0f;
                    }
                }
                const long long o = (long long)il * pitch + j0;
#pragma unroll
                for (int k = 0; k < 9; ++k) *reinterpret_cast<float2 *>(a.dst + k * plane + o) = make_float2(g[0][k], g[1][k]);
                if (EMIT) {
                    *reinterpret_cast<float2 *>(a.rho + o) = make_float2(rho[0], rho[1]);
                    *reinterpret_cast<float2 *>(a.ux + o) = make_float2(ux[0], ux[1]);
                    *reinterpret_cast<float2 *>(a.uy + o) = make_float2(uy[0], uy[1]);
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const float m2 = vmag2_strict(ux[c], uy[c]);
                        vnan |= (m2 != m2);
                        vmax = fmaxf(vmax, m2);
                    }
                }
            }
            if (touches_ring) {
                TileSink sink;
                sink.sm_f = nullptr;
                sink.sm_mac = nullptr;
                sink.il0 = sink.j0 = sink.bx = sink.by = sink.row_hi = sink.col_lo = sink.col_hi = 0;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int j = j0 + c;
                    if (j < 1 || j > ny - 2) continue;
                    if (!(j == 1 || j == ny - 2 || edge_col)) continue;
                    ring_from_owner(a.ring, &sink, EMIT, il, j, &own[c], ramp, &vmax, &vnan);
                }
            }
        }
    }
    cp_wait<0>();
    if (EMIT) {
        for (int s = 16; s > 0; s >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, s));
        const bool any_nan = __any_sync(0xffffffffu, vnan != 0);
        if (lane == 0) {
            const unsigned bits = __float_as_uint(vmax);
            if (bits > *a.maxv_bits) atomicMax(a.maxv_bits, bits);
            if (any_nan) a.maxv_bits[1] = 1u;
        }
    }
}

}  // namespace lbm
