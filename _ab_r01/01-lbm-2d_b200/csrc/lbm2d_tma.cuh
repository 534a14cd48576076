// Persistent TMA / mbarrier variant of the fused step (sm_100a).
//
// One CTA per SM loops over (BX x BY)-cell tiles.  A producer warp streams the 9 *already shifted*
// source planes of the next tiles into a ring of shared-memory stages with
// cp.async.bulk.tensor (TMA): plane k is fetched at tile origin - e_k, so the hardware does the
// pull-streaming, including the unaligned +-1 shift in y that costs shuffles and edge loads in the
// register variant, and zero-fills out-of-range coordinates.  Consumer warps collide one cell per
// thread from shared memory into an output tile, which one thread hands back to TMA
// (cp.async.bulk.tensor shared -> global).  Loads, math and stores of different tiles overlap; the
// bytes in flight per SM are set by the stage count, not by registers or occupancy.
#pragma once
#include <cuda.h>

#include "lbm2d_kernels.cuh"

namespace lbm {

constexpr int kTileBX = 8;     // columns per tile
constexpr int kTileBY = 128;   // rows per tile (fast dimension, 512-byte TMA rows)
constexpr int kTileCells = kTileBX * kTileBY;
constexpr int kInStages = 3;
constexpr int kOutStages = 2;
constexpr int kConsumerWarps = 16;
constexpr int kConsumers = 32 * kConsumerWarps;
constexpr int kTmaThreads = 32 * (2 + kConsumerWarps);   // + load-producer warp + store warp
// TMA needs a 16-byte aligned start address, so the +-1 float shift of the pull in y cannot be put
// into the box origin.  Planes with e_ky != 0 are fetched with a 4-float apron on both sides of the
// tile rows (aligned origin j0 - 4, rows of BY + 8 floats) and read at offset 4 - e_ky; the shift in
// x is a whole row of the tensor and goes into the box origin directly.
constexpr int kHaloY = 4;
constexpr int kRowHalo = kTileBY + 2 * kHaloY;
// per-plane row length and float offset inside a stage (e_ky != 0 for k = 2, 4, 5, 6, 7, 8)
constexpr int kRowN = kTileBY * kTileBX, kRowH = kRowHalo * kTileBX;
__device__ constexpr int kPlaneRow[9] = {kTileBY, kTileBY, kRowHalo, kTileBY, kRowHalo, kRowHalo, kRowHalo, kRowHalo, kRowHalo};
__device__ constexpr int kPlaneOff[10] = {0,
                                          kRowN,
                                          2 * kRowN,
                                          2 * kRowN + kRowH,
                                          3 * kRowN + kRowH,
                                          3 * kRowN + 2 * kRowH,
                                          3 * kRowN + 3 * kRowH,
                                          3 * kRowN + 4 * kRowH,
                                          3 * kRowN + 5 * kRowH,
                                          3 * kRowN + 6 * kRowH};
constexpr int kCodeOff = (3 * kRowN + 6 * kRowH) * 4;                 // byte offset of the cell-code tile
constexpr int kStageInBytes = kCodeOff + kTileCells;                  // 9 fp32 planes + 1 byte cell codes
constexpr int kStageInStride = (kStageInBytes + 127) / 128 * 128;
constexpr int kStageOutBytes = 12 * kTileCells * 4;                   // 9 f planes + rho, ux, uy
constexpr int kTmaSmemBytes = kInStages * kStageInStride + kOutStages * kStageOutBytes + 1024;

struct TmaArgs {
    const float *__restrict__ damp_x;
    const float *__restrict__ damp_y;
    const float *__restrict__ ramp_tab;
    const int *ctr_in;
    int *ctr_out;
    unsigned *maxv_bits;
    const RingCtx *ring;
    int nx_local, ny, pitch;
    int n_tx, n_ty, n_tiles;
    int col_lo, col_hi, row_hi;   // store-tensor extent (local columns [col_lo, col_hi), rows [0, row_hi))
    int west_ring, east_ring;
    int warmup;
    Physics phys;
};

// ---- thin PTX wrappers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, int c0, int c1, int c2, const void *src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map), "r"(c0),
                 "r"(c1), "r"(c2), "r"(smem_u32(src))
                 : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Tensor maps (all fp32 except the cell codes), built on the host by lbm2d_capi.cu:
//   map_src : 3-D (pitch, nx_local, 9) over the source buffer, box (BY, BX, 1)      (planes with e_ky == 0)
//   map_srch: same tensor, box (BY + 8, BX, 1)                                       (planes with e_ky != 0)
//   map_code: 2-D (pitch, nx_local) uint8, box (BY, BX)
//   map_dst : 3-D (row_hi, col_hi - col_lo, 9) over the destination buffer starting at column col_lo, box (BY, BX, 9)
//   map_mac : 3-D (row_hi, col_hi - col_lo, 3) over the rho / ux / uy planes (EMIT), box (BY, BX, 3)
template <bool STRICT, bool EMIT>
__global__ void __launch_bounds__(kTmaThreads, 1)
step_tma_kernel(const __grid_constant__ CUtensorMap map_src, const __grid_constant__ CUtensorMap map_srch,
                const __grid_constant__ CUtensorMap map_code,
                const __grid_constant__ CUtensorMap map_dst, const __grid_constant__ CUtensorMap map_mac,
                const TmaArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *in_base = smem;
    float *out_base = reinterpret_cast<float *>(smem + kInStages * kStageInStride);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kInStages * kStageInStride + kOutStages * kStageOutBytes);
    uint64_t *full = bars, *empty = bars + kInStages;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t *ofull = bars + 2 * kInStages, *oempty = ofull + kOutStages;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kInStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kConsumerWarps);
        }
        for (int o = 0; o < kOutStages; ++o) {
            mbar_init(&ofull[o], kConsumerWarps);
            mbar_init(&oempty[o], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (blockIdx.x == 0) *a.ctr_out = *a.ctr_in + 1;  // ref:440
    }
    __syncthreads();

    if (warp == 0) {
        // ===================== load producer: TMA loads of the shifted planes =====================
        if (lane == 0) {
            int it = 0;
            for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x, ++it) {
                const int s = it % kInStages;
                const uint32_t ph = (it / kInStages) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                const int il0 = a.col_lo + (t / a.n_ty) * kTileBX, j0 = (t % a.n_ty) * kTileBY;
                unsigned char *st = in_base + s * kStageInStride;
                mbar_expect_tx(&full[s], kStageInBytes);
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    if (kEy[k] == 0) tma_load_3d(st + kPlaneOff[k] * 4, &map_src, j0, il0 - kEx[k], k, &full[s]);
                    else tma_load_3d(st + kPlaneOff[k] * 4, &map_srch, j0 - kHaloY, il0 - kEx[k], k, &full[s]);
                }
                tma_load_2d(st + kCodeOff, &map_code, j0, il0, &full[s]);
            }
        }
        return;
    }
    if (warp == 1) {
        // ===================== store warp: one TMA store per finished output tile ==================
        if (lane == 0) {
            int it = 0;
            for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x, ++it) {
                const int o = it % kOutStages;
                const uint32_t ph = (it / kOutStages) & 1;
                const int sx = (t / a.n_ty) * kTileBX, j0 = (t % a.n_ty) * kTileBY;
                const float *out = out_base + o * (kStageOutBytes / 4);
                mbar_wait(&ofull[o], ph);
                tma_store_3d(&map_dst, j0, sx, 0, out);                       // box (BY, BX, 9): all planes at once
                if (EMIT) tma_store_3d(&map_mac, j0, sx, 0, out + 9 * kTileCells);  // box (BY, BX, 3)
                tma_commit();
                tma_wait_read<0>();   // shared memory of this stage has been read: consumers may refill it
                mbar_arrive(&oempty[o]);
            }
            tma_wait_all<0>();        // every store of this CTA has landed before the grid ends
        }
        return;
    }

    // ========================= consumers: collide from smem, write the output tile ===============
    const int ctid = threadIdx.x - 64;  // 0 .. kConsumers-1
    const int ny = a.ny;
    constexpr int kIter = kTileCells / kConsumers;   // cells per thread per tile (2)
    static_assert(kTileCells % kConsumers == 0 && kConsumers % kTileBY == 0, "tile / thread mapping");
    const int y = ctid % kTileBY, xb = ctid / kTileBY;          // this thread's row is the same in every tile
    float vmax = 0.0f;
    int vnan = 0;
    const float ramp = __ldg(a.ramp_tab + min(*a.ctr_in + 1, a.warmup));
    int it = 0;
    for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x, ++it) {
        const int s = it % kInStages;
        const uint32_t ph = (it / kInStages) & 1;
        const int o = it % kOutStages;
        const uint32_t oph = (it / kOutStages) & 1;
        const int il0 = a.col_lo + (t / a.n_ty) * kTileBX, j0 = (t % a.n_ty) * kTileBY;
        const float *in = reinterpret_cast<const float *>(in_base + s * kStageInStride);
        const unsigned char *codes = in_base + s * kStageInStride + kCodeOff;
        float *out = out_base + o * (kStageOutBytes / 4);
        const int j = j0 + y;

        // sponge damping of this thread's cells (ref:364-380), fetched before waiting on the tile
        float dmp[kIter];
        {
            const float dy = (j < ny) ? __ldg(a.damp_y + j) : 0.0f;
#pragma unroll
            for (int i = 0; i < kIter; ++i) {
                const int il = il0 + xb + i * (kConsumers / kTileBY);
                dmp[i] = fmaxf((il < a.nx_local) ? __ldg(a.damp_x + il) : 0.0f, dy);
            }
        }
        TileSink sink;
        sink.sm_f = out;
        sink.sm_mac = out + 9 * kTileCells;
        sink.il0 = il0; sink.j0 = j0; sink.bx = kTileBX; sink.by = kTileBY;
        sink.row_hi = a.row_hi; sink.col_lo = a.col_lo; sink.col_hi = a.col_hi;

        mbar_wait(&full[s], ph);          // the tile's shifted planes have landed
        mbar_wait(&oempty[o], oph ^ 1);   // the output stage is no longer being read by an earlier store

        float fin[kIter][9], g[kIter][9], rho[kIter], ux[kIter], uy[kIter];
        bool interior[kIter];
#pragma unroll
        for (int i = 0; i < kIter; ++i) {
            const int x = xb + i * (kConsumers / kTileBY);
            const int il = il0 + x;
            interior[i] = (il >= 1) && (il <= a.nx_local - 2) && (j >= 1) && (j <= ny - 2);
#pragma unroll
            for (int k = 0; k < 9; ++k)
                fin[i][k] = in[kPlaneOff[k] + x * kPlaneRow[k] + y + (kEy[k] == 0 ? 0 : kHaloY - kEy[k])];
        }
#pragma unroll
        for (int i = 0; i < kIter; ++i) {
            if (STRICT) collide_strict(a.phys, fin[i], dmp[i], g[i]);
            else collide_fast(a.phys, fin[i], dmp[i], g[i]);
            const int x = xb + i * (kConsumers / kTileBY);
            const int il = il0 + x;
            const bool owner = (j == 1) || (j == ny - 2) || (il == 1 && a.west_ring) || (il == a.nx_local - 2 && a.east_ring);
            rho[i] = ux[i] = uy[i] = 0.0f;
            if (EMIT || owner || (codes[x * kTileBY + y] & 1)) macro_from_f<STRICT>(g[i], rho[i], ux[i], uy[i]);
        }
#pragma unroll
        for (int i = 0; i < kIter; ++i) {
            if (!interior[i]) continue;   // ring cells are written by their owners, the rest is clipped by the store
            const int x = xb + i * (kConsumers / kTileBY);
            const int il = il0 + x, c = x * kTileBY + y;
            const bool owner = (j == 1) || (j == ny - 2) || (il == 1 && a.west_ring) || (il == a.nx_local - 2 && a.east_ring);
            if (owner) {  // rare: produce the ring cells hanging off this cell from its un-refilled state
                Cell me;
#pragma unroll
                for (int k = 0; k < 9; ++k) me.f[k] = g[i][k];
                me.rho = rho[i]; me.ux = ux[i]; me.uy = uy[i];
                ring_from_owner(a.ring, &sink, EMIT, il, j, &me, ramp, &vmax, &vnan);
            }
            if (codes[c] & 1) {  // obstacle refill, ref:452-455
                ux[i] = 0.0f; uy[i] = 0.0f;
#pragma unroll
                for (int k = 0; k < 9; ++k) g[i][k] = __fmul_rn(kW[k], rho[i]);
            }
#pragma unroll
            for (int k = 0; k < 9; ++k) out[k * kTileCells + c] = g[i][k];
            if (EMIT) {
                out[9 * kTileCells + c] = rho[i];
                out[10 * kTileCells + c] = ux[i];
                out[11 * kTileCells + c] = uy[i];
                const float m2 = vmag2_strict(ux[i], uy[i]);
                vnan |= (m2 != m2);
                vmax = fmaxf(vmax, m2);
            }
        }
        // hand the input stage back to the producer and the output tile to the store warp: the writes
        // above go through the generic proxy, TMA reads through the async proxy -> fence, then arrive
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&empty[s]);
            mbar_arrive(&ofull[o]);
        }
    }

    if (EMIT) {
        for (int sft = 16; sft > 0; sft >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, sft));
        const bool any_nan = __any_sync(0xffffffffu, vnan != 0);
        if (lane == 0) {
            const unsigned bits = __float_as_uint(vmax);
            if (bits > *a.maxv_bits) atomicMax(a.maxv_bits, bits);
            if (any_nan) a.maxv_bits[1] = 1u;
        }
    }
}

}  // namespace lbm
