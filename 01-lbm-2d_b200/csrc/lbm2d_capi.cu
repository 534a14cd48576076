// C ABI of the B200-native D2Q9 MRT-LES step (see include/lbm2d.h).  Host side: owns the device
// buffers, derives the fp32 constants the way the reference's Taichi program does, launches the
// kernels of lbm2d_kernels.cuh.  No CPU fallback: every path below needs a CUDA device.
#include "../../include/lbm2d.h"

#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <sys/mman.h>
#include <thread>
#include <unordered_map>
#include <vector>

#include "lbm2d_export.cuh"
#include "lbm2d_tma.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return fail(LBM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));         \
    } while (0)

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace

// ---- NCCL, resolved at run time from the library torch has already loaded (no link-time dependency) ----
namespace nccl {
typedef struct { char internal[128]; } UniqueId;
typedef void *Comm;
enum { kInt32 = 2, kFloat32 = 7 };
struct Api {
    int (*GetUniqueId)(UniqueId *) = nullptr;
    int (*CommInitRank)(Comm *, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(Comm) = nullptr;
    int (*Send)(const void *, size_t, int, int, Comm, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, Comm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};
inline Api &api() {
    static Api a;
    static bool tried = false;
    if (tried) return a;
    tried = true;
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return a;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(lib, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(lib, "ncclCommDestroy");
    a.Send = (decltype(a.Send))dlsym(lib, "ncclSend");
    a.Recv = (decltype(a.Recv))dlsym(lib, "ncclRecv");
    a.GroupStart = (decltype(a.GroupStart))dlsym(lib, "ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))dlsym(lib, "ncclGroupEnd");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(lib, "ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.Send && a.Recv && a.GroupStart && a.GroupEnd;
    return a;
}
}  // namespace nccl

// Device memory of destroyed single-GPU handles is kept for the next handle of the process.  A dataset sweep creates and
// destroys a solver per case, several at a time on one GPU (batch.py): ~30 cudaMalloc and, worse, ~30 cudaFree per case --
// and cudaFree synchronises the whole device, i.e. it stalls the steps of the OTHER cases in flight.  Blocks are recycled
// whole (never sub-allocated), zeroed on reuse (what fresh cudaMalloc memory is in practice), capped in size and total;
// x-slab handles (their buffers are exported through CUDA IPC and may be large) keep plain cudaMalloc / cudaFree.
namespace devpool {
constexpr size_t kMaxBlock = (size_t)64 << 20, kMaxCached = (size_t)1 << 30;
struct Pool {
    std::mutex mu;
    std::multimap<std::pair<int, size_t>, void *> free_blocks;
    std::unordered_map<void *, std::pair<int, size_t>> live;
    size_t cached = 0;
};
inline Pool &pool() {
    static Pool *p = new Pool;   // never destroyed: the CUDA context may be gone by the time static destructors run
    return *p;
}
inline bool enabled() {
    static const bool on = !std::getenv("LBM2D_NO_POOL");
    return on;
}
inline cudaError_t alloc(void **ptr, size_t bytes, bool pooled) {
    const size_t sz = (std::max<size_t>(bytes, 1) + 255) / 256 * 256;
    if (!pooled || !enabled() || sz > kMaxBlock) return cudaMalloc(ptr, bytes);
    int dev = 0;
    cudaGetDevice(&dev);
    Pool &p = pool();
    void *q = nullptr;
    {
        std::lock_guard<std::mutex> lk(p.mu);
        auto it = p.free_blocks.find({dev, sz});
        if (it != p.free_blocks.end()) {
            q = it->second;
            p.free_blocks.erase(it);
            p.cached -= sz;
        }
    }
    if (q) {   // its previous owner synchronised its stream before giving it back (lbm_destroy)
        cudaError_t e = cudaMemsetAsync(q, 0, sz, cudaStreamLegacy);
        if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamLegacy);   // the handle's own stream is non-blocking: order by hand
        if (e != cudaSuccess) return e;
    } else {
        cudaError_t e = cudaMalloc(&q, sz);
        if (e != cudaSuccess) return e;
    }
    {
        std::lock_guard<std::mutex> lk(p.mu);
        p.live[q] = {dev, sz};
    }
    *ptr = q;
    return cudaSuccess;
}
inline void release(void *ptr) {
    if (!ptr) return;
    Pool &p = pool();
    {
        std::lock_guard<std::mutex> lk(p.mu);
        auto it = p.live.find(ptr);
        if (it != p.live.end()) {
            const auto key = it->second;
            p.live.erase(it);
            if (p.cached + key.second <= kMaxCached) {
                p.free_blocks.emplace(key, ptr);
                p.cached += key.second;
                return;
            }
        }
    }
    cudaFree(ptr);
}
}  // namespace devpool

struct LbmSolver {
    LbmParams p{};
    nccl::Comm comm = nullptr;
    int rank = 0, nranks = 1;
    // peer-memory halo path (lbm_peer_connect): neighbour buffers mapped through CUDA IPC
    bool peer_mode = false;
    bool halo_wait_pending = false;
    void *peer_base[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};   // [side][f0, f1, inbox] as opened
    int peer_nx_local[2] = {0, 0};
    unsigned long long *inbox = nullptr;      // [0] from the west neighbour, [1] from the east neighbour, [2] error flag (low 32 bits)
    int64_t steps_total = 0;                  // steps since creation (never reset: the inbox counters are cumulative)
    cudaStream_t stream_e = nullptr;   // edge columns + halo exchange, overlapped with the interior
    cudaEvent_t ev_m = nullptr, ev_e = nullptr, ev_e_prev = nullptr, ev_x = nullptr;
    bool ev_e_prev_valid = false;
    int device = 0;
    cudaStream_t stream = nullptr;
    int nx_local = 0, ny = 0, pitch = 0, nseg = 0, n_items = 0;
    int own0 = 0;  // first owned local column
    int x_off = 0;
    bool west_ring = true, east_ring = true;
    long long plane = 0;
    float *f[2] = {nullptr, nullptr};
    uint8_t *code = nullptr;
    uint32_t *code_bits = nullptr;
    uint8_t *links8 = nullptr;            // bounce-back mode only
    float *damp_x = nullptr, *damp_y = nullptr, *ramp_tab = nullptr;
    int *ctr = nullptr;
    std::vector<float> ramp_host;          // ramp_at(t), t = 0 .. warmup_steps (ref:442-443)
    float *mac = nullptr;  // rho | ux | uy, three consecutive planes (one TMA store tensor)
    float *rho = nullptr, *ux = nullptr, *uy = nullptr;
    unsigned *maxv = nullptr;
    lbm::RingCtx *ring_ctx = nullptr;  // [2], one per destination buffer
    bool use_tma = false;
    bool use_pdl = true;
    long long early_min_ctas = 2500;      // grids with fewer CTAs keep the plain PDL hand-over
    int early_target = 1500;              // CTAs that may start on the progress counter (0 = early start off)
    // Launch-bound grids: a whole lbm_run(steps) batch is replayed as ONE CUDA graph (see lbm_run); key = steps * 2 + parity
    bool use_graph = true;
    int graph_min_steps = 8;
    std::map<long long, cudaGraphExec_t> graphs;
    int64_t graph_replays = 0;
    bool pooled = false;                  // device memory through devpool (single-GPU handles)
    template <class T> cudaError_t dalloc(T **ptr, size_t bytes) { return devpool::alloc((void **)ptr, bytes, pooled); }
    unsigned long long *progress = nullptr;   // device counter, see step_kernel
    unsigned long long progress_total = 0;    // its value once every step launched so far has signalled
    int tma_grid = 0;
    CUtensorMap map_src[2], map_srch[2], map_dst[2], map_code, map_mac;
    lbm::TmaArgs tma_args{};
    // export reduction state (lbm_export_*)
    bool exp_ready = false;
    lbm::ExportGeom exp_geom{};
    lbm::AreaEntry *exp_xtab = nullptr, *exp_ytab = nullptr;
    int *exp_xoff = nullptr, *exp_yoff = nullptr;
    float *exp_tmp = nullptr, *exp_frame = nullptr;
    double *exp_sum = nullptr, *exp_velsq = nullptr, *exp_vor = nullptr, *exp_minmax = nullptr;
    float *exp_halo = nullptr;      // [4][3*th]: send-west, send-east, recv-from-west (left), recv-from-east (right)
    int *exp_ecount = nullptr;      // [2] device scratch for the one-off extension-width handshake
    int exp_send_cols = 0, exp_recv_cols = 0;   // ROI columns sent to the west / received from the east neighbour
    int64_t exp_count = 0;
    lbm::Link *links = nullptr;
    int n_links = 0;
    double *force_partial = nullptr;
    float *force_out = nullptr;
    float *staging = nullptr;
    size_t staging_floats = 0;
    char *pinned[2] = {nullptr, nullptr};   // host staging of the large device -> host getters
    cudaEvent_t pin_ev[2] = {nullptr, nullptr};
    int64_t steps_done = 0;
    int64_t launches = 0;
    bool inited = false;
    lbm::Physics phys{};

    ~LbmSolver() {
        cudaSetDevice(device);
        for (int side = 0; side < 2; ++side)
            for (int i = 0; i < 3; ++i)
                if (peer_base[side][i]) cudaIpcCloseMemHandle(peer_base[side][i]);
        if (comm) nccl::api().CommDestroy(comm);
        for (cudaEvent_t ev : {ev_m, ev_e, ev_e_prev, ev_x})
            if (ev) cudaEventDestroy(ev);
        if (stream_e) cudaStreamDestroy(stream_e);
        for (auto &kv : graphs) cudaGraphExecDestroy(kv.second);
        for (void *ptr : {(void *)f[0], (void *)f[1], (void *)code, (void *)damp_x, (void *)damp_y, (void *)ramp_tab,
                          (void *)ctr, (void *)mac, (void *)ring_ctx, (void *)maxv, (void *)links,
                          (void *)force_partial, (void *)force_out, (void *)staging, (void *)exp_xtab, (void *)exp_ytab,
                          (void *)exp_xoff, (void *)exp_yoff, (void *)exp_tmp, (void *)exp_frame, (void *)exp_sum,
                          (void *)exp_velsq, (void *)exp_vor, (void *)exp_minmax, (void *)exp_halo, (void *)exp_ecount,
                          (void *)progress, (void *)code_bits, (void *)links8, (void *)inbox})
            if (ptr) devpool::release(ptr);
        for (int i = 0; i < 2; ++i) {
            if (pinned[i]) cudaFreeHost(pinned[i]);
            if (pin_ev[i]) cudaEventDestroy(pin_ev[i]);
        }
        if (stream) cudaStreamDestroy(stream);
    }
};

namespace {

constexpr int kForceBlocks = 128;

int ensure_staging(LbmSolver *s, size_t floats) {
    if (s->staging_floats >= floats) return LBM_OK;
    if (s->staging) devpool::release(s->staging);
    s->staging = nullptr;
    s->staging_floats = 0;
    CUDA_TRY(s->dalloc(&s->staging, floats * sizeof(float)));
    s->staging_floats = floats;
    return LBM_OK;
}

// Device -> caller-owned host array.  The reference hands out FRESH numpy arrays (they are queued to the writer
// thread), so the destination is pageable and untouched: a plain cudaMemcpy runs at ~5 GB/s there (page faults +
// the driver's own staging).  Large copies therefore go through two pinned chunks: the DMA of chunk i+1 overlaps
// a multi-threaded copy (and first touch) of chunk i into the caller's array.
constexpr size_t kPinChunk = 32u << 20;
int d2h(LbmSolver *s, void *host, const void *dev, size_t bytes) {
    bool pinned_dst = false;
    if (bytes >= 2 * kPinChunk) {   // destination from lbm_host_alloc (the binding's frame pool): plain DMA at PCIe speed
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, host) == cudaSuccess) pinned_dst = at.type == cudaMemoryTypeHost;
        else (void)cudaGetLastError();
    }
    if (bytes < 2 * kPinChunk || pinned_dst) {
        CUDA_TRY(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        return LBM_OK;
    }
    for (int i = 0; i < 2; ++i) {
        if (!s->pinned[i]) CUDA_TRY(cudaHostAlloc((void **)&s->pinned[i], kPinChunk, cudaHostAllocDefault));
        if (!s->pin_ev[i]) CUDA_TRY(cudaEventCreateWithFlags(&s->pin_ev[i], cudaEventDisableTiming));
    }
    {   // the destination is usually a fresh allocation: ask for huge pages so that first touch is ~500x fewer faults
        const uintptr_t huge = (uintptr_t)2 << 20;
        const uintptr_t lo = ((uintptr_t)host + huge - 1) & ~(huge - 1), hi = ((uintptr_t)host + bytes) & ~(huge - 1);
        if (hi > lo) (void)madvise((void *)lo, hi - lo, MADV_HUGEPAGE);
    }
    const unsigned nthreads = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    auto scatter = [&](const char *src, char *dst, size_t n) {
        std::vector<std::thread> pool;
        const size_t per = (n / nthreads + 4095) / 4096 * 4096;
        for (unsigned t = 0; t < nthreads; ++t) {
            const size_t lo = std::min(n, (size_t)t * per), hi = std::min(n, lo + per);
            if (hi > lo) pool.emplace_back([=] { std::memcpy(dst + lo, src + lo, hi - lo); });
        }
        for (auto &th : pool) th.join();
    };
    const size_t nchunks = (bytes + kPinChunk - 1) / kPinChunk;
    for (size_t i = 0; i <= nchunks; ++i) {
        if (i < nchunks) {
            const size_t off = i * kPinChunk, n = std::min(kPinChunk, bytes - off);
            CUDA_TRY(cudaMemcpyAsync(s->pinned[i & 1], (const char *)dev + off, n, cudaMemcpyDeviceToHost, s->stream));
            CUDA_TRY(cudaEventRecord(s->pin_ev[i & 1], s->stream));
        }
        if (i > 0) {
            const size_t j = i - 1, off = j * kPinChunk, n = std::min(kPinChunk, bytes - off);
            CUDA_TRY(cudaEventSynchronize(s->pin_ev[j & 1]));
            scatter(s->pinned[j & 1], (char *)host + off, n);
        }
    }
    return LBM_OK;
}

// The reference's sponge profile (ref:364-378), evaluated in fp32 exactly as the kernel would.
float sponge_1d(int i, int n, int w_lo, int w_hi, float strength, bool lo_first) {
    // x: `if i > n - w_hi ... elif i < w_lo`;  y: `if j < w_lo ... elif j > n - w_hi`
    auto hi = [&]() { float c = (float)(i - (n - w_hi)) / (float)w_hi; return strength * (c * c); };
    auto lo = [&]() { float c = (float)(w_lo - i) / (float)w_lo; return strength * (c * c); };
    if (lo_first) {
        if (i < w_lo) return lo();
        if (i > n - w_hi) return hi();
    } else {
        if (i > n - w_hi) return hi();
        if (i < w_lo) return lo();
    }
    return 0.0f;
}

// Cosine soft start (ref:442-443) for frame_count = t; the cosine is the correctly rounded fp32 of
// the double cosine (see oracle/lbm_oracle_np.py).
float ramp_at(int t, int warmup) {
    float progress = (warmup == 0) ? 1.0f : std::fmin(1.0f, (float)t / (float)warmup);
    const float arg = (float)(0.5 * 3.14159265) * progress;
    const float c = (float)std::cos((double)arg);
    return 1.0f - c;
}

lbm::StepArgs make_args(const LbmSolver *s, int par_override = -1) {
    lbm::StepArgs a{};
    const int par = par_override >= 0 ? par_override : (int)(s->steps_done & 1);
    static const int ex[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
    a.src = s->f[par];
    a.dst = s->f[par ^ 1];
    for (int k = 0; k < 9; ++k) {
        a.srcp[k] = a.src + k * s->plane - (long long)ex[k] * s->pitch;
        a.dstp[k] = a.dst + k * s->plane;
    }
    a.code = s->code;
    a.code_bits = s->code_bits;
    a.links8 = s->links8;
    a.damp_x = s->damp_x;
    a.damp_y = s->damp_y;
    a.ramp = s->ramp_host[std::min<int64_t>(s->steps_done + 1, s->p.warmup_steps)];
    a.rho = s->rho;
    a.ux = s->ux;
    a.uy = s->uy;
    a.maxv_bits = s->maxv;
    a.plane = s->plane;
    a.nx_local = s->nx_local;
    a.ny = s->ny;
    a.pitch = s->pitch;
    a.nseg = s->nseg;
    a.x_off = s->x_off;
    a.west_ring = s->west_ring;
    a.east_ring = s->east_ring;
    a.ring = s->ring_ctx + (par ^ 1);
    a.il0 = 1;
    a.il_step = 1;
    a.il_count = s->nx_local - 2;
    a.n_ring = lbm::ring_cell_count(a.il0, a.il_step, a.il_count, s->nx_local, s->ny, a.west_ring, a.east_ring);
    a.progress = s->progress;
    a.col_split = -1;
    a.edge_il[0] = a.edge_il[1] = -1;
    a.edge_row[0] = a.edge_row[1] = -1;
    a.phys = s->phys;
    if (s->peer_mode) {
        // grid order [1 .. E | W/E ring block | nx_local - 2 | E + 1 .. nx_local - 3] (set by lbm_run through col_split)
        static const int halo_plane[2][3] = {{3, 6, 7}, {1, 5, 8}};
        for (int side = 0; side < 2; ++side) {
            if (!s->peer_base[side][0]) continue;
            a.edge_il[side] = side == 0 ? 1 : s->nx_local - 2;
            const int nb_nx = s->peer_nx_local[side];
            const long long nb_plane = (long long)nb_nx * s->pitch;
            float *nb_dst = (float *)s->peer_base[side][par ^ 1];
            const long long halo_col = side == 0 ? (long long)(nb_nx - 1) * s->pitch : 0;   // west neighbour's EAST halo / east neighbour's WEST halo
            for (int q = 0; q < 3; ++q) a.peer_dst[side][q] = nb_dst + halo_plane[side][q] * nb_plane + halo_col;
            a.peer_inbox[side] = (unsigned long long *)s->peer_base[side][2] + (side == 0 ? 1 : 0);   // I am its east / west neighbour
        }
        a.inbox = s->inbox;
        a.peer_error = (unsigned *)(s->inbox + 2);
        const int gx = (s->nseg + lbm::kWarpsPerBlock - 1) / lbm::kWarpsPerBlock;
        a.inbox_expected = (unsigned long long)s->steps_total * (unsigned long long)(gx + 1);   // edge column CTAs + its ring row CTA, per step
    }
    return a;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_map(CUtensorMap *map, CUtensorMapDataType dt, int rank, void *base, const cuuint64_t *dims,
               const cuuint64_t *strides_bytes, const cuuint32_t *box) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !sym)
            return fail(LBM_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
        fn = (EncodeTiledFn)sym;
    }
    const cuuint32_t ones[3] = {1, 1, 1};
    CUresult r = fn(map, dt, (cuuint32_t)rank, base, dims, strides_bytes, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(LBM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return LBM_OK;
}

// Tensor maps + static arguments of the TMA variant (see lbm2d_tma.cuh).
int setup_tma(LbmSolver *s) {
    using namespace lbm;
    const int west_halo = s->west_ring ? 0 : 1, east_halo = s->east_ring ? 0 : 1;
    int col_lo = west_halo, col_hi = s->nx_local - east_halo, row_hi = s->ny;
    // a ring row / column that would start a tile of its own is written by its owners with scalar stores
    if ((s->ny - 1) % kTileBY == 0) row_hi = s->ny - 1;
    if (s->east_ring && (s->nx_local - 1 - col_lo) % kTileBX == 0) col_hi = s->nx_local - 1;
    TmaArgs &a = s->tma_args;
    a.damp_x = s->damp_x;
    a.damp_y = s->damp_y;
    a.ramp_tab = s->ramp_tab;
    a.maxv_bits = s->maxv;
    a.nx_local = s->nx_local;
    a.ny = s->ny;
    a.pitch = s->pitch;
    a.col_lo = col_lo;
    a.col_hi = col_hi;
    a.row_hi = row_hi;
    a.n_tx = (col_hi - col_lo + kTileBX - 1) / kTileBX;
    a.n_ty = (row_hi + kTileBY - 1) / kTileBY;
    a.n_tiles = a.n_tx * a.n_ty;
    a.west_ring = s->west_ring;
    a.east_ring = s->east_ring;
    a.warmup = s->p.warmup_steps;
    a.phys = s->phys;

    const cuuint64_t pitch_b = (cuuint64_t)s->pitch * 4, plane_b = (cuuint64_t)s->plane * 4;
    const cuuint32_t box3[3] = {(cuuint32_t)kTileBY, (cuuint32_t)kTileBX, 1};
    for (int b = 0; b < 2; ++b) {
        const cuuint64_t dsrc[3] = {(cuuint64_t)s->pitch, (cuuint64_t)s->nx_local, 9};
        const cuuint64_t st[2] = {pitch_b, plane_b};
        if (int rc = encode_map(&s->map_src[b], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, s->f[b], dsrc, st, box3)) return rc;
        const cuuint32_t box3h[3] = {(cuuint32_t)kRowHalo, (cuuint32_t)kTileBX, 1};
        if (int rc = encode_map(&s->map_srch[b], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, s->f[b], dsrc, st, box3h)) return rc;
        const cuuint64_t ddst[3] = {(cuuint64_t)row_hi, (cuuint64_t)(col_hi - col_lo), 9};
        const cuuint32_t box9[3] = {(cuuint32_t)kTileBY, (cuuint32_t)kTileBX, 9};
        if (int rc = encode_map(&s->map_dst[b], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, s->f[b] + (size_t)col_lo * s->pitch, ddst, st, box9))
            return rc;
    }
    {
        const cuuint64_t d[3] = {(cuuint64_t)row_hi, (cuuint64_t)(col_hi - col_lo), 3};
        const cuuint64_t st[2] = {pitch_b, plane_b};
        const cuuint32_t box3m[3] = {(cuuint32_t)kTileBY, (cuuint32_t)kTileBX, 3};
        if (int rc = encode_map(&s->map_mac, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, s->mac + (size_t)col_lo * s->pitch, d, st, box3m)) return rc;
        const cuuint64_t dc[2] = {(cuuint64_t)s->pitch, (cuuint64_t)s->nx_local};
        const cuuint64_t stc[1] = {(cuuint64_t)s->pitch};
        const cuuint32_t box2[2] = {(cuuint32_t)kTileBY, (cuuint32_t)kTileBX};
        if (int rc = encode_map(&s->map_code, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, s->code, dc, stc, box2)) return rc;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device);
    s->tma_grid = std::min(a.n_tiles, sms);
    for (auto fnp : {(const void *)step_tma_kernel<false, false>, (const void *)step_tma_kernel<false, true>,
                     (const void *)step_tma_kernel<true, false>, (const void *)step_tma_kernel<true, true>})
        CUDA_TRY(cudaFuncSetAttribute(fnp, cudaFuncAttributeMaxDynamicSharedMemorySize, kTmaSmemBytes));
    return LBM_OK;
}

lbm::ExportArgs make_export_args(const LbmSolver *s) {
    lbm::ExportArgs a{};
    const int par = (int)(s->steps_done & 1);
    a.cur = s->f[par];
    a.prev = s->f[par ^ 1];
    a.code = s->code;
    a.damp_x = s->damp_x;
    a.damp_y = s->damp_y;
    a.plane = s->plane;
    a.nx_local = s->nx_local;
    a.ny = s->ny;
    a.pitch = s->pitch;
    a.il0 = s->own0;
    a.il1 = s->own0 + s->p.nx;
    a.x_off = s->x_off;
    a.nx_global = s->p.nx_global;
    a.have_prev = s->steps_done > 0;
    a.strict = s->p.arith == LBM_ARITH_STRICT;
    a.phys = s->phys;
    return a;
}

#define NCCL_TRY(expr)                                                                              \
    do {                                                                                            \
        int r__ = (expr);                                                                           \
        if (r__ != 0) {                                                                             \
            const char *m__ = nccl::api().GetErrorString ? nccl::api().GetErrorString(r__) : "?";   \
            return fail(LBM_ERR_NCCL, std::string(#expr) + ": " + m__);                             \
        }                                                                                           \
    } while (0)

// One halo column per interface: the populations that stream across it.  `buf` is the
// buffer the step just wrote; sends read the first / last OWNED column, receives fill the halo columns.
int exchange_halos(LbmSolver *s, float *buf, cudaStream_t st) {
    if (!s->comm || s->nranks == 1) return LBM_OK;
    nccl::Api &n = nccl::api();
    static const int east_going[3] = {1, 5, 8}, west_going[3] = {3, 6, 7};
    const size_t cnt = (size_t)s->pitch;
    const long long pl = s->plane;
    NCCL_TRY(n.GroupStart());
    if (!s->east_ring) {  // east neighbour = rank + 1
        for (int q = 0; q < 3; ++q) {
            NCCL_TRY(n.Send(buf + east_going[q] * pl + (long long)(s->nx_local - 2) * s->pitch, cnt, nccl::kFloat32, s->rank + 1, s->comm, st));
            NCCL_TRY(n.Recv(buf + west_going[q] * pl + (long long)(s->nx_local - 1) * s->pitch, cnt, nccl::kFloat32, s->rank + 1, s->comm, st));
        }
    }
    if (!s->west_ring) {  // west neighbour = rank - 1
        for (int q = 0; q < 3; ++q) {
            NCCL_TRY(n.Send(buf + west_going[q] * pl + (long long)s->pitch, cnt, nccl::kFloat32, s->rank - 1, s->comm, st));
            NCCL_TRY(n.Recv(buf + east_going[q] * pl, cnt, nccl::kFloat32, s->rank - 1, s->comm, st));
        }
    }
    NCCL_TRY(n.GroupEnd());
    return LBM_OK;
}

typedef void (*StepFn)(const lbm::StepArgs);
StepFn step_fn(bool strict, bool emit, bool bb = false, bool peer = false) {
#define LBM_PICK2(S, E, B) (peer ? (StepFn)lbm::step_kernel<S, E, B, true> : (StepFn)lbm::step_kernel<S, E, B, false>)
#define LBM_PICK(S, E) (bb ? LBM_PICK2(S, E, true) : LBM_PICK2(S, E, false))
    return strict ? (emit ? LBM_PICK(true, true) : LBM_PICK(true, false)) : (emit ? LBM_PICK(false, true) : LBM_PICK(false, false));
#undef LBM_PICK
#undef LBM_PICK2
}

// Launch one step of the register variant.  `pdl`: programmatic dependent launch -- the grid may start being
// scheduled before the previous kernel in the stream has drained (it synchronises on it itself, see the kernel).
cudaError_t launch_step(StepFn fn, dim3 grid, cudaStream_t st, const lbm::StepArgs &a, bool pdl) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(lbm::kThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, fn, a);
}

int check_handle(LbmHandle h, bool need_init) {
    if (!h) return fail(LBM_ERR_INVALID, "null handle");
    if (need_init && !h->inited) return fail(LBM_ERR_STATE, "lbm_init() has not been called");
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) return fail(LBM_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    return LBM_OK;
}

}  // namespace

extern "C" {

int lbm_abi_version(void) { return LBM2D_ABI_VERSION; }
const char *lbm_last_error(void) { return g_err.c_str(); }

int lbm_device_count(int *count) {
    if (!count) return fail(LBM_ERR_INVALID, "count is null");
    CUDA_TRY(cudaGetDeviceCount(count));
    return LBM_OK;
}

int lbm_create(const LbmParams *params, const uint8_t *mask_xy, LbmHandle *out) {
    if (!params || !out) return fail(LBM_ERR_INVALID, "params / out is null");
    const LbmParams &p = *params;
    if (p.nx < 1 || p.ny < 3) return fail(LBM_ERR_INVALID, "need nx >= 1 (owned) and ny >= 3");
    if (p.nx_global < 3) return fail(LBM_ERR_INVALID, "need nx_global >= 3");
    if (p.slab_x0 < 0 || p.slab_x0 + p.nx > p.nx_global) return fail(LBM_ERR_INVALID, "slab outside the global domain");
    if (p.warmup_steps < 0) return fail(LBM_ERR_INVALID, "warmup_steps < 0");
    if (p.obstacle_mode != LBM_OBSTACLE_REFILL && p.obstacle_mode != LBM_OBSTACLE_BOUNCE_BACK)
        return fail(LBM_ERR_INVALID, "unsupported obstacle_mode");
    if (p.obstacle_mode == LBM_OBSTACLE_BOUNCE_BACK && p.kernel != LBM_KERNEL_AUTO && p.kernel != LBM_KERNEL_REGISTER)
        return fail(LBM_ERR_INVALID, "obstacle_mode bounce-back: the default (register) kernel only");
    if (p.arith != LBM_ARITH_FAST && p.arith != LBM_ARITH_STRICT) return fail(LBM_ERR_INVALID, "unsupported arith");
    if (p.warmup_steps > (1 << 26)) return fail(LBM_ERR_INVALID, "warmup_steps too large for the ramp table");

    int ndev = 0;
    cudaError_t e0 = cudaGetDeviceCount(&ndev);
    if (e0 != cudaSuccess || ndev == 0)
        return fail(LBM_ERR_CUDA, std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e0));
    int dev = p.device;
    if (dev < 0) CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= ndev) return fail(LBM_ERR_INVALID, "device ordinal out of range");
    CUDA_TRY(cudaSetDevice(dev));

    LbmSolver *s = new (std::nothrow) LbmSolver();
    if (!s) return fail(LBM_ERR_INVALID, "out of host memory");
    s->p = p;
    s->device = dev;
    const bool west_halo = p.slab_x0 > 0, east_halo = p.slab_x0 + p.nx < p.nx_global;
    s->west_ring = !west_halo;
    s->east_ring = !east_halo;
    s->own0 = west_halo ? 1 : 0;
    s->x_off = p.slab_x0 - (west_halo ? 1 : 0);
    s->nx_local = p.nx + (west_halo ? 1 : 0) + (east_halo ? 1 : 0);
    if (s->nx_local < 3) {
        delete s;
        return fail(LBM_ERR_INVALID, "a slab needs at least 3 local columns");
    }
    s->ny = p.ny;
    s->pitch = round_up(p.ny, lbm::kSegCells);   // every lane of every segment warp stays inside its column
    s->plane = (long long)s->nx_local * s->pitch;
    s->nseg = s->pitch / lbm::kSegCells;
    s->n_items = (s->nx_local - 2) * s->nseg;

    // fp32 constants, derived like the reference's Python scope + Taichi f32 casts
    const double tau0 = 3.0 * p.nu + 0.5;                       // ref:44
    s->phys.tau0 = (float)tau0;
    s->phys.tau0_sq = (float)(tau0 * tau0);                     // ref:348 (python-scope power, then f32)
    s->phys.cs_factor = (float)(18.0 * (p.c_smag * p.c_smag));  // ref:79
    s->phys.s_ghost = (float)p.s_ghost;
    s->phys.les_on = p.c_smag > 0.001;                          // ref:342
    s->phys.rho_in = (float)p.rho_in;
    s->phys.rho_out = (float)p.rho_out;
    for (int d = 0; d < 4; ++d) {
        s->phys.bc_type[d] = p.bc_type[d];
        s->phys.bc_val[d][0] = p.bc_value[d][0];
        s->phys.bc_val[d][1] = p.bc_value[d][1];
    }
    s->phys.nx_global = p.nx_global;
    {   // where the packed division / square-root sequences of the strict collision need no range checks (Lane2)
        auto in = [](double v, double lo, double hi) { return v >= lo && v <= hi; };
        const double strength = (float)p.sponge_strength;
        s->phys.fast_div_ok = in(s->phys.tau0, 0x1p-10, 0x1p20) && in(s->phys.tau0_sq, 0x1p-20, 0x1p40) &&
                              (!s->phys.les_on || in(s->phys.cs_factor, 0x1p-20, 0x1p20)) && in(strength, 0.0, 0x1p30) &&
                              !std::getenv("LBM2D_NO_FAST_DIV");
    }

#define CREATE_TRY(expr)                                                                            \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            delete s;                                                                               \
            return fail(LBM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));         \
        }                                                                                           \
    } while (0)

    CREATE_TRY(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    s->pooled = p.nx == p.nx_global;   // x-slabs export their buffers through CUDA IPC: plain allocations
    const size_t fbytes = ((size_t)9 * s->plane + 64) * sizeof(float);  // +64: the last segment may prefetch past the end
    CREATE_TRY(s->dalloc(&s->f[0], fbytes));
    CREATE_TRY(s->dalloc(&s->f[1], fbytes));
    CREATE_TRY(cudaMemset(s->f[0], 0, fbytes));
    CREATE_TRY(cudaMemset(s->f[1], 0, fbytes));
    CREATE_TRY(s->dalloc(&s->mac, (size_t)3 * s->plane * sizeof(float)));
    s->rho = s->mac;
    s->ux = s->mac + s->plane;
    s->uy = s->mac + 2 * s->plane;
    CREATE_TRY(s->dalloc(&s->ring_ctx, 2 * sizeof(lbm::RingCtx)));
    CREATE_TRY(s->dalloc(&s->code, (size_t)s->plane + 64));
    CREATE_TRY(s->dalloc(&s->damp_x, s->nx_local * sizeof(float)));
    CREATE_TRY(s->dalloc(&s->damp_y, s->pitch * sizeof(float)));
    CREATE_TRY(s->dalloc(&s->ramp_tab, ((size_t)p.warmup_steps + 1) * sizeof(float)));
    CREATE_TRY(s->dalloc(&s->ctr, 2 * sizeof(int)));
    CREATE_TRY(s->dalloc(&s->progress, sizeof(unsigned long long)));
    CREATE_TRY(cudaMemset(s->progress, 0, sizeof(unsigned long long)));
    CREATE_TRY(s->dalloc(&s->maxv, 2 * sizeof(unsigned)));
    CREATE_TRY(s->dalloc(&s->inbox, 4 * sizeof(unsigned long long)));
    CREATE_TRY(cudaMemset(s->inbox, 0, 4 * sizeof(unsigned long long)));
    CREATE_TRY(s->dalloc(&s->force_partial, kForceBlocks * 2 * sizeof(double)));
    CREATE_TRY(s->dalloc(&s->force_out, 2 * sizeof(float)));

    // cell codes (bit0 = solid), padded to the pitch
    std::vector<uint8_t> code((size_t)s->plane + 64, 0);
    if (mask_xy)
        for (int il = 0; il < s->nx_local; ++il)
            for (int j = 0; j < s->ny; ++j) code[(size_t)il * s->pitch + j] = mask_xy[(size_t)il * s->ny + j] ? 1 : 0;
    CREATE_TRY(cudaMemcpy(s->code, code.data(), code.size(), cudaMemcpyHostToDevice));
    if (p.obstacle_mode == LBM_OBSTACLE_BOUNCE_BACK) {   // solid upstream neighbours of every fluid cell
        static const int ex[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1}, ey[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
        std::vector<uint8_t> links((size_t)s->plane + 64, 0);
        for (int il = 1; il < s->nx_local - 1; ++il)
            for (int j = 1; j < s->ny - 1; ++j) {
                const size_t o = (size_t)il * s->pitch + j;
                if (code[o] & 1) continue;
                uint8_t l = 0;
                for (int k = 1; k < 9; ++k)
                    if (code[(size_t)(il - ex[k]) * s->pitch + (j - ey[k])] & 1) l |= (uint8_t)(1u << (k - 1));
                links[o] = l;
            }
        CREATE_TRY(s->dalloc(&s->links8, links.size()));
        CREATE_TRY(cudaMemcpy(s->links8, links.data(), links.size(), cudaMemcpyHostToDevice));
    }
    {   // bit-packed copy for the interior warps (1/8 of the bytes per step)
        std::vector<uint32_t> bits(((size_t)s->plane + 31) / 32 + 2, 0u);
        for (size_t o = 0; o < (size_t)s->plane; ++o)
            if (code[o] & 1) bits[o >> 5] |= 1u << (o & 31);
        CREATE_TRY(s->dalloc(&s->code_bits, bits.size() * sizeof(uint32_t)));
        CREATE_TRY(cudaMemcpy(s->code_bits, bits.data(), bits.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }

    // sponge tables (ref:90-94 widths are max(1, cfg))
    const int w_in = std::max(1, p.sponge_in), w_out = std::max(1, p.sponge_out);
    const int w_top = std::max(1, p.sponge_top), w_bot = std::max(1, p.sponge_bot);
    const float strength = (float)p.sponge_strength;
    std::vector<float> dx(s->nx_local), dy(s->pitch, 0.0f);
    for (int il = 0; il < s->nx_local; ++il) dx[il] = sponge_1d(s->x_off + il, p.nx_global, w_in, w_out, strength, false);
    for (int j = 0; j < s->ny; ++j) dy[j] = sponge_1d(j, s->ny, w_bot, w_top, strength, true);
    CREATE_TRY(cudaMemcpy(s->damp_x, dx.data(), dx.size() * sizeof(float), cudaMemcpyHostToDevice));
    CREATE_TRY(cudaMemcpy(s->damp_y, dy.data(), dy.size() * sizeof(float), cudaMemcpyHostToDevice));

    std::vector<float> &ramp = s->ramp_host;
    ramp.resize((size_t)p.warmup_steps + 1);
    for (int t = 0; t <= p.warmup_steps; ++t) ramp[t] = ramp_at(t, p.warmup_steps);
    CREATE_TRY(cudaMemcpy(s->ramp_tab, ramp.data(), ramp.size() * sizeof(float), cudaMemcpyHostToDevice));

    // solid-fluid links of the momentum-exchange force (ref:597-641), for solids in OWNED columns
    {
        static const int inv[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
        static const int ex[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1}, ey[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
        std::vector<lbm::Link> links;
        for (int il = s->own0; il < s->own0 + p.nx; ++il)
            for (int j = 0; j < s->ny; ++j) {
                if (!code[(size_t)il * s->pitch + j]) continue;
                for (int k = 1; k < 9; ++k) {
                    const int nl = il + ex[k], nj = j + ey[k], ng = s->x_off + nl;
                    if (ng < 0 || ng >= p.nx_global || nj < 0 || nj >= s->ny) continue;
                    if (code[(size_t)nl * s->pitch + nj]) continue;
                    const bool ring = ng == 0 || ng == p.nx_global - 1 || nj == 0 || nj == s->ny - 1;
                    lbm::Link l;
                    l.offset = (int)((long long)nl * s->pitch + nj);
                    l.packed = inv[k] | ((int)ring << 4) | ((-ex[k] + 1) << 5) | ((-ey[k] + 1) << 7);
                    links.push_back(l);
                }
            }
        s->n_links = (int)links.size();
        if (s->plane > 0x7fffffffLL) {
            delete s;
            return fail(LBM_ERR_INVALID, "slab too large for 32-bit link offsets");
        }
        if (s->n_links) {
            CREATE_TRY(s->dalloc(&s->links, links.size() * sizeof(lbm::Link)));
            CREATE_TRY(cudaMemcpy(s->links, links.data(), links.size() * sizeof(lbm::Link), cudaMemcpyHostToDevice));
        }
    }
    {
        lbm::RingCtx rc[2];
        for (int b = 0; b < 2; ++b) {
            rc[b].phys = s->phys;
            rc[b].dst = s->f[b];
            rc[b].rho = s->rho;
            rc[b].ux = s->ux;
            rc[b].uy = s->uy;
            rc[b].code = s->code;
            rc[b].plane = s->plane;
            rc[b].nx_local = s->nx_local;
            rc[b].ny = s->ny;
            rc[b].pitch = s->pitch;
            rc[b].x_off = s->x_off;
            rc[b].west_ring = s->west_ring;
            rc[b].east_ring = s->east_ring;
        }
        CREATE_TRY(cudaMemcpy(s->ring_ctx, rc, sizeof(rc), cudaMemcpyHostToDevice));
    }
#undef CREATE_TRY
    // kernel variant: the persistent TMA pipeline needs enough tiles to occupy every SM
    {
        const long long tiles = ((long long)s->nx_local + lbm::kTileBX - 1) / lbm::kTileBX * ((s->ny + lbm::kTileBY - 1) / lbm::kTileBY);
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        (void)tiles;
        s->use_tma = p.kernel == LBM_KERNEL_TMA;
        s->use_pdl = !std::getenv("LBM2D_NO_PDL");
        s->use_graph = !std::getenv("LBM2D_NO_GRAPH");
        if (const char *e = std::getenv("LBM2D_GRAPH_MIN_STEPS")) s->graph_min_steps = std::max(1, std::atoi(e));
        if (const char *e = std::getenv("LBM2D_EARLY_CTAS")) s->early_target = std::max(0, std::atoi(e));
        if (const char *e = std::getenv("LBM2D_EARLY_MIN_CTAS")) s->early_min_ctas = std::max(0, std::atoi(e));
        if (p.kernel < LBM_KERNEL_AUTO || p.kernel > LBM_KERNEL_TMA) {
            delete s;
            return fail(LBM_ERR_INVALID, "unsupported kernel variant");
        }
        if (s->use_tma)
            if (int rc = setup_tma(s)) {
                delete s;
                return rc;
            }
    }
    *out = s;
    return LBM_OK;
}

int lbm_comm_unique_id(uint8_t out[LBM_COMM_ID_BYTES]) {
    if (!out) return fail(LBM_ERR_INVALID, "out is null");
    nccl::Api &n = nccl::api();
    if (!n.ok) return fail(LBM_ERR_NCCL, "libnccl.so.2 could not be loaded");
    nccl::UniqueId id;
    NCCL_TRY(n.GetUniqueId(&id));
    std::memcpy(out, id.internal, LBM_COMM_ID_BYTES);
    return LBM_OK;
}

int lbm_comm_connect(LbmHandle h, int rank, int nranks, const uint8_t id_bytes[LBM_COMM_ID_BYTES]) {
    if (int rc = check_handle(h, false)) return rc;
    if (!id_bytes || nranks < 1 || rank < 0 || rank >= nranks) return fail(LBM_ERR_INVALID, "bad rank / nranks / id");
    if ((rank == 0) != h->west_ring || (rank == nranks - 1) != h->east_ring)
        return fail(LBM_ERR_INVALID, "slabs must be ordered west to east by rank and tile the global domain");
    nccl::Api &n = nccl::api();
    if (!n.ok) return fail(LBM_ERR_NCCL, "libnccl.so.2 could not be loaded");
    nccl::UniqueId id;
    std::memcpy(id.internal, id_bytes, LBM_COMM_ID_BYTES);
    NCCL_TRY(n.CommInitRank(&h->comm, nranks, id, rank));
    h->rank = rank;
    h->nranks = nranks;
    if (nranks > 1) {   // NCCL sets its point-to-point channels up lazily (~1-2 s): do it here, not inside the first export / step
        nccl::Api &na = nccl::api();
        int *scratch = nullptr;
        CUDA_TRY(cudaMalloc(&scratch, 4 * sizeof(int)));
        CUDA_TRY(cudaMemsetAsync(scratch, 0, 4 * sizeof(int), h->stream));
        NCCL_TRY(na.GroupStart());
        if (!h->east_ring) {
            NCCL_TRY(na.Send(scratch, 1, nccl::kInt32, rank + 1, h->comm, h->stream));
            NCCL_TRY(na.Recv(scratch + 1, 1, nccl::kInt32, rank + 1, h->comm, h->stream));
        }
        if (!h->west_ring) {
            NCCL_TRY(na.Send(scratch + 2, 1, nccl::kInt32, rank - 1, h->comm, h->stream));
            NCCL_TRY(na.Recv(scratch + 3, 1, nccl::kInt32, rank - 1, h->comm, h->stream));
        }
        NCCL_TRY(na.GroupEnd());
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        cudaFree(scratch);
    }
    if (nranks > 1 && !std::getenv("LBM2D_NO_OVERLAP")) {
        int lo = 0, hi = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CUDA_TRY(cudaStreamCreateWithPriority(&h->stream_e, cudaStreamNonBlocking, hi));  // edge + comm first
        for (cudaEvent_t *ev : {&h->ev_m, &h->ev_e, &h->ev_e_prev, &h->ev_x})
            CUDA_TRY(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
    }
    return LBM_OK;
}

// ---- peer-memory halo path: CUDA IPC handles of this slab's buffers, exchanged by the host (torch.distributed) --------
namespace {
struct PeerBlob {
    cudaIpcMemHandle_t f0, f1, inbox;
    int32_t nx_local, pitch, device, pad;
};
static_assert(sizeof(PeerBlob) <= LBM_PEER_HANDLE_BYTES, "blob size");
}  // namespace

int lbm_peer_export(LbmHandle h, uint8_t out[LBM_PEER_HANDLE_BYTES]) {
    if (int rc = check_handle(h, false)) return rc;
    if (!out) return fail(LBM_ERR_INVALID, "out is null");
    PeerBlob b{};
    CUDA_TRY(cudaIpcGetMemHandle(&b.f0, h->f[0]));
    CUDA_TRY(cudaIpcGetMemHandle(&b.f1, h->f[1]));
    CUDA_TRY(cudaIpcGetMemHandle(&b.inbox, h->inbox));
    b.nx_local = h->nx_local;
    b.pitch = h->pitch;
    b.device = h->device;
    std::memset(out, 0, LBM_PEER_HANDLE_BYTES);
    std::memcpy(out, &b, sizeof(b));
    return LBM_OK;
}

int lbm_peer_connect(LbmHandle h, const uint8_t *west, const uint8_t *east) {
    if (int rc = check_handle(h, false)) return rc;
    if ((west != nullptr) == h->west_ring || (east != nullptr) == h->east_ring)
        return fail(LBM_ERR_INVALID, "a neighbour blob is needed exactly on the sides that are halos");
    if (h->use_tma) return fail(LBM_ERR_INVALID, "peer-memory halos: register kernel only");
    const uint8_t *blobs[2] = {west, east};
    for (int side = 0; side < 2; ++side) {
        if (!blobs[side]) continue;
        PeerBlob b;
        std::memcpy(&b, blobs[side], sizeof(b));
        if (b.pitch != h->pitch || b.nx_local < 3) return fail(LBM_ERR_INVALID, "neighbour slab has a different ny / is too narrow");
        int can = 0;
        if (b.device != h->device) {
            CUDA_TRY(cudaDeviceCanAccessPeer(&can, h->device, b.device));
            if (!can) return fail(LBM_ERR_CUDA, "no peer access between the GPUs of neighbouring slabs");
        }
        const cudaIpcMemHandle_t *hs[3] = {&b.f0, &b.f1, &b.inbox};
        for (int i = 0; i < 3; ++i) CUDA_TRY(cudaIpcOpenMemHandle(&h->peer_base[side][i], *hs[i], cudaIpcMemLazyEnablePeerAccess));
        h->peer_nx_local[side] = b.nx_local;
    }
    h->peer_mode = true;
    return LBM_OK;
}

int lbm_destroy(LbmHandle h) {
    if (!h) return LBM_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    delete h;
    return LBM_OK;
}

int lbm_init(LbmHandle h) {
    if (int rc = check_handle(h, false)) return rc;
    lbm::init_kernel<<<1184, 256, 0, h->stream>>>(h->f[0], h->f[1], h->rho, h->ux, h->uy, h->plane, h->ny, h->pitch);
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    CUDA_TRY(cudaMemsetAsync(h->ctr, 0, 2 * sizeof(int), h->stream));
    CUDA_TRY(cudaMemsetAsync(h->maxv, 0, 2 * sizeof(unsigned), h->stream));
    h->steps_done = 0;
    h->inited = true;
    return LBM_OK;
}

int lbm_run(LbmHandle h, int steps) {
    if (int rc = check_handle(h, true)) return rc;
    if (steps < 0) return fail(LBM_ERR_INVALID, "steps < 0");
    const bool strict = h->p.arith == LBM_ARITH_STRICT;
    const int ncols = h->nx_local - 2;
    // grid of the register variant: x = segment blocks of a column, y (z) = rows, see step_kernel
    // `early_cols` > 0: early-start order -- the first column groups, then the W/E ring block, then the other groups --
    // and the rows up to one whole group (its top / bottom ring row included) past the ring signal the progress
    // counter: they are what the next step's early columns read (columns c-1..c+1, the ring cells of column
    // early_cols + 1 included) and overwrite (source columns whose last readers are that group's ring warps).  Returns the CTAs that signal per launch in `signals`.
    auto grid_for = [&](lbm::StepArgs &a, int early_cols = 0, unsigned long long *signals = nullptr) {
        const int gx = (h->nseg + lbm::kWarpsPerBlock - 1) / lbm::kWarpsPerBlock;
        const int G = lbm::kRingGroup;
        a.n_ring = lbm::ring_cell_count(a.il0, a.il_step, a.il_count, h->nx_local, h->ny, a.west_ring, a.east_ring);
        // rows: groups of G columns + 1 top/bottom ring row each; the W/E ring block behind the early groups
        const int vrows = (a.il_count + G - 1) / G * (G + 1);
        const int we_cells = a.n_ring - 2 * a.il_count;
        const int we_ctas = (we_cells + 32 * lbm::kWarpsPerBlock - 1) / (32 * lbm::kWarpsPerBlock);
        a.ring_rows = (we_ctas + gx - 1) / gx;
        a.ring_row0 = early_cols > 0 ? early_cols / G * (G + 1) : vrows;
        a.early_rows = 0;
        a.low_rows = early_cols > 0 ? a.ring_row0 + a.ring_rows + (G + 1) : 0;
        if (signals) *signals = (unsigned long long)a.low_rows * gx;
        const int rows = vrows + a.ring_rows;
        return dim3(gx, std::min(rows, 65535), (rows + 65534) / 65535);
    };
    // Early start needs PDL, one launch per step and a grid of many waves (the early columns must be a small
    // prefix whose inputs the previous step finished long before its tail).
    int early_cols = 0;
    {
        const int gx = (h->nseg + lbm::kWarpsPerBlock - 1) / lbm::kWarpsPerBlock;
        const int want = std::min((h->early_target + gx - 1) / gx, ncols / 3) / lbm::kRingGroup * lbm::kRingGroup;  // whole groups
        const bool single = !(h->comm && h->nranks > 1) || h->peer_mode;   // one launch per step
        // below ~2 waves of CTAs the previous step's first columns are not done when its last CTAs start: the
        // check would always fall through to the wait and the counter update would only lengthen the step
        const long long total_ctas = (long long)ncols * gx;
        if (h->use_pdl && single && !h->use_tma && want >= 4 && total_ctas >= h->early_min_ctas) early_cols = want;
    }
    lbm::StepArgs a_all = make_args(h);
    unsigned long long signals_all = 0;
    const dim3 blocks_all = grid_for(a_all, early_cols, &signals_all);
    // peer-memory slabs: the east edge column sits right behind the early block (see StepArgs::col_split)
    const int col_split = (h->peer_mode && h->peer_base[1][0]) ? early_cols : -1;
    if (!h->peer_mode && h->comm && h->nranks > 1 && h->stream_e) {  // the side stream starts behind everything already queued
        CUDA_TRY(cudaEventRecord(h->ev_m, h->stream));
        h->ev_e_prev_valid = false;
    }
    // Launch-bound grids (no early start: under ~2 waves of CTAs a step is a few microseconds, about what one launch costs
    // the host): once the soft-start ramp has reached its final value every per-step kernel argument repeats with the
    // buffer parity, so the whole batch -- K - 1 plain steps chained by programmatic edges, the max|u| reset, the EMIT
    // step -- is captured ONCE per (K, parity) and replayed with ONE cudaGraphLaunch per lbm_run: the host cost of a batch
    // no longer grows with K, which is what lets several cases in flight (one stream each, batch.py) fill the GPU
    // instead of queueing on the context's launch lock.  The diagnostic step counter is set behind the graph.
    const bool graphable = h->use_graph && !h->use_tma && !(h->comm && h->nranks > 1) && !h->peer_mode && early_cols == 0 &&
                           steps >= h->graph_min_steps && h->steps_done + 1 >= (int64_t)h->p.warmup_steps;
    auto enqueue = [&]() -> int {
    for (int it = 0; it < steps; ++it) {
        const bool emit = (it == steps - 1);
        if (emit) CUDA_TRY(cudaMemsetAsync(h->maxv, 0, 2 * sizeof(unsigned), h->stream));
        if (h->use_tma) {
            const int par = (int)(h->steps_done & 1);
            lbm::TmaArgs ta = h->tma_args;
            ta.ctr_in = h->ctr + par;
            ta.ctr_out = h->ctr + (par ^ 1);
            ta.ring = h->ring_ctx + (par ^ 1);
            const CUtensorMap &ms = h->map_src[par], &mh = h->map_srch[par], &md = h->map_dst[par ^ 1];
            const dim3 grid(h->tma_grid), block(lbm::kTmaThreads);
            const size_t sm = lbm::kTmaSmemBytes;
            if (strict) {
                if (emit) lbm::step_tma_kernel<true, true><<<grid, block, sm, h->stream>>>(ms, mh, h->map_code, md, h->map_mac, ta);
                else lbm::step_tma_kernel<true, false><<<grid, block, sm, h->stream>>>(ms, mh, h->map_code, md, h->map_mac, ta);
            } else {
                if (emit) lbm::step_tma_kernel<false, true><<<grid, block, sm, h->stream>>>(ms, mh, h->map_code, md, h->map_mac, ta);
                else lbm::step_tma_kernel<false, false><<<grid, block, sm, h->stream>>>(ms, mh, h->map_code, md, h->map_mac, ta);
            }
            h->steps_done++;
            h->steps_total++;
            h->launches++;
            if (int rc = exchange_halos(h, h->f[par ^ 1], h->stream)) return rc;
            continue;
        }
        lbm::StepArgs a = make_args(h);
        if (col_split >= 0) { a.col_split = col_split; }
        else if (h->peer_mode) { a.col_split = ncols; }   // identity order il = 1 + col
        if (h->peer_mode) {   // grid rows of the edge columns: col -> row = groups of 33, shifted past the W/E ring block
            auto row_of = [&](int col) {
                const int vrow = col / lbm::kRingGroup * (lbm::kRingGroup + 1) + col % lbm::kRingGroup;
                return vrow < a_all.ring_row0 ? vrow : vrow + a_all.ring_rows;
            };
            if (a.edge_il[0] >= 0) a.edge_row[0] = row_of(a.col_split == 0 ? 1 : 0);             // il = 1
            if (a.edge_il[1] >= 0) a.edge_row[1] = row_of(a.col_split < ncols ? a.col_split : ncols - 1);   // il = nx_local - 2
            if (ncols == 1)   // one owned column: it is both edges
                for (int side = 0; side < 2; ++side)
                    if (a.edge_il[side] >= 0) a.edge_row[side] = row_of(0);
        }
        const bool overlap = !h->peer_mode && h->comm && h->nranks > 1 && h->nx_local >= 6 && h->stream_e;
        cudaStream_t st = h->stream;
        dim3 blocks = blocks_all;
        a.n_ring = a_all.n_ring; a.ring_row0 = a_all.ring_row0; a.ring_rows = a_all.ring_rows; a.low_rows = a_all.low_rows;
        if (overlap) {
            // Edge columns (1 and nx_local-2) first on the side stream, halo exchange right behind them, the
            // interior on the main stream meanwhile.  edge(n) needs interior(n-1) and exchange(n-1);
            // interior(n) needs edge(n-1) and interior(n-1); see DESIGN.md section 5 "slabs".
            lbm::StepArgs e = a;
            e.il0 = 1; e.il_step = h->nx_local - 3; e.il_count = 2;
            if (emit) CUDA_TRY(cudaEventRecord(h->ev_m, h->stream));  // orders the max|u| reset before the edge kernel
            CUDA_TRY(cudaStreamWaitEvent(h->stream_e, h->ev_m, 0));
            const dim3 eb = grid_for(e);
            CUDA_TRY(launch_step(step_fn(strict, emit, h->links8 != nullptr), eb, h->stream_e, e, false));
            CUDA_TRY(cudaEventRecord(h->ev_e, h->stream_e));
            if (int rc = exchange_halos(h, a.dst, h->stream_e)) return rc;
            h->launches++;
            a.il0 = 2; a.il_step = 1; a.il_count = h->nx_local - 4;
            blocks = grid_for(a);
            CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_e_prev_valid ? h->ev_e_prev : h->ev_e, 0));
        }
        {
            // PDL between consecutive plain steps of a batch (not across the max|u| memset of an EMIT step)
            const bool pdl = h->use_pdl && !overlap && !emit && it > 0;
            // early start only straight behind a step that signalled (it > 0: the previous launch of this loop)
            a.early_rows = (pdl && early_cols > 0) ? a_all.ring_row0 : 0;
            a.progress_expected = h->progress_total;
            static const bool force_peer_kernel = std::getenv("LBM2D_FORCE_PEER_KERNEL") != nullptr;   // experiment: cost of the PEER code on one GPU
            if (force_peer_kernel && !h->peer_mode) a.col_split = ncols;
            CUDA_TRY(launch_step(step_fn(strict, emit, h->links8 != nullptr, h->peer_mode || force_peer_kernel), blocks, st, a, pdl));
            if (!overlap) h->progress_total += signals_all;
        }
        h->steps_done++;
        h->steps_total++;
        h->launches++;
        if (overlap) {
            CUDA_TRY(cudaEventRecord(h->ev_m, h->stream));
            std::swap(h->ev_e, h->ev_e_prev);   // interior(n+1) waits on edge(n)
            h->ev_e_prev_valid = true;
        } else if (!h->peer_mode) {
            if (int rc = exchange_halos(h, a.dst, h->stream)) return rc;
        }
    }
    return LBM_OK;
    };
    if (!graphable) {
        if (int rc = enqueue()) return rc;
    } else {
        const long long key = (long long)steps * 2 + (long long)(h->steps_done & 1);
        auto hit = h->graphs.find(key);
        if (hit == h->graphs.end()) {
            if (h->graphs.size() >= 8) {   // a run loop uses one or two batch sizes; a caller cycling through many must not pile up graphs
                for (auto &kv : h->graphs) cudaGraphExecDestroy(kv.second);
                h->graphs.clear();
            }
            // thread-local capture: other cases (threads) of the same process keep allocating / launching meanwhile
            CUDA_TRY(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
            const int64_t done0 = h->steps_done, total0 = h->steps_total, launches0 = h->launches;
            const int rc = enqueue();
            cudaGraph_t graph = nullptr;
            const cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
            h->steps_done = done0; h->steps_total = total0; h->launches = launches0;   // nothing has run yet
            if (rc != LBM_OK) { if (graph) cudaGraphDestroy(graph); (void)cudaGetLastError(); return rc; }
            if (ce != cudaSuccess) return fail(LBM_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
            cudaGraphExec_t exec = nullptr;
            const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ie != cudaSuccess) return fail(LBM_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie));
            hit = h->graphs.emplace(key, exec).first;
        }
        CUDA_TRY(cudaGraphLaunch(hit->second, h->stream));
        h->graph_replays++;
        h->steps_done += steps;
        h->steps_total += steps;
        h->launches += steps;
    }
    if (steps > 0 && !h->use_tma) {   // the diagnostic device copy of frame_count (the TMA kernel keeps its own)
        lbm::set_counter_kernel<<<1, 1, 0, h->stream>>>(h->ctr + (h->steps_done & 1), (int)h->steps_done);
        h->launches++;
    }
    if (h->peer_mode && steps > 0) h->halo_wait_pending = true;   // see halo_ready()
    if (!h->peer_mode && h->comm && h->nranks > 1 && h->stream_e) {  // later work on the main stream sees the last exchange
        CUDA_TRY(cudaEventRecord(h->ev_x, h->stream_e));
        CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_x, 0));
    }
    CUDA_TRY(cudaGetLastError());
    return LBM_OK;
}

// peer-memory slabs: kernels that read the halo columns of the CURRENT buffer (force links, exports of solid cells) must
// run behind the neighbours' stores of the final step; the step kernels themselves wait in their edge CTAs.
static int halo_ready(LbmHandle h) {
    if (!h->peer_mode || !h->halo_wait_pending) return LBM_OK;
    lbm::StepArgs w = make_args(h);
    lbm::halo_wait_kernel<<<1, 32, 0, h->stream>>>(w);
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    h->halo_wait_pending = false;
    return LBM_OK;
}

// peer-memory slabs: a halo wait that timed out (a neighbour died or fell 2 s behind) poisons the run; say so.
static int check_peer_error(LbmHandle h) {
    if (!h->peer_mode) return LBM_OK;
    unsigned long long flag = 0;
    CUDA_TRY(cudaMemcpyAsync(&flag, h->inbox + 2, sizeof(flag), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (flag & 0xffffffffULL) return fail(LBM_ERR_NCCL, "peer-memory halo exchange: a neighbour's halo column did not arrive within 2 s");
    return LBM_OK;
}

int lbm_synchronize(LbmHandle h) {
    if (int rc = check_handle(h, false)) return rc;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return check_peer_error(h);
}

int lbm_step_count(LbmHandle h, int64_t *steps) {
    if (int rc = check_handle(h, true)) return rc;
    if (!steps) return fail(LBM_ERR_INVALID, "steps is null");
    int v = 0;  // the device copy of frame_count (diagnostic: written by the step kernels / behind a replayed graph)
    CUDA_TRY(cudaMemcpyAsync(&v, h->ctr + (h->steps_done & 1), sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    *steps = v;
    return LBM_OK;
}

int lbm_get_force(LbmHandle h, float out_xy[2]) {
    if (int rc = check_handle(h, true)) return rc;
    if (!out_xy) return fail(LBM_ERR_INVALID, "out is null");
    out_xy[0] = out_xy[1] = 0.0f;
    if (h->n_links == 0) return lbm_synchronize(h);
    if (int rc = halo_ready(h)) return rc;
    const int par = (int)(h->steps_done & 1);
    lbm::force_kernel<<<kForceBlocks, 256, 0, h->stream>>>(h->f[par], h->plane, h->links, h->n_links, h->force_partial);
    lbm::force_final_kernel<<<1, 32, 0, h->stream>>>(h->force_partial, kForceBlocks, h->force_out);
    CUDA_TRY(cudaGetLastError());
    h->launches += 2;
    CUDA_TRY(cudaMemcpyAsync(out_xy, h->force_out, 2 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return LBM_OK;
}

int lbm_get_max_velocity(LbmHandle h, float *out) {
    if (int rc = check_handle(h, true)) return rc;
    if (!out) return fail(LBM_ERR_INVALID, "out is null");
    unsigned v[2] = {0, 0};
    CUDA_TRY(cudaMemcpyAsync(v, h->maxv, sizeof(v), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    float m2;
    std::memcpy(&m2, &v[0], sizeof(float));
    *out = v[1] ? NAN : std::sqrt(m2);  // max of sqrt == sqrt of max (ref:652-653)
    return check_peer_error(h);
}

static int get_planes(LbmHandle h, const float *p0, const float *p1, int nch, float *out) {
    if (int rc = check_handle(h, true)) return rc;
    if (!out) return fail(LBM_ERR_INVALID, "out is null");
    const size_t n = (size_t)h->p.nx * h->ny * nch;
    if (int rc = ensure_staging(h, n)) return rc;
    dim3 grid(h->p.nx, (h->ny + 127) / 128);
    lbm::pack_planes_kernel<<<grid, 128, 0, h->stream>>>(p0, p1, nch, h->own0, h->ny, h->pitch, h->staging);
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    return d2h(h, out, h->staging, n * sizeof(float));
}

int lbm_get_vel(LbmHandle h, float *out) { return h ? get_planes(h, h->ux, h->uy, 2, out) : fail(LBM_ERR_INVALID, "null handle"); }
int lbm_get_rho(LbmHandle h, float *out) { return h ? get_planes(h, h->rho, nullptr, 1, out) : fail(LBM_ERR_INVALID, "null handle"); }

int lbm_get_mask(LbmHandle h, float *out) {
    if (int rc = check_handle(h, false)) return rc;
    if (!out) return fail(LBM_ERR_INVALID, "out is null");
    const size_t n = (size_t)h->p.nx * h->ny;
    if (int rc = ensure_staging(h, n)) return rc;
    dim3 grid(h->p.nx, (h->ny + 127) / 128);
    lbm::mask_to_float_kernel<<<grid, 128, 0, h->stream>>>(h->code, h->own0, h->ny, h->pitch, h->staging);
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    return d2h(h, out, h->staging, n * sizeof(float));
}

static int export9(LbmHandle h, int mode, float *out) {
    if (int rc = check_handle(h, true)) return rc;
    if (!out) return fail(LBM_ERR_INVALID, "out is null");
    const size_t n = (size_t)h->p.nx * h->ny * 9;
    if (int rc = ensure_staging(h, n)) return rc;
    if (int rc = halo_ready(h)) return rc;
    const lbm::ExportArgs a = make_export_args(h);
    dim3 grid(h->p.nx, (h->ny + 127) / 128);
    lbm::export9_kernel<<<grid, 128, 0, h->stream>>>(a, mode, h->staging);
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    return d2h(h, out, h->staging, n * sizeof(float));
}

int lbm_get_viz_fields(LbmHandle h, const double *weights, int radius, float *out_mag, float *out_vor) {
    if (int rc = check_handle(h, true)) return rc;
    if (!out_mag || !out_vor) return fail(LBM_ERR_INVALID, "out is null");
    if (radius < 0 || radius > 4096 || (radius > 0 && !weights)) return fail(LBM_ERR_INVALID, "bad filter radius / weights");
    const bool slabs = h->comm && h->nranks > 1;
    if (!slabs && !(h->west_ring && h->east_ring)) return fail(LBM_ERR_STATE, "viz fields on a slab handle need lbm_comm_connect() first");
    const int nx = h->p.nx, ny = h->ny, NX = h->p.nx_global, x0 = h->p.slab_x0;
    // x-slabs: the x pass of the filter reaches `radius` columns into the neighbours and the vorticity one column
    // further, so radius + 1 raw velocity columns come from each neighbour (NCCL, contiguous: y is the fast index);
    // reflection and the one-sided differences happen at the GLOBAL edges, so the owned columns are bit-identical to
    // the single-GPU result.
    const int hw = h->west_ring ? 0 : radius + 1, he = h->east_ring ? 0 : radius + 1;   // raw halo columns
    const int ew = h->west_ring ? 0 : 1, ee = h->east_ring ? 0 : 1;                     // filtered halo columns
    if (slabs && radius + 1 > nx) return fail(LBM_ERR_INVALID, "viz fields: the filter radius exceeds the slab width");
    const int nxe = hw + nx + he, nxo = ew + nx + ee;
    const size_t n = (size_t)nx * ny, no = (size_t)nxo * ny, wfloats = 2 * ((size_t)radius + 2);   // the weights ride in front (8-byte aligned)
    const size_t ext = (size_t)nxe * h->pitch;
    if (int rc = ensure_staging(h, wfloats + 2 * ext + 4 * no + 2 * n)) return rc;
    double *w = reinterpret_cast<double *>(h->staging);
    float *e0 = h->staging + wfloats, *e1 = e0 + ext, *t0 = e1 + ext, *t1 = t0 + no, *v0 = t1 + no, *v1 = v0 + no, *mag = v1 + no, *vor = mag + n;
    // raw velocity with neighbour columns: row 0 = global column x0 - hw, pitched like the planes
    const float *rx = h->ux + (size_t)h->own0 * h->pitch, *ry = h->uy + (size_t)h->own0 * h->pitch;
    int gin0 = x0;
    if (slabs) {
        const size_t own = (size_t)nx * h->pitch;
        CUDA_TRY(cudaMemcpyAsync(e0 + (size_t)hw * h->pitch, rx, own * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
        CUDA_TRY(cudaMemcpyAsync(e1 + (size_t)hw * h->pitch, ry, own * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
        nccl::Api &nc = nccl::api();
        const size_t cnt = (size_t)(radius + 1) * h->pitch;
        NCCL_TRY(nc.GroupStart());
        if (!h->west_ring) {
            NCCL_TRY(nc.Send(rx, cnt, nccl::kFloat32, h->rank - 1, h->comm, h->stream));
            NCCL_TRY(nc.Send(ry, cnt, nccl::kFloat32, h->rank - 1, h->comm, h->stream));
            NCCL_TRY(nc.Recv(e0, cnt, nccl::kFloat32, h->rank - 1, h->comm, h->stream));
            NCCL_TRY(nc.Recv(e1, cnt, nccl::kFloat32, h->rank - 1, h->comm, h->stream));
        }
        if (!h->east_ring) {
            NCCL_TRY(nc.Send(rx + own - cnt, cnt, nccl::kFloat32, h->rank + 1, h->comm, h->stream));
            NCCL_TRY(nc.Send(ry + own - cnt, cnt, nccl::kFloat32, h->rank + 1, h->comm, h->stream));
            NCCL_TRY(nc.Recv(e0 + (size_t)(hw + nx) * h->pitch, cnt, nccl::kFloat32, h->rank + 1, h->comm, h->stream));
            NCCL_TRY(nc.Recv(e1 + (size_t)(hw + nx) * h->pitch, cnt, nccl::kFloat32, h->rank + 1, h->comm, h->stream));
        }
        NCCL_TRY(nc.GroupEnd());
        rx = e0;
        ry = e1;
        gin0 = x0 - hw;
    }
    const dim3 grid(nx, (ny + 127) / 128), grid_o(nxo, (ny + 127) / 128);
    const float *vx = rx, *vy = ry;
    long long sx = h->pitch;
    int gv0 = gin0;   // global column of row 0 of (vx, vy)
    if (radius > 0) {
        CUDA_TRY(cudaMemcpyAsync(w, weights, ((size_t)radius + 1) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        lbm::viz_blur_kernel<<<grid_o, 128, 0, h->stream>>>(rx, ry, h->pitch, nxo, ny, 0, radius, w, t0, t1, x0 - ew, gin0, NX);
        lbm::viz_blur_kernel<<<grid_o, 128, 0, h->stream>>>(t0, t1, ny, nxo, ny, 1, radius, w, v0, v1, 0, 0, NX);
        vx = v0; vy = v1; sx = ny; gv0 = x0 - ew;
        h->launches += 2;
    }
    lbm::viz_fields_kernel<<<grid, 128, 0, h->stream>>>(vx, vy, sx, nx, ny, mag, vor, x0, gv0, NX);
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    if (int rc = d2h(h, out_mag, mag, n * sizeof(float))) return rc;
    return d2h(h, out_vor, vor, n * sizeof(float));
}

int lbm_get_moments(LbmHandle h, float *out) { return export9(h, 0, out); }
int lbm_get_f(LbmHandle h, int which, float *out) {
    if (which != 0 && which != 1) return fail(LBM_ERR_INVALID, "which must be 0 (f_old) or 1 (f_new)");
    return export9(h, which == 0 ? 1 : 2, out);
}

// computeResizeAreaTab of OpenCV's resize.cpp (see oracle/writer_oracle.py::area_tab), grouped by destination
static double area_tab(int ssize, int dsize, std::vector<lbm::AreaEntry> &tab, std::vector<int> &off) {
    const double inv = (double)dsize / (double)ssize, scale = 1.0 / inv;
    tab.clear();
    off.assign(dsize + 1, 0);
    for (int dx = 0; dx < dsize; ++dx) {
        off[dx] = (int)tab.size();
        const double fsx1 = dx * scale, fsx2 = fsx1 + scale, cell = std::min(scale, ssize - fsx1);
        int sx1 = (int)std::ceil(fsx1), sx2 = (int)std::floor(fsx2);
        sx2 = std::min(sx2, ssize - 1);
        sx1 = std::min(sx1, sx2);
        if (sx1 - fsx1 > 1e-3) tab.push_back({sx1 - 1, (float)((sx1 - fsx1) / cell)});
        for (int sx = sx1; sx < sx2; ++sx) tab.push_back({sx, (float)(1.0 / cell)});
        if (fsx2 - sx2 > 1e-3) tab.push_back({sx2, (float)(std::min(std::min(fsx2 - sx2, 1.0), cell) / cell)});
    }
    off[dsize] = (int)tab.size();
    return scale;
}

int lbm_export_configure(LbmHandle h, const LbmExportConfig *cfg) {
    if (int rc = check_handle(h, false)) return rc;
    if (!cfg) return fail(LBM_ERR_INVALID, "cfg is null");
    const bool slabs = h->p.nx != h->p.nx_global;
    if (slabs && !h->comm) return fail(LBM_ERR_STATE, "slab handles need lbm_comm_connect() before lbm_export_configure()");
    const int X0 = cfg->x0, X1 = cfg->x1, cw_g = X1 - X0, ch = cfg->y1 - cfg->y0;
    if (X0 < 0 || cfg->y0 < 0 || X1 > h->p.nx_global || cfg->y1 > h->ny || cw_g <= 0 || ch <= 0)
        return fail(LBM_ERR_INVALID, "export ROI outside the grid or empty");
    if (cfg->target_w < 1 || cfg->target_h < 1 || cfg->target_w > cw_g || cfg->target_h > ch)
        return fail(LBM_ERR_INVALID, "INTER_AREA export supports shrinking only (1 <= target <= crop)");
    for (void **ptr : {(void **)&h->exp_xtab, (void **)&h->exp_ytab, (void **)&h->exp_xoff, (void **)&h->exp_yoff, (void **)&h->exp_tmp,
                       (void **)&h->exp_frame, (void **)&h->exp_sum, (void **)&h->exp_velsq, (void **)&h->exp_vor,
                       (void **)&h->exp_minmax, (void **)&h->exp_halo, (void **)&h->exp_ecount}) {
        if (*ptr) devpool::release(*ptr);
        *ptr = nullptr;
    }
    h->exp_ready = false;

    std::vector<lbm::AreaEntry> xt, yt;
    std::vector<int> xo, yo;
    const double sx = area_tab(cw_g, cfg->target_w, xt, xo), sy = area_tab(ch, cfg->target_h, yt, yo);
    lbm::ExportGeom &g = h->exp_geom;
    g.tw_g = cfg->target_w;
    g.th = cfg->target_h;
    g.ch = ch;
    g.y0 = cfg->y0;
    g.ix = (int)std::lrint(sx);
    g.iy = (int)std::lrint(sy);
    g.fast = std::fabs(sx - g.ix) < 2.220446049250313e-16 && std::fabs(sy - g.iy) < 2.220446049250313e-16;
    // this rank's share: the ROI columns it owns, and the output columns whose first source column is one of them
    const int gx0 = h->p.slab_x0, gx1 = h->p.slab_x0 + h->p.nx;
    const int rx0 = std::min(std::max(gx0, X0), X1), rx1 = std::max(std::min(gx1, X1), rx0);
    g.own_cols = rx1 - rx0;
    g.x0 = rx0 - h->x_off;
    g.src_shift = rx0 - X0;
    int dlo = 0, dhi = 0;
    for (int dx = 0; dx < g.tw_g; ++dx) {
        const int first = X0 + xt[xo[dx]].si;
        if (first < rx0) ++dlo;
        if (first < rx1) ++dhi;
    }
    if (g.own_cols == 0) dhi = dlo;
    g.dlo = dlo;
    g.dhi = dhi;
    int e_recv = 0;
    if (dhi > dlo) e_recv = std::max(0, X0 + xt[xo[dhi] - 1].si - (rx1 - 1));
    int e_send = 0;
    CUDA_TRY(h->dalloc(&h->exp_ecount, 2 * sizeof(int)));
    if (slabs) {  // tell the east neighbour how many of its first ROI columns this rank needs
        nccl::Api &n = nccl::api();
        CUDA_TRY(cudaMemcpyAsync(h->exp_ecount, &e_recv, sizeof(int), cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaMemsetAsync(h->exp_ecount + 1, 0, sizeof(int), h->stream));
        NCCL_TRY(n.GroupStart());
        if (!h->east_ring) NCCL_TRY(n.Send(h->exp_ecount, 1, nccl::kInt32, h->rank + 1, h->comm, h->stream));
        if (!h->west_ring) NCCL_TRY(n.Recv(h->exp_ecount + 1, 1, nccl::kInt32, h->rank - 1, h->comm, h->stream));
        NCCL_TRY(n.GroupEnd());
        CUDA_TRY(cudaMemcpyAsync(&e_send, h->exp_ecount + 1, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        if (e_send > g.own_cols) return fail(LBM_ERR_INVALID, "slab too narrow for the export reduction (neighbour needs more ROI columns than this rank owns)");
    } else if (e_recv != 0) {
        return fail(LBM_ERR_INVALID, "internal: single-GPU export needs no extension");
    }
    h->exp_send_cols = e_send;
    h->exp_recv_cols = e_recv;
    g.cw = g.own_cols + e_recv;

    const size_t npx = (size_t)std::max(1, dhi - dlo) * g.th;
    CUDA_TRY(h->dalloc(&h->exp_xtab, xt.size() * sizeof(lbm::AreaEntry)));
    CUDA_TRY(h->dalloc(&h->exp_ytab, yt.size() * sizeof(lbm::AreaEntry)));
    CUDA_TRY(h->dalloc(&h->exp_xoff, xo.size() * sizeof(int)));
    CUDA_TRY(h->dalloc(&h->exp_yoff, yo.size() * sizeof(int)));
    CUDA_TRY(h->dalloc(&h->exp_tmp, (size_t)9 * std::max(1, g.cw) * ch * sizeof(float)));
    CUDA_TRY(h->dalloc(&h->exp_frame, 9 * npx * sizeof(float)));
    CUDA_TRY(h->dalloc(&h->exp_sum, 9 * npx * sizeof(double)));
    CUDA_TRY(h->dalloc(&h->exp_velsq, npx * sizeof(double)));
    CUDA_TRY(h->dalloc(&h->exp_vor, npx * sizeof(double)));
    CUDA_TRY(h->dalloc(&h->exp_minmax, 18 * sizeof(double)));
    CUDA_TRY(h->dalloc(&h->exp_halo, (size_t)4 * 3 * g.th * sizeof(float)));
    CUDA_TRY(cudaMemcpy(h->exp_xtab, xt.data(), xt.size() * sizeof(lbm::AreaEntry), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(h->exp_ytab, yt.data(), yt.size() * sizeof(lbm::AreaEntry), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(h->exp_xoff, xo.data(), xo.size() * sizeof(int), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(h->exp_yoff, yo.data(), yo.size() * sizeof(int), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemset(h->exp_sum, 0, 9 * npx * sizeof(double)));
    CUDA_TRY(cudaMemset(h->exp_velsq, 0, npx * sizeof(double)));
    CUDA_TRY(cudaMemset(h->exp_vor, 0, npx * sizeof(double)));
    CUDA_TRY(cudaMemset(h->exp_halo, 0, (size_t)4 * 3 * g.th * sizeof(float)));
    double mm[18];
    for (int c = 0; c < 9; ++c) { mm[c] = INFINITY; mm[9 + c] = -INFINITY; }
    CUDA_TRY(cudaMemcpy(h->exp_minmax, mm, sizeof(mm), cudaMemcpyHostToDevice));
    h->exp_count = 0;
    h->exp_ready = true;
    return LBM_OK;
}

int lbm_export_layout(LbmHandle h, int32_t *dlo, int32_t *dhi, int32_t *target_h) {
    if (!h || !h->exp_ready) return fail(LBM_ERR_STATE, "lbm_export_configure() has not been called");
    if (dlo) *dlo = h->exp_geom.dlo;
    if (dhi) *dhi = h->exp_geom.dhi;
    if (target_h) *target_h = h->exp_geom.th;
    return LBM_OK;
}

static int export_frame_impl(LbmHandle h, float *out_chw);

int lbm_export_frame(LbmHandle h, float *out_chw) { return export_frame_impl(h, out_chw); }

int lbm_export_frame_device(LbmHandle h, const float **frame_chw_dev) {
    if (!frame_chw_dev) return fail(LBM_ERR_INVALID, "frame_chw_dev is null");
    if (int rc = export_frame_impl(h, nullptr)) return rc;
    *frame_chw_dev = (h->exp_geom.dhi - h->exp_geom.dlo) > 0 ? h->exp_frame : nullptr;
    return LBM_OK;
}

static int export_frame_impl(LbmHandle h, float *out_chw) {
    if (int rc = check_handle(h, true)) return rc;
    if (!h->exp_ready) return fail(LBM_ERR_STATE, "lbm_export_configure() has not been called");
    if (int rc = halo_ready(h)) return rc;
    const lbm::ExportGeom g = h->exp_geom;
    const lbm::ExportArgs a = make_export_args(h);
    const int twl = g.dhi - g.dlo;
    const size_t npx = (size_t)twl * g.th;
    const bool slabs = h->comm && h->nranks > 1;
    nccl::Api &n = nccl::api();
    if (g.own_cols > 0) {
        lbm::roi_moments_kernel<<<dim3(g.own_cols, (g.ch + 127) / 128), 128, 0, h->stream>>>(a, g, h->exp_tmp);
        h->launches++;
    }
    if (slabs && (h->exp_send_cols > 0 || h->exp_recv_cols > 0)) {
        // the last output columns of a rank reach into the first ROI columns of its east neighbour
        const size_t pl = (size_t)g.cw * g.ch;
        NCCL_TRY(n.GroupStart());
        for (int c = 0; c < 9; ++c) {
            if (h->exp_send_cols > 0) NCCL_TRY(n.Send(h->exp_tmp + c * pl, (size_t)h->exp_send_cols * g.ch, nccl::kFloat32, h->rank - 1, h->comm, h->stream));
            if (h->exp_recv_cols > 0) NCCL_TRY(n.Recv(h->exp_tmp + c * pl + (size_t)g.own_cols * g.ch, (size_t)h->exp_recv_cols * g.ch, nccl::kFloat32, h->rank + 1, h->comm, h->stream));
        }
        NCCL_TRY(n.GroupEnd());
    }
    if (twl > 0) {
        const dim3 rgrid(twl, (g.th + 63) / 64, 9);
        if (g.fast) lbm::area_fast_kernel<<<rgrid, 64, 0, h->stream>>>(h->exp_tmp, g, h->exp_frame);
        else lbm::area_resize_kernel<<<rgrid, 64, 0, h->stream>>>(h->exp_tmp, g, h->exp_xtab, h->exp_xoff, h->exp_ytab, h->exp_yoff, h->exp_frame);
        h->launches++;
    }
    // x-gradient halos: rho, jx, jy of the neighbours' adjacent output columns
    float *send_w = h->exp_halo, *send_e = h->exp_halo + 3 * g.th, *left = h->exp_halo + 6 * g.th, *right = h->exp_halo + 9 * g.th;
    const bool nb_w = slabs && twl > 0 && g.dlo > 0, nb_e = slabs && twl > 0 && g.dhi < g.tw_g;
    if (slabs) {
        // every rank takes part; a rank without output columns forwards nothing (its neighbours' columns are then
        // not adjacent to it, which only happens at the ends of the ROI where the global edge rule applies)
        if (nb_w) lbm::export_pack_column_kernel<<<(g.th + 127) / 128, 128, 0, h->stream>>>(h->exp_frame, twl, g.th, 0, send_w);
        if (nb_e) lbm::export_pack_column_kernel<<<(g.th + 127) / 128, 128, 0, h->stream>>>(h->exp_frame, twl, g.th, twl - 1, send_e);
        NCCL_TRY(n.GroupStart());
        if (!h->west_ring) {
            NCCL_TRY(n.Send(send_w, 3 * g.th, nccl::kFloat32, h->rank - 1, h->comm, h->stream));
            NCCL_TRY(n.Recv(left, 3 * g.th, nccl::kFloat32, h->rank - 1, h->comm, h->stream));
        }
        if (!h->east_ring) {
            NCCL_TRY(n.Send(send_e, 3 * g.th, nccl::kFloat32, h->rank + 1, h->comm, h->stream));
            NCCL_TRY(n.Recv(right, 3 * g.th, nccl::kFloat32, h->rank + 1, h->comm, h->stream));
        }
        NCCL_TRY(n.GroupEnd());
    }
    if (twl > 0) {
        lbm::export_stats_kernel<<<dim3((twl + 127) / 128, g.th), 128, 0, h->stream>>>(h->exp_frame, twl, g.th, g.dlo, g.tw_g, left, right, h->exp_sum, h->exp_velsq, h->exp_vor);
        lbm::export_minmax_kernel<<<9, 256, 0, h->stream>>>(h->exp_frame, (long long)npx, h->exp_minmax, h->exp_minmax + 9);
        h->launches += 2;
    }
    CUDA_TRY(cudaGetLastError());
    h->exp_count++;
    if (out_chw && twl > 0) CUDA_TRY(cudaMemcpyAsync(out_chw, h->exp_frame, 9 * npx * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return LBM_OK;
}

int lbm_export_stats(LbmHandle h, double *running_sum_chw, double *vel_sq_sum_hw, double *abs_vor_sum_hw, double *min9,
                     double *max9, int64_t *count) {
    if (int rc = check_handle(h, false)) return rc;
    if (!h->exp_ready) return fail(LBM_ERR_STATE, "lbm_export_configure() has not been called");
    const size_t npx = (size_t)(h->exp_geom.dhi - h->exp_geom.dlo) * h->exp_geom.th;
    if (npx > 0) {
        if (running_sum_chw) CUDA_TRY(cudaMemcpyAsync(running_sum_chw, h->exp_sum, 9 * npx * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        if (vel_sq_sum_hw) CUDA_TRY(cudaMemcpyAsync(vel_sq_sum_hw, h->exp_velsq, npx * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        if (abs_vor_sum_hw) CUDA_TRY(cudaMemcpyAsync(abs_vor_sum_hw, h->exp_vor, npx * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    if (min9) CUDA_TRY(cudaMemcpyAsync(min9, h->exp_minmax, 9 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (max9) CUDA_TRY(cudaMemcpyAsync(max9, h->exp_minmax + 9, 9 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (count) *count = h->exp_count;
    return LBM_OK;
}

int lbm_static_mask(LbmHandle h, int32_t x0, int32_t x1, int32_t y0, int32_t y1, int32_t tw, int32_t th, float *out,
                    int32_t *degenerate) {
    if (int rc = check_handle(h, false)) return rc;
    if (!out || !degenerate) return fail(LBM_ERR_INVALID, "out / degenerate is null");
    if (h->p.nx != h->p.nx_global) return fail(LBM_ERR_INVALID, "static mask: single GPU only (the mask of a slab is partial)");
    const int cw = x1 - x0, ch = y1 - y0;
    if (x0 < 0 || y0 < 0 || x1 > h->p.nx || y1 > h->ny || cw <= 0 || ch <= 0 || tw < 1 || th < 1)
        return fail(LBM_ERR_INVALID, "static mask: ROI outside the grid or empty target");
    // cv::resize INTER_NEAREST (resizeNN): x_ofs[x] = min(cvFloor(x * ifx), ssize.width - 1), ifx = 1 / (dsize / ssize)
    std::vector<int> xs(tw), ys(th);
    const double ifx = 1.0 / ((double)tw / (double)cw), ify = 1.0 / ((double)th / (double)ch);
    for (int x = 0; x < tw; ++x) xs[x] = x0 + std::min((int)std::floor(x * ifx), cw - 1);
    for (int y = 0; y < th; ++y) ys[y] = y0 + std::min((int)std::floor(y * ify), ch - 1);
    const size_t n = (size_t)tw * th;
    // scratch layout (bytes): xs | ys | counts | small | g | d_fluid | d_solid | out
    const size_t off_ys = (size_t)tw * 4, off_cnt = off_ys + (size_t)th * 4, off_small = (off_cnt + 4 + 15) / 16 * 16;
    const size_t off_g = (off_small + n + 15) / 16 * 16, off_df = (off_g + n * 4 + 15) / 16 * 16, off_ds = off_df + n * 8;
    const size_t off_out = off_ds + n * 8, total = off_out + 2 * n * 4;
    if (int rc = ensure_staging(h, (total + 3) / 4)) return rc;
    char *base = reinterpret_cast<char *>(h->staging);
    int *d_xs = (int *)base, *d_ys = (int *)(base + off_ys), *d_cnt = (int *)(base + off_cnt), *d_g = (int *)(base + off_g);
    uint8_t *d_small = (uint8_t *)(base + off_small);
    double *d_df = (double *)(base + off_df), *d_ds = (double *)(base + off_ds);
    float *d_out = (float *)(base + off_out);
    CUDA_TRY(cudaMemcpyAsync(d_xs, xs.data(), (size_t)tw * 4, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(d_ys, ys.data(), (size_t)th * 4, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemsetAsync(d_cnt, 0, 4, h->stream));
    const dim3 grid2((tw + 127) / 128, th);
    lbm::mask_nearest_kernel<<<grid2, 128, 0, h->stream>>>(h->code, h->pitch, d_xs, d_ys, tw, th, d_small, d_cnt);
    int solids = 0;
    CUDA_TRY(cudaMemcpyAsync(&solids, d_cnt, 4, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->launches++;
    // no solid or no fluid pixel: scipy's transform of an image without background is an artefact of its algorithm, not
    // a distance -- the caller reproduces it with scipy itself
    *degenerate = (solids == 0 || (size_t)solids == n) ? 1 : 0;
    if (*degenerate) return LBM_OK;
    lbm::edt_columns_kernel<<<(tw + 127) / 128, 128, 0, h->stream>>>(d_small, tw, th, 1, d_g);   // fluid pixels: distance to the nearest solid
    lbm::edt_rows_kernel<<<grid2, 128, 0, h->stream>>>(d_g, tw, th, d_df);
    lbm::edt_columns_kernel<<<(tw + 127) / 128, 128, 0, h->stream>>>(d_small, tw, th, 0, d_g);   // solid pixels: distance to the nearest fluid
    lbm::edt_rows_kernel<<<grid2, 128, 0, h->stream>>>(d_g, tw, th, d_ds);
    lbm::sdf_combine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(d_small, d_df, d_ds, (long long)n, d_out);
    CUDA_TRY(cudaGetLastError());
    h->launches += 5;
    CUDA_TRY(cudaMemcpyAsync(out, d_out, 2 * n * 4, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return LBM_OK;
}

int lbm_device_view(LbmHandle h, LbmDeviceView *out) {
    if (int rc = check_handle(h, false)) return rc;
    if (!out) return fail(LBM_ERR_INVALID, "out is null");
    const int par = (int)(h->steps_done & 1);
    out->f_cur = h->f[par];
    out->f_prev = h->f[par ^ 1];
    out->rho = h->rho;
    out->ux = h->ux;
    out->uy = h->uy;
    out->cell_code = h->code;
    out->nx_local = h->nx_local;
    out->ny = h->ny;
    out->pitch = h->pitch;
    out->plane_stride = h->plane;
    out->stream = (void *)h->stream;
    return LBM_OK;
}

int lbm_host_alloc(size_t bytes, void **out) {
    if (!out || bytes == 0) return fail(LBM_ERR_INVALID, "bad argument");
    CUDA_TRY(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return LBM_OK;
}

int lbm_host_free(void *ptr) {
    if (ptr) CUDA_TRY(cudaFreeHost(ptr));
    return LBM_OK;
}

int lbm_selftest_arith(int64_t pairs, uint64_t seed, int64_t mismatches[3]) {
    if (!mismatches || pairs < 0) return fail(LBM_ERR_INVALID, "bad argument");
    unsigned long long *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, 3 * sizeof(unsigned long long)));
    CUDA_TRY(cudaMemset(d, 0, 3 * sizeof(unsigned long long)));
    const int blocks = 1184, threads = 256;
    const long long per = (pairs + (long long)blocks * threads - 1) / ((long long)blocks * threads);
    lbm::selftest_arith_kernel<<<blocks, threads>>>(per, seed, d);
    unsigned long long h[3] = {0, 0, 0};
    cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(LBM_ERR_CUDA, cudaGetErrorString(e));
    for (int k = 0; k < 3; ++k) mismatches[k] = (int64_t)h[k];
    return LBM_OK;
}

int lbm_launch_count(LbmHandle h, int64_t *launches) {
    if (!h || !launches) return fail(LBM_ERR_INVALID, "null argument");
    *launches = h->launches;
    return LBM_OK;
}

int lbm_graph_replay_count(LbmHandle h, int64_t *replays) {
    if (!h || !replays) return fail(LBM_ERR_INVALID, "null argument");
    *replays = h->graph_replays;
    return LBM_OK;
}

}  // extern "C"
