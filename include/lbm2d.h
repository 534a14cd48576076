/*
 * lbm2d.h -- C ABI of the B200-native D2Q9 MRT-LES lattice-Boltzmann time step.
 *
 * Drop-in boundary for the hot path of ms-112-scott/01-lbm-2d.  The reference has no FFI layer
 * of its own: its boundary is the Python class `LBM2D_MRT_LES`
 * (src/lbm_mrt_les/core/LBM2D_MRT_LES.py:10) driven by `run_simulation_loop`
 * (src/lbm_mrt_les/core/simulation_ops.py:60).  Every entry point below replaces one member of
 * that class (cited per function as ref:LINE of LBM2D_MRT_LES.py); the ctypes binding a
 * maintainer would add is in INTEGRATION.md and shipped as 01-lbm-2d_b200/_capi.py.
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success, non-zero LbmStatus on error;
 *     lbm_last_error() gives the message for the calling thread;
 *   - the handle owns all device memory; the caller owns every host buffer it passes;
 *   - host arrays use the reference's layout: C order with y fastest -- (nx,ny), (nx,ny,2),
 *     (nx,ny,9) -- fp32;
 *   - calls are stream-ordered on the handle's stream; lbm_run() is asynchronous, every getter
 *     synchronises before returning;
 *   - one host thread drives a handle at a time (as in the reference).
 *   - There is NO CPU fallback: without a CUDA device lbm_create() fails with LBM_ERR_CUDA.
 */
#ifndef LBM2D_H_
#define LBM2D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LBM2D_ABI_VERSION 2

typedef struct LbmSolver *LbmHandle;

typedef enum {
    LBM_OK = 0,
    LBM_ERR_INVALID = 1, /* bad argument / unsupported parameter */
    LBM_ERR_CUDA = 2,    /* CUDA runtime error (message holds cudaGetErrorString) */
    LBM_ERR_STATE = 3,   /* call sequence error (e.g. run before init) */
    LBM_ERR_NCCL = 4     /* halo exchange / collective error */
} LbmStatus;

typedef enum {
    LBM_ARITH_FAST = 0,  /* FMA + re-associated sparse transforms, SFU reciprocal / sqrt (tolerance-level parity) */
    LBM_ARITH_STRICT = 1 /* reference evaluation order, every operation individually rounded, no FMA:
                            bit-identical to the fp32 oracle */
} LbmArith;

typedef enum {
    LBM_KERNEL_AUTO = 0,     /* the fastest measured variant (REGISTER) */
    LBM_KERNEL_REGISTER = 1, /* one warp per 64-cell column segment, 2 cells per thread: 64-bit accesses, warp shuffles for
                                the y shift, packed fp32 pairs (FADD2) in the strict collision */
    LBM_KERNEL_TMA = 2       /* persistent CTAs, cp.async.bulk.tensor tiles through shared memory, mbarrier pipeline */
} LbmKernel;

typedef enum {
    LBM_OBSTACLE_REFILL = 0,     /* wet-node equilibrium refill, ref:452-455 (what the reference does; parity mode) */
    LBM_OBSTACLE_BOUNCE_BACK = 1 /* half-way bounce-back on solid links, solids frozen at rest: NOT reference behaviour,
                                    optional (single GPU, default kernel); checked against oracle/lbm_oracle_np.py */
} LbmObstacleMode;

/*
 * The keys the reference solver reads from the per-case YAML (ref:32-94, :114-119), passed as the
 * Python floats they are (double); the library derives the fp32 kernel constants from them the
 * way Taichi does (tau0 = f32(3 nu + 0.5), tau0^2 and 18 Cs^2 folded in double, then rounded).
 */
typedef struct {
    int32_t nx, ny;          /* simulation.nx / ny                                  ref:37-38 */
    int32_t warmup_steps;    /* simulation.warmup_steps (0 -> ramp == 1)             ref:40, :442 */
    double nu;               /* simulation.nu                                        ref:43 */
    double rho_in, rho_out;  /* simulation.rho_in / rho_out                          ref:53-54 */
    double c_smag;           /* simulation.smagorinsky_constant (LES on if > 0.001)  ref:78, :342 */
    double s_ghost;          /* simulation.ghost_moments_s                           ref:82 */
    int32_t sponge_in, sponge_out, sponge_top, sponge_bot; /* raw; max(1, .) applied  ref:90-93 */
    double sponge_strength;  /*                                                      ref:94 */
    int32_t bc_type[4];      /* boundary_condition.type  [W, top, E, bottom]         ref:115-118 */
    float bc_value[4][2];    /* boundary_condition.value                             ref:116-119 */
    int32_t arith;           /* LbmArith */
    int32_t obstacle_mode;   /* LbmObstacleMode */
    int32_t device;          /* CUDA device ordinal, -1 = current device */
    int32_t kernel;          /* LbmKernel */
    /* x-slab decomposition (single GPU: nx_global = nx, slab_x0 = 0).  A slab owns global columns
     * [slab_x0, slab_x0 + nx); `nx` above is then the OWNED width, and one halo column is kept on
     * every side that is not a domain boundary. */
    int32_t nx_global;
    int32_t slab_x0;
} LbmParams;

/* ABI / build information. */
int lbm_abi_version(void);
const char *lbm_last_error(void);
int lbm_device_count(int *count);

/* ctor, ref:13-29 (+ :97-128 fields, :131-201 constants).  `mask_xy` is nx*ny bytes, y fastest,
 * non-zero = solid (the reference's bool mask, ref:107-111); NULL = all fluid.  For a slab the mask
 * covers the owned columns plus the halo columns that exist (west halo first). */
int lbm_create(const LbmParams *params, const uint8_t *mask_xy, LbmHandle *out);
int lbm_destroy(LbmHandle h);

/* init(), ref:235-241: rho = 1, u = 0, f_old = f_new = f_eq, frame_count = 0. */
int lbm_init(LbmHandle h);

/* run_step(steps), ref:552-573: `steps` fused collide+stream+macro+BC+refill passes.  Asynchronous.
 * The last pass of a call also materialises rho / vel and the max|u| reduction (ref:648-654).
 * One kernel launch per step; on grids of less than ~2 waves of CTAs a whole call is replayed as one CUDA graph
 * (captured once per `steps` value and buffer parity) once the soft-start ramp is over -- same results, one host call. */
int lbm_run(LbmHandle h, int steps);
int lbm_synchronize(LbmHandle h);

/* frame_count[None], ref:122 / :440 */
int lbm_step_count(LbmHandle h, int64_t *steps);

/* get_force(), ref:644-646 (momentum exchange over solid-fluid links, ref:588-641). */
int lbm_get_force(LbmHandle h, float out_xy[2]);
/* get_max_velocity(), ref:656-660; NaN in the field gives NaN. */
int lbm_get_max_velocity(LbmHandle h, float *out);
/* vel.to_numpy() / get_physical_fields(), ref:207-212: (nx,ny,2). */
int lbm_get_vel(LbmHandle h, float *out_nxny2);
/* rho.to_numpy(): (nx,ny). */
int lbm_get_rho(LbmHandle h, float *out_nxny);
/* mask.to_numpy(): (nx,ny) float32 0/1, ref:107-111. */
int lbm_get_mask(LbmHandle h, float *out_nxny);
/* get_moments_numpy(), ref:739-741 (= compute_moments_for_output, ref:667-737): (nx,ny,9),
 * channel order [rho, e, eps, jx, qx, jy, qy, pxx, pxy], taken from the reference's f_new:
 * post-collision values at interior cells (solids included), the initial equilibrium on the ring. */
int lbm_get_moments(LbmHandle h, float *out_nxny9);
/* The numeric part of the reference's video frame, Taichi_Gui_Viz.process_frame
 * (src/lbm_mrt_les/visualization/Taichi_Gui_Viz.py:22-34), computed on the device instead of on the host from
 * get_physical_fields(): scipy.ndimage.gaussian_filter of both velocity components (restated operation for
 * operation: separable, axis 0 then 1, double accumulation, mode "reflect", float32 between the passes),
 * |u| and the np.gradient vorticity.  weights[0..radius] = the taps of scipy's kernel at distance 0..radius
 * (float64, computed by the caller the way scipy does); radius = 0 / weights = NULL = no filter (viz_sigma <= 0).
 * out_mag, out_vor: (nx, ny) float32 each.  Single GPU (the filter reaches across slab borders). */
int lbm_get_viz_fields(LbmHandle h, const double *weights, int radius, float *out_mag_nxny, float *out_vor_nxny);
/* f_old.to_numpy() (which = 0) / f_new.to_numpy() (which = 1): (nx,ny,9).  Parity / debugging. */
int lbm_get_f(LbmHandle h, int which, float *out_nxny9);

/* ---- on-device export reduction: io/lbm_writer.py:135-251 of the reference -------------------------
 * lbm_export_configure(): ROI [x0,x1) x [y0,y1) in grid cells (the writer's crop, lbm_writer.py:37-42) and the
 * target size (target_w, target_h) it derives from save_resolution_height (lbm_writer.py:52-58); resets
 * the running statistics.  lbm_export_frame(): the 9 moments of get_moments_numpy() over the ROI,
 * down-sampled per channel exactly like cv2.resize(..., INTER_AREA) (lbm_writer.py:150-163), accumulated
 * into the running sum / min / max / sum(u^2+v^2) / sum|vorticity| (lbm_writer.py:176-210) on the device;
 * `out_chw` (9, target_h, target_w) fp32 receives the frame (may be NULL).  lbm_export_stats(): the
 * accumulators for finalize() (lbm_writer.py:212-251); any pointer may be NULL. */
typedef struct {
    int32_t x0, x1, y0, y1;
    int32_t target_w, target_h;
} LbmExportConfig;
int lbm_export_configure(LbmHandle h, const LbmExportConfig *cfg);
/* x-slabs: ROI and target are GLOBAL; every rank calls the export functions collectively and holds the output
 * columns [dlo, dhi) of the global (9, target_h, target_w) frame (those whose first source column it owns), so
 * the concatenation over ranks is the single-GPU / cv2 result.  lbm_export_frame() then fills (9, target_h, dhi-dlo). */
int lbm_export_layout(LbmHandle h, int32_t *dlo, int32_t *dhi, int32_t *target_h);
int lbm_export_frame(LbmHandle h, float *out_chw);
/* The same frame left on the device: *frame_chw_dev points at the handle's (9, target_h, dhi-dlo) float32 buffer (NULL
 * when this rank holds no output column), complete when the call returns and valid until the next export call.  For the
 * x-slab writer: the ranks' column ranges are gathered GPU to GPU (NCCL) and only rank 0 copies the frame to the host. */
int lbm_export_frame_device(LbmHandle h, const float **frame_chw_dev);
int lbm_export_stats(LbmHandle h, double *running_sum_chw, double *vel_sq_sum_hw, double *abs_vor_sum_hw,
                     double *min9, double *max9, int64_t *count);

/* ---- multi-GPU x-slabs: one process per GPU, one handle per slab ---------------
 * Every rank creates its handle with nx_global / slab_x0 set (slabs in rank order, west to east), then all
 * ranks call lbm_comm_connect() collectively with the id rank 0 obtained from lbm_comm_unique_id() (moved
 * between processes by the host, e.g. torch.distributed broadcast).  From then on lbm_run() exchanges one
 * halo column per interface and step over NCCL (NVLink): the three populations that cross it in each
 * direction (f1,f5,f8 eastward, f3,f6,f7 westward), ny floats each.  Global force / max|u| are reduced by
 * the host (they are per-batch scalars). */
#define LBM_COMM_ID_BYTES 128
int lbm_comm_unique_id(uint8_t out[LBM_COMM_ID_BYTES]);
int lbm_comm_connect(LbmHandle h, int rank, int nranks, const uint8_t id[LBM_COMM_ID_BYTES]);

/* `static_mask` of the case file, io/lbm_writer.py:74-110: the ROI [x0,x1) x [y0,y1) of the obstacle mask, transposed to
 * image order and resized to (target_h, target_w) with cv2.INTER_NEAREST, and its signed Euclidean distance field
 * (scipy distance_transform_edt of the fluid minus that of the solid: positive in the fluid) -- out is (2, th, tw)
 * float32, bit-identical to the reference's cv2 + scipy result.  *degenerate = 1 (out untouched) when the resized mask
 * has no solid or no fluid pixel; the binding then falls back to cv2 + scipy.  Single GPU. */
int lbm_static_mask(LbmHandle h, int32_t x0, int32_t x1, int32_t y0, int32_t y1, int32_t target_w, int32_t target_h,
                    float *out, int32_t *degenerate);

/* Peer-memory halo path (preferred on NVLink / NVSwitch nodes): after lbm_comm_connect(), every rank exports CUDA IPC
 * handles of its two population buffers and its inbox counters (lbm_peer_export), the host exchanges the blobs
 * (torch.distributed all_gather) and each rank maps its neighbours' buffers (lbm_peer_connect; NULL on a side that is a
 * domain boundary).  From then on a slab step is ONE kernel launch: the CTAs of the first / last owned column store the
 * populations that cross the interface straight into the neighbour's halo column over NVLink and publish a per-step
 * counter; the neighbour's edge CTAs of the next step wait on it (bounded; a timeout surfaces as LBM_ERR_NCCL from the
 * next getter).  No NCCL call is left on the per-step path, and programmatic dependent launch + early start work as on
 * one GPU.  If lbm_peer_connect() fails (no peer access) the handle keeps the NCCL exchange above.  All ranks must
 * call lbm_init() and reach a host barrier before any of them steps (slab.py does), and must not destroy a handle while a
 * neighbour may still be stepping. */
#define LBM_PEER_HANDLE_BYTES 256
int lbm_peer_export(LbmHandle h, uint8_t out[LBM_PEER_HANDLE_BYTES]);
int lbm_peer_connect(LbmHandle h, const uint8_t *west_blob, const uint8_t *east_blob);

/* ---- HBM-resident access (no host copies): device pointers valid until lbm_destroy ---------- */
typedef struct {
    float *f_cur;      /* 9 planes, plane stride `plane_stride` floats, (nx_local, pitch) y fastest */
    float *f_prev;
    float *rho, *ux, *uy;
    uint8_t *cell_code;
    int32_t nx_local, ny, pitch;
    int64_t plane_stride;
    void *stream;      /* cudaStream_t the handle launches on */
} LbmDeviceView;
int lbm_device_view(LbmHandle h, LbmDeviceView *out);

/* Kernel launches issued by this handle since creation (bench.py's gpu_launches). */
int lbm_launch_count(LbmHandle h, int64_t *launches);
/* lbm_run() calls that were replayed as one CUDA graph (launch-bound grids, see lbm_run); their kernels are in
 * lbm_launch_count() as well. */
int lbm_graph_replay_count(LbmHandle h, int64_t *replays);

/* Page-locked host memory for the large device -> host getters.  The reference hands out a FRESH (nx, ny, 9) numpy array
 * per export (ref:739-741; it is queued to the writer thread, io/lbm_writer.py:260-287), which as a pageable allocation
 * costs first-touch page faults plus a staged copy -- 3.5 GB/s instead of PCIe speed.  The Python binding therefore keeps
 * a small pool of these buffers and wraps one per call in a caller-owned array that returns it to the pool when it is
 * garbage collected; every getter DMAs straight into a destination allocated here. */
int lbm_host_alloc(size_t bytes, void **out);
int lbm_host_free(void *ptr);

/* Self test of the strict kernel's inline packed division / square root (two cells per FFMA2 pair) against CUDA's
 * correctly rounded __fdiv_rn / __fsqrt_rn on `pairs` random operand pairs from the range the kernel admits to the
 * inline sequences; mismatches[0..2] = a / b, 1 / t, sqrt(x).  No reference counterpart: it guards the bit-exactness
 * claim of LBM_ARITH_STRICT against ref:281-284, 348-356. */
int lbm_selftest_arith(int64_t pairs, uint64_t seed, int64_t mismatches[3]);

#ifdef __cplusplus
}
#endif
#endif /* LBM2D_H_ */
