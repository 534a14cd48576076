// On-device export reduction: what the reference's HDF5 writer does on the
// host with the full (nx, ny, 9) moments frame -- ROI crop, per-channel cv2.INTER_AREA down-sampling
// and the running statistics of io/lbm_writer.py:135-210 (cited as writer:LINE) -- done on the GPU so that
// only the (9, H, W) export frame crosses PCIe (604 MB -> ~11 MB per frame at 8192x2048).
//
// The resize restates OpenCV's INTER_AREA for single-channel float32 shrinking operation for operation
// (strict fp32, same summation order: see oracle/writer_oracle.py, which is tested bit-for-bit against
// cv2), one thread per output pixel walking its slice of the x / y area tables.
#pragma once
#include "lbm2d_kernels.cuh"

namespace lbm {

struct AreaEntry {
    int si;       // source index
    float alpha;  // weight (float, as OpenCV's DecimateAlpha)
};

// Geometry of one rank's share of the export.  Single GPU: own_cols = cw = crop width, dlo = 0, dhi = tw_g.
// x-slabs: a rank computes the moments of its own ROI columns, receives the few columns its last output
// pixels reach into from the east neighbour, and produces the output columns [dlo, dhi) whose FIRST source
// column it owns -- so the assembled frame is the global cv2 result, independent of the decomposition.
struct ExportGeom {
    int x0, y0;          // first own ROI column (local index), first ROI row
    int own_cols;        // ROI columns computed locally
    int cw, ch;          // columns held in tmp (own_cols + received extension), rows
    int src_shift;       // global ROI-relative source column  -  src_shift  =  tmp column
    int tw_g, th;        // GLOBAL target width, target height
    int dlo, dhi;        // output columns of this rank
    int fast;            // both scales integer: resizeAreaFast_ path
    int ix, iy;          // integer scales (fast path)
};

// 9 moments of the reference's f_new over the ROI -> tmp[c][x][y] (y fastest)
__global__ void roi_moments_kernel(const ExportArgs a, ExportGeom g, float *__restrict__ tmp) {
    const int y = blockIdx.y * blockDim.x + threadIdx.x;
    const int x = blockIdx.x;   // < own_cols (grid x: no 65 535 limit on the number of columns)
    if (y >= g.ch) return;
    float f[9], m[9];
    load_f_new(a, g.x0 + x, g.y0 + y, f);
    moments_strict(f, m);
    const long long n = (long long)g.cw * g.ch, o = (long long)x * g.ch + y;
#pragma unroll
    for (int k = 0; k < 9; ++k) tmp[k * n + o] = m[k];
}

// ResizeArea_Invoker: out[c][dy][dx] = sum_j beta_j * (sum_k alpha_k * S[sy_j][sx_k]), every product
// and sum individually rounded, in table order.
__global__ void area_resize_kernel(const float *__restrict__ tmp, ExportGeom g, const AreaEntry *__restrict__ xtab,
                                   const int *__restrict__ xoff, const AreaEntry *__restrict__ ytab,
                                   const int *__restrict__ yoff, float *__restrict__ out) {
    const int dy = blockIdx.y * blockDim.x + threadIdx.x;  // lanes along y: neighbouring source rows
    const int dxl = blockIdx.x, dx = g.dlo + dxl, c = blockIdx.z;
    if (dy >= g.th) return;
    const float *S = tmp + (long long)c * g.cw * g.ch;
    float total = 0.0f;
    const int k0 = xoff[dx], k1 = xoff[dx + 1];
    for (int j = yoff[dy]; j < yoff[dy + 1]; ++j) {
        const int sy = ytab[j].si;
        float buf = 0.0f;
        for (int k = k0; k < k1; ++k)
            buf = __fadd_rn(buf, __fmul_rn(S[(long long)(xtab[k].si - g.src_shift) * g.ch + sy], xtab[k].alpha));
        total = __fadd_rn(total, __fmul_rn(ytab[j].alpha, buf));
    }
    out[((long long)c * g.th + dy) * (g.dhi - g.dlo) + dxl] = total;
}

// resizeAreaFast_: integer scales.  2x2: ((a+b)+(c+d))*0.25 (the SIMD kernel); otherwise the scalar loop
// unrolled by four the way OpenCV writes it, times 1/area.
__global__ void area_fast_kernel(const float *__restrict__ tmp, ExportGeom g, float *__restrict__ out) {
    const int dy = blockIdx.y * blockDim.x + threadIdx.x;
    const int dxl = blockIdx.x, dx = g.dlo + dxl, c = blockIdx.z;
    if (dy >= g.th) return;
    const float *S = tmp + (long long)c * g.cw * g.ch;
    auto at = [&](int sy, int sx) { return S[(long long)(sx - g.src_shift) * g.ch + sy]; };
    float r;
    if (g.ix == 2 && g.iy == 2) {
        const float a = at(2 * dy, 2 * dx), b = at(2 * dy, 2 * dx + 1), cc = at(2 * dy + 1, 2 * dx), d = at(2 * dy + 1, 2 * dx + 1);
        r = __fmul_rn(__fadd_rn(__fadd_rn(a, b), __fadd_rn(cc, d)), 0.25f);
    } else {
        const int area = g.ix * g.iy;
        float sum = 0.0f;
        int k = 0;
        auto val = [&](int q) { return at(dy * g.iy + q / g.ix, dx * g.ix + q % g.ix); };
        for (; k <= area - 4; k += 4)
            sum = __fadd_rn(sum, __fadd_rn(__fadd_rn(__fadd_rn(val(k), val(k + 1)), val(k + 2)), val(k + 3)));
        for (; k < area; ++k) sum = __fadd_rn(sum, val(k));
        r = __fmul_rn(sum, 1.0f / (float)area);
    }
    out[((long long)c * g.th + dy) * (g.dhi - g.dlo) + dxl] = r;
}

// Channels 0 (rho), 3 (jx), 5 (jy) of one frame column -> buf[3][th] (halo for the neighbour's x-gradient).
__global__ void export_pack_column_kernel(const float *__restrict__ frame, int twl, int th, int xl, float *__restrict__ buf) {
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= th) return;
    const long long n = (long long)twl * th, q = (long long)y * twl + xl;
    buf[y] = frame[q];
    buf[th + y] = frame[3 * n + q];
    buf[2 * th + y] = frame[5 * n + q];
}

// writer:176-210: running sum of the frame (float64), sum of u^2+v^2, sum of |vorticity| on the
// down-sampled grid (np.gradient: central differences inside, one-sided at the GLOBAL edges, float32).
// `left` / `right`: rho, jx, jy of the output columns dlo-1 / dhi held by the neighbouring ranks (or null).
__global__ void export_stats_kernel(const float *__restrict__ frame, int twl, int th, int dlo, int tw_g,
                                    const float *__restrict__ left, const float *__restrict__ right,
                                    double *__restrict__ running_sum, double *__restrict__ vel_sq_sum,
                                    double *__restrict__ abs_vor_sum) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= twl) return;
    const long long n = (long long)twl * th, o = (long long)y * twl + x;
#pragma unroll
    for (int c = 0; c < 9; ++c) running_sum[c * n + o] += (double)frame[c * n + o];
    auto uv = [&](int yy, int xx, float &u, float &v) {   // xx = -1 / twl read the neighbours' halo columns
        float rho, jx, jy;
        if (xx < 0) { rho = left[yy]; jx = left[th + yy]; jy = left[2 * th + yy]; }
        else if (xx >= twl) { rho = right[yy]; jx = right[th + yy]; jy = right[2 * th + yy]; }
        else {
            const long long q = (long long)yy * twl + xx;
            rho = frame[q]; jx = frame[3 * n + q]; jy = frame[5 * n + q];
        }
        const float rs = fmaxf(rho, 1e-6f);
        u = __fdiv_rn(jx, rs);
        v = __fdiv_rn(jy, rs);
    };
    float u, v;
    uv(y, x, u, v);
    vel_sq_sum[o] += (double)__fadd_rn(__fmul_rn(u, u), __fmul_rn(v, v));
    // dv/dx along W (axis 1), du/dy along H (axis 0)
    float ua, va, ub, vb, dvdx = 0.0f, dudy = 0.0f;
    const int xg = dlo + x;
    if (tw_g > 1) {
        const bool first = xg == 0, last = xg == tw_g - 1;
        uv(y, first ? x : x - 1, ua, va);
        uv(y, last ? x : x + 1, ub, vb);
        dvdx = (first || last) ? __fsub_rn(vb, va) : __fdiv_rn(__fsub_rn(vb, va), 2.0f);
    }
    if (th > 1) {
        const int ya = y == 0 ? 0 : y - 1, yb = y == th - 1 ? th - 1 : y + 1;
        uv(ya, x, ua, va);
        uv(yb, x, ub, vb);
        dudy = (y == 0 || y == th - 1) ? __fsub_rn(ub, ua) : __fdiv_rn(__fsub_rn(ub, ua), 2.0f);
    }
    abs_vor_sum[o] += (double)fabsf(__fsub_rn(dvdx, dudy));
}

// per-channel min / max of the frame folded into the running global min / max (writer:181-184)
__global__ void export_minmax_kernel(const float *__restrict__ frame, long long n, double *__restrict__ gmin, double *__restrict__ gmax) {
    const int c = blockIdx.x;
    float lo = INFINITY, hi = -INFINITY;
    bool nan = false;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = frame[c * n + i];
        nan |= (v != v);
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
    }
    __shared__ float slo[32], shi[32];
    __shared__ int snan;
    if (threadIdx.x == 0) snan = 0;
    __syncthreads();
    for (int s = 16; s > 0; s >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, s));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, s));
    }
    if (nan) atomicOr(&snan, 1);
    if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { lo = fminf(lo, slo[w]); hi = fmaxf(hi, shi[w]); }
        if (snan) { gmin[c] = NAN; gmax[c] = NAN; }   // np.minimum / np.maximum propagate NaN
        else {
            gmin[c] = fmin(gmin[c], (double)lo);
            gmax[c] = fmax(gmax[c], (double)hi);
        }
    }
}

// ---- video-frame fields (Taichi_Gui_Viz.process_frame, viz:22-34) ---------------------------------------------
// mode="reflect" of scipy.ndimage: (d c b a | a b c d | d c b a), any distance past the edge
__device__ __forceinline__ int reflect_index(int i, int n) {
    const int p = 2 * n;
    i %= p;
    if (i < 0) i += p;
    return i < n ? i : p - 1 - i;
}

// One pass of NI_Correlate1D (symmetric kernel): centre tap first, then the pairs from the outermost inwards,
// (in[-j] + in[+j]) * w[j], accumulated in double with every operation rounded on its own; float32 out.
// in: element (x, y) at x * sx + y; out: (nx, ny) unpitched.  axis 0 = along x, 1 = along y.  Two fields at once.
// x-slabs: row x of the OUTPUT is global column gout0 + x, row q of the INPUT is global column gin0 + q, and the
// reflection along x happens at the GLOBAL edges [0, nx_global) -- the input carries enough neighbour columns.
__global__ void viz_blur_kernel(const float *__restrict__ in0, const float *__restrict__ in1, long long sx, int nx, int ny,
                                int axis, int radius, const double *__restrict__ w, float *__restrict__ out0,
                                float *__restrict__ out1, int gout0, int gin0, int nx_global) {
    const int y = blockIdx.y * blockDim.x + threadIdx.x, x = blockIdx.x;
    if (y >= ny) return;
    const int pos = axis == 0 ? gout0 + x : y, n = axis == 0 ? nx_global : ny;
    auto at = [&](const float *in, int q) {
        return (double)(axis == 0 ? in[(long long)(q - gin0) * sx + y] : in[(long long)x * sx + q]);
    };
    double a0 = __dmul_rn(at(in0, pos), w[0]), a1 = __dmul_rn(at(in1, pos), w[0]);
    for (int j = radius; j > 0; --j) {
        const int lo = reflect_index(pos - j, n), hi = reflect_index(pos + j, n);
        a0 = __dadd_rn(a0, __dmul_rn(__dadd_rn(at(in0, lo), at(in0, hi)), w[j]));
        a1 = __dadd_rn(a1, __dmul_rn(__dadd_rn(at(in1, lo), at(in1, hi)), w[j]));
    }
    const long long o = (long long)x * ny + y;
    out0[o] = __double2float_rn(a0);
    out1[o] = __double2float_rn(a1);
}

// vel_mag = sqrt(vx^2 + vy^2); vor = np.gradient(vx)[1] - np.gradient(vy)[0] (float32: central differences / 2 inside,
// one-sided at the edges).  vx, vy: element (x, y) at x * sx + y, row 0 = global column gin0; output row x = global
// column gout0 + x (x-slabs: the input holds one neighbour column on every side that is not a global edge).
__global__ void viz_fields_kernel(const float *__restrict__ vx, const float *__restrict__ vy, long long sx, int nx, int ny,
                                  float *__restrict__ mag, float *__restrict__ vor, int gout0, int gin0, int nx_global) {
    const int y = blockIdx.y * blockDim.x + threadIdx.x, x = blockIdx.x;
    if (y >= ny) return;
    const int g = gout0 + x;
    auto U = [&](int gg, int yy) { return vx[(long long)(gg - gin0) * sx + yy]; };
    auto Vv = [&](int gg, int yy) { return vy[(long long)(gg - gin0) * sx + yy]; };
    const float u = U(g, y), v = Vv(g, y);
    const float dudy = (y == 0)        ? __fsub_rn(U(g, 1), U(g, 0))
                       : (y == ny - 1) ? __fsub_rn(U(g, ny - 1), U(g, ny - 2))
                                       : __fdiv_rn(__fsub_rn(U(g, y + 1), U(g, y - 1)), 2.0f);
    const float dvdx = (g == 0)               ? __fsub_rn(Vv(1, y), Vv(0, y))
                       : (g == nx_global - 1) ? __fsub_rn(Vv(nx_global - 1, y), Vv(nx_global - 2, y))
                                              : __fdiv_rn(__fsub_rn(Vv(g + 1, y), Vv(g - 1, y)), 2.0f);
    const long long o = (long long)x * ny + y;
    mag[o] = __fsqrt_rn(__fadd_rn(__fmul_rn(u, u), __fmul_rn(v, v)));
    vor[o] = __fsub_rn(dudy, dvdx);
}

// ---- static mask + signed distance field of the case file (io/lbm_writer.py:74-110) ------------------------------
// small[y][x] = mask(x0 + xs[x], y0 + ys[y]) -- cv2.INTER_NEAREST of the transposed ROI (index tables from the host, see
// lbm_static_mask); image order (th, tw).
__global__ void mask_nearest_kernel(const uint8_t *__restrict__ code, int pitch, const int *__restrict__ xs,
                                    const int *__restrict__ ys, int tw, int th, uint8_t *__restrict__ small, int *counts) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= tw) return;
    const uint8_t v = code[(long long)xs[x] * pitch + ys[y]] & 1;
    small[(long long)y * tw + x] = v;
    if (v) atomicAdd(counts, 1);
}
// Exact Euclidean distance transform, scipy.ndimage.distance_transform_edt semantics: for every pixel that is NOT
// background, the distance to the nearest background pixel (0 on the background).  Two separable passes on integers:
// per column the distance along y to the nearest background pixel of that column, then per pixel the minimum over the
// row of dx^2 + g^2; the square root is taken in double like scipy's.  bg = the value of `small` that is background.
__global__ void edt_columns_kernel(const uint8_t *__restrict__ small, int tw, int th, int bg, int *__restrict__ g) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= tw) return;
    const int kInf = 1 << 20;
    int d = kInf;
    for (int y = 0; y < th; ++y) {
        d = small[(long long)y * tw + x] == bg ? 0 : (d >= kInf ? kInf : d + 1);
        g[(long long)y * tw + x] = d;
    }
    d = kInf;
    for (int y = th - 1; y >= 0; --y) {
        d = small[(long long)y * tw + x] == bg ? 0 : (d >= kInf ? kInf : d + 1);
        if (d < g[(long long)y * tw + x]) g[(long long)y * tw + x] = d;
    }
}
__global__ void edt_rows_kernel(const int *__restrict__ g, int tw, int th, double *__restrict__ dist) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= tw) return;
    const int *row = g + (long long)y * tw;
    long long best = -1;
    for (int q = 0; q < tw; ++q) {
        const long long gy = row[q];
        if (gy >= (1 << 20)) continue;
        const long long dx = x - q, d2 = dx * dx + gy * gy;
        if (best < 0 || d2 < best) best = d2;
    }
    dist[(long long)y * tw + x] = sqrt((double)best);
}
// out (2, th, tw): channel 0 = mask, channel 1 = float32(dist_fluid - dist_solid)  (fluid positive)
__global__ void sdf_combine_kernel(const uint8_t *__restrict__ small, const double *__restrict__ d_fluid,
                                   const double *__restrict__ d_solid, long long n, float *__restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = small[i] ? 1.0f : 0.0f;
    out[n + i] = __double2float_rn(__dsub_rn(d_fluid[i], d_solid[i]));
}

}  // namespace lbm
