// Device-side building blocks of the fused D2Q9 MRT-LES step (sm_100a).
//
// Written from the algorithm description in SURVEY.md section 3.4, not from the Taichi kernels: SoA
// planes, one pass f_src -> f_dst, boundary ring and obstacle refill fused into the same pass.
// `ref:LINE` cites /root/reference/src/lbm_mrt_les/core/LBM2D_MRT_LES.py for parity checking.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lbm {

// D2Q9 directions, ref:146-158: 0 rest; 1 E; 2 N; 3 W; 4 S; 5 NE; 6 NW; 7 SW; 8 SE
__device__ constexpr int kEx[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
__device__ constexpr int kEy[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
// Lallemand-Luo rows, ref:167-180; moment order [rho, e, eps, jx, qx, jy, qy, pxx, pxy]
__device__ constexpr int kM[9][9] = {
    {1, 1, 1, 1, 1, 1, 1, 1, 1},      {-4, -1, -1, -1, -1, 2, 2, 2, 2}, {4, -2, -2, -2, -2, 1, 1, 1, 1},
    {0, 1, 0, -1, 0, 1, -1, -1, 1},   {0, -2, 0, 2, 0, 1, -1, -1, 1},   {0, 0, 1, 0, -1, 1, 1, -1, -1},
    {0, 0, -2, 0, 2, 1, 1, -1, -1},   {0, 1, -1, 1, -1, 0, 0, 0, 0},    {0, 0, 0, 0, 0, 1, -1, 1, -1}};
// ||row||^2 of M: M^-1 = M^T diag(1/norm)  (rows are orthogonal); ref:182 computes it numerically,
// its non-zero entries are exactly these correctly rounded fractions.
__device__ constexpr double kMNorm[9] = {9, 36, 36, 6, 12, 6, 12, 4, 4};
__device__ constexpr float kW[9] = {(float)(4.0 / 9.0),  (float)(1.0 / 9.0),  (float)(1.0 / 9.0),
                                    (float)(1.0 / 9.0),  (float)(1.0 / 9.0),  (float)(1.0 / 36.0),
                                    (float)(1.0 / 36.0), (float)(1.0 / 36.0), (float)(1.0 / 36.0)};

__device__ __forceinline__ constexpr float inv_m(int r, int c) { return (float)((double)kM[c][r] / kMNorm[c]); }

// ---------------------------------------------------------------------------------------------
// Arithmetic policies.  Strict: every operation individually rounded, never contracted -- the
// reference's left-to-right order gives results bit-identical to the fp32 oracle.  Fast: plain
// operators (nvcc contracts to FFMA), reciprocal / rsqrt through the SFU.
// ---------------------------------------------------------------------------------------------
struct Strict {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
};

// SFU square root / reciprocal for the fast flavour (MUFU.SQRT / MUFU.RCP, <= 2 ulp / 1 ulp; the
// .ftz forms avoid the denormal rescaling sequences -- every operand here is O(1e-9..10)).
__device__ __forceinline__ float fast_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Parameters every cell needs (kernel argument, lives in the constant bank).
struct Physics {
    float tau0;        // f32(3 nu + 0.5)                      ref:44
    float tau0_sq;     // f32(tau0_f64^2)                      ref:348
    float cs_factor;   // f32(18 Cs^2)                         ref:79
    float s_ghost;     // ghost_moments_s                      ref:82
    int les_on;        // C_smag > 0.001                       ref:342
    float rho_in, rho_out;
    int bc_type[4];
    float bc_val[4][2];
    int nx_global;     // ibc == nx-1 test, ref:495
};

// ---------------------------------------------------------------------------------------------
// Collision, strict flavour: literal restatement of ref:266-420 for one cell.
// in: pulled populations f[9], damping = max(damp_x, damp_y).  out: post-collision g[9].
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void collide_strict(const Physics &P, const float (&f)[9], float damp, float (&g)[9]) {
    using A = Strict;
    float m[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        float val = 0.0f;
#pragma unroll
        for (int c = 0; c < 9; ++c) val = A::add(val, A::mul((float)kM[r][c], f[c]));
        m[r] = val;
    }
    const float rho = m[0];
    float u = 0.0f, v = 0.0f;
    if (rho > 0.0f) {
        u = A::div(m[3], rho);
        v = A::div(m[5], rho);
    }
    const float u2 = A::add(A::mul(u, u), A::mul(v, v));
    float meq[9];
    meq[0] = rho;
    meq[1] = A::mul(rho, A::add(-2.0f, A::mul(3.0f, u2)));
    meq[2] = A::mul(rho, A::sub(1.0f, A::mul(3.0f, u2)));
    meq[3] = A::mul(rho, u);
    meq[4] = A::mul(-rho, u);
    meq[5] = A::mul(rho, v);
    meq[6] = A::mul(-rho, v);
    meq[7] = A::mul(rho, A::sub(A::mul(u, u), A::mul(v, v)));
    meq[8] = A::mul(A::mul(rho, u), v);
    const float n7 = A::sub(m[7], meq[7]);
    const float n8 = A::sub(m[8], meq[8]);
    const float norm = A::sqrt(A::add(A::mul(A::mul(2.0f, n7), n7), A::mul(A::mul(2.0f, n8), n8)));
    float tau_eff = P.tau0;
    if (P.les_on) {
        const float term = A::add(P.tau0_sq, A::div(A::mul(P.cs_factor, norm), rho));
        const float tau_eddy = A::mul(0.5f, A::sub(A::sqrt(term), P.tau0));
        tau_eff = A::add(P.tau0, tau_eddy);
    }
    tau_eff = A::add(tau_eff, damp);
    const float s_eff = A::div(1.0f, tau_eff);
    const float S[9] = {0.0f, P.s_ghost, P.s_ghost, 0.0f, P.s_ghost, 0.0f, P.s_ghost, s_eff, s_eff};
    float ms[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) ms[k] = A::sub(m[k], A::mul(S[k], A::sub(m[k], meq[k])));
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        float val = 0.0f;
#pragma unroll
        for (int c = 0; c < 9; ++c) val = A::add(val, A::mul(inv_m(r, c), ms[c]));
        g[r] = val;
    }
}

// ---------------------------------------------------------------------------------------------
// Collision, fast flavour: same mathematics, sparse integer transforms with shared partial sums,
// conserved moments passed through, SFU reciprocal / sqrt.  ~95 FP instructions per cell.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void collide_fast(const Physics &P, const float (&f)[9], float damp, float (&g)[9]) {
    const float a13 = f[1] + f[3], a24 = f[2] + f[4], a57 = f[5] + f[7], a68 = f[6] + f[8];
    const float d13 = f[1] - f[3], d24 = f[2] - f[4], p = f[5] - f[7], q = f[6] - f[8];
    const float s1 = a13 + a24, s2 = a57 + a68;
    const float rho = f[0] + s1 + s2;
    const float e = 2.0f * s2 - s1 - 4.0f * f[0];
    const float eps = 4.0f * f[0] - 2.0f * s1 + s2;
    const float pq = p - q, pp = p + q;
    const float jx = d13 + pq, qx = pq - 2.0f * d13;
    const float jy = d24 + pp, qy = pp - 2.0f * d24;
    const float pxx = a13 - a24, pxy = a57 - a68;

    const float inv_rho = (rho > 0.0f) ? fast_rcp(rho) : 0.0f;     // ref:281-284: u = v = 0 if rho <= 0
    const float jx2 = jx * jx, jy2 = jy * jy;
    const float ru2 = (jx2 + jy2) * inv_rho;                      // rho * (u^2 + v^2)
    const float meq1 = 3.0f * ru2 - 2.0f * rho;
    const float meq2 = rho - 3.0f * ru2;
    const float meq7 = (jx2 - jy2) * inv_rho;
    const float meq8 = jx * jy * inv_rho;
    const float n7 = pxx - meq7, n8 = pxy - meq8;
    float tau_eff = P.tau0;
    if (P.les_on) {
        const float norm = fast_sqrt(2.0f * (n7 * n7 + n8 * n8));
        // NB: the reference divides by rho_l itself (inf / nan if rho <= 0), ref:348
        const float term = P.tau0_sq + P.cs_factor * norm * fast_rcp(rho);
        tau_eff = 0.5f * (P.tau0 + fast_sqrt(term));
    }
    tau_eff += damp;
    const float s_eff = fast_rcp(tau_eff);
    const float sg = P.s_ghost;
    // relaxed moments already scaled by 1/||row||^2 for the inverse transform
    const float r0 = rho * (float)(1.0 / 9.0);
    const float r1 = (e - sg * (e - meq1)) * (float)(1.0 / 36.0);
    const float r2 = (eps - sg * (eps - meq2)) * (float)(1.0 / 36.0);
    const float r3 = jx * (float)(1.0 / 6.0);
    const float r4 = (qx - sg * (qx + jx)) * (float)(1.0 / 12.0);
    const float r5 = jy * (float)(1.0 / 6.0);
    const float r6 = (qy - sg * (qy + jy)) * (float)(1.0 / 12.0);
    const float r7 = (pxx - s_eff * n7) * 0.25f;
    const float r8 = (pxy - s_eff * n8) * 0.25f;
    const float A = r0 - r1 - 2.0f * r2, B = r0 + 2.0f * r1 + r2;
    const float X = r3 - 2.0f * r4, Y = r5 - 2.0f * r6, Xd = r3 + r4, Yd = r5 + r6;
    g[0] = r0 - 4.0f * r1 + 4.0f * r2;
    const float Ap = A + r7, Am = A - r7;
    g[1] = Ap + X;
    g[3] = Ap - X;
    g[2] = Am + Y;
    g[4] = Am - Y;
    const float Bp = B + r8, Bm = B - r8;
    const float XpY = Xd + Yd, XmY = Xd - Yd;
    g[5] = Bp + XpY;
    g[7] = Bp - XpY;
    g[8] = Bm + XmY;
    g[6] = Bm - XmY;
}

// Macroscopic values from post-collision populations, ref:425-436 (sequential sums).
template <bool STRICT>
__device__ __forceinline__ void macro_from_f(const float (&g)[9], float &rho, float &ux, float &uy) {
    if (STRICT) {
        using A = Strict;
        float lr = 0.0f, lx = 0.0f, ly = 0.0f;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            lr = A::add(lr, g[k]);
            lx = A::add(lx, A::mul((float)kEx[k], g[k]));
            ly = A::add(ly, A::mul((float)kEy[k], g[k]));
        }
        rho = lr;
        if (lr > 0.0f) {
            ux = A::div(lx, lr);
            uy = A::div(ly, lr);
        } else {
            ux = 0.0f;
            uy = 0.0f;
        }
    } else {
        const float lr = g[0] + g[1] + g[2] + g[3] + g[4] + g[5] + g[6] + g[7] + g[8];
        const float lx = g[1] - g[3] + g[5] - g[6] - g[7] + g[8];
        const float ly = g[2] - g[4] + g[5] + g[6] - g[7] - g[8];
        const float inv = (lr > 0.0f) ? fast_rcp(lr) : 0.0f;
        rho = lr;
        ux = lx * inv;
        uy = ly * inv;
    }
}

// ---------------------------------------------------------------------------------------------
// Boundary ring.  Always strict arithmetic: O(perimeter) work, and every branch of the
// reference's apply_bc_core (ref:457-550) is mirrored including its coordinate quirks.
// ---------------------------------------------------------------------------------------------
struct Cell {
    float f[9];
    float rho, ux, uy;
};

__device__ __forceinline__ void f_eq_strict(float rho, float ux, float uy, float (&out)[9]) {  // ref:214-218
    using A = Strict;
    const float uv = A::add(A::mul(ux, ux), A::mul(uy, uy));
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float eu = A::add(A::mul((float)kEx[k], ux), A::mul((float)kEy[k], uy));
        float t = A::add(1.0f, A::mul(3.0f, eu));
        t = A::add(t, A::mul(A::mul(4.5f, eu), eu));
        t = A::sub(t, A::mul(1.5f, uv));
        out[k] = A::mul(A::mul(kW[k], rho), t);
    }
}

// State of a ring cell the reference never writes (no-op boundary types): its init value.
__device__ __forceinline__ void cell_rest(Cell &c) {
#pragma unroll
    for (int k = 0; k < 9; ++k) c.f[k] = kW[k];
    c.rho = 1.0f;
    c.ux = 0.0f;
    c.uy = 0.0f;
}

// bc <- apply_bc_core(dr, ibc, ., inb, .) given the neighbour's fresh (pre-refill) state.
// `ibc`, `inb` are GLOBAL x coordinates; `bc` must hold the rest state on entry.
__device__ __forceinline__ void bc_core(const Physics &P, int dr, int ibc, int inb, const Cell &nb, Cell &bc, float ramp) {
    using A = Strict;
    const int t = P.bc_type[dr];
    float eb[9], en[9];
    if (t == 0) {
        if (ibc == 0) {  // Zou-He pressure inlet, ref:461-486
            const float rc = A::add(1.0f, A::mul(A::sub(P.rho_in, 1.0f), ramp));
            const float f0 = nb.f[0], f2 = nb.f[2], f3 = nb.f[3], f4 = nb.f[4], f6 = nb.f[6], f7 = nb.f[7];
            const float s = A::add(A::add(A::add(f0, f2), f4), A::mul(2.0f, A::add(A::add(f3, f6), f7)));
            const float ux = A::sub(1.0f, A::div(s, rc));
            const float c23 = A::mul(A::mul((float)(2.0 / 3.0), rc), ux);
            const float c16 = A::mul(A::mul((float)(1.0 / 6.0), rc), ux);
            const float h = A::mul(0.5f, A::sub(f2, f4));
            bc.rho = rc;
            bc.ux = ux;
            bc.uy = 0.0f;
            f_eq_strict(rc, ux, 0.0f, bc.f);
            bc.f[1] = A::add(f3, c23);
            bc.f[5] = A::add(A::sub(f7, h), c16);
            bc.f[8] = A::add(A::add(f6, h), c16);
        } else {  // velocity Dirichlet by non-equilibrium extrapolation, ref:487-492
            bc.ux = A::mul(P.bc_val[dr][0], ramp);
            bc.uy = A::mul(P.bc_val[dr][1], ramp);
            bc.rho = nb.rho;
            f_eq_strict(bc.rho, bc.ux, bc.uy, eb);
            f_eq_strict(nb.rho, nb.ux, nb.uy, en);
#pragma unroll
            for (int k = 0; k < 9; ++k) bc.f[k] = A::add(A::sub(eb[k], en[k]), nb.f[k]);
        }
    } else if (t == 1) {
        if (ibc == P.nx_global - 1) {  // Zou-He pressure outlet, ref:495-527
            const float ro = P.rho_out;
            const float f0 = nb.f[0], f1 = nb.f[1], f2 = nb.f[2], f4 = nb.f[4], f5 = nb.f[5], f8 = nb.f[8];
            const float s = A::add(A::add(A::add(f0, f2), f4), A::mul(2.0f, A::add(A::add(f1, f5), f8)));
            const float ux = A::add(-1.0f, A::div(s, ro));
            if (ux < 0.0f) {  // backflow guard, ref:508-516
                bc.ux = nb.ux;
                bc.uy = nb.uy;
                bc.rho = ro;
                f_eq_strict(bc.rho, bc.ux, bc.uy, eb);
                f_eq_strict(nb.rho, nb.ux, nb.uy, en);
#pragma unroll
                for (int k = 0; k < 9; ++k) bc.f[k] = A::add(A::sub(eb[k], en[k]), nb.f[k]);
            } else {
                const float c23 = A::mul(A::mul((float)(2.0 / 3.0), ro), ux);
                const float c16 = A::mul(A::mul((float)(1.0 / 6.0), ro), ux);
                const float h = A::mul(0.5f, A::sub(f2, f4));
                bc.rho = ro;
                bc.ux = ux;
                bc.uy = 0.0f;
                f_eq_strict(ro, ux, 0.0f, bc.f);
                bc.f[3] = A::sub(f1, c23);
                bc.f[6] = A::sub(A::sub(f8, h), c16);
                bc.f[7] = A::sub(A::add(f5, h), c16);
            }
        }
        // type 1 anywhere else: the reference does nothing -> rest state
    } else if (t == 2) {  // free slip, ref:529-550
        if (ibc == inb) {
            bc.ux = nb.ux;
            bc.uy = 0.0f;
        } else {
            bc.ux = 0.0f;
            bc.uy = nb.uy;
        }
        bc.rho = nb.rho;
        f_eq_strict(bc.rho, bc.ux, bc.uy, eb);
        f_eq_strict(nb.rho, nb.ux, nb.uy, en);
#pragma unroll
        for (int k = 0; k < 9; ++k) bc.f[k] = A::add(A::sub(eb[k], en[k]), nb.f[k]);
    }
    // any other type (3 = "no-slip" is documented but not implemented in the reference): rest state
}

// Obstacle refill, ref:452-455: u = 0, f = f_eq(rho, 0) = (w_k rho) * 1.
__device__ __forceinline__ void refill(Cell &c) {
    c.ux = 0.0f;
    c.uy = 0.0f;
#pragma unroll
    for (int k = 0; k < 9; ++k) c.f[k] = __fmul_rn(kW[k], c.rho);
}

// 9 moments by the reference's hand-expanded rows, ref:682-737 (strict, export path only).
__device__ __forceinline__ void moments_strict(const float (&f)[9], float (&o)[9]) {
    using A = Strict;
    float rho = 0.0f;
#pragma unroll
    for (int k = 0; k < 9; ++k) rho = A::add(rho, f[k]);
    const float s14 = A::add(A::add(A::add(f[1], f[2]), f[3]), f[4]);
    const float s58 = A::add(A::add(A::add(f[5], f[6]), f[7]), f[8]);
    o[0] = rho;
    o[1] = A::add(A::sub(A::mul(-4.0f, f[0]), s14), A::mul(2.0f, s58));
    o[2] = A::add(A::sub(A::mul(4.0f, f[0]), A::mul(2.0f, s14)), s58);
    const float t56 = A::sub(f[5], f[6]);
    o[3] = A::add(A::sub(A::sub(A::add(A::sub(f[1], f[3]), f[5]), f[6]), f[7]), f[8]);
    o[4] = A::add(A::sub(A::sub(A::add(A::add(A::mul(-2.0f, f[1]), A::mul(2.0f, f[3])), f[5]), f[6]), f[7]), f[8]);
    o[5] = A::sub(A::sub(A::add(A::add(A::sub(f[2], f[4]), f[5]), f[6]), f[7]), f[8]);
    o[6] = A::sub(A::sub(A::add(A::add(A::add(A::mul(-2.0f, f[2]), A::mul(2.0f, f[4])), f[5]), f[6]), f[7]), f[8]);
    o[7] = A::sub(A::add(A::sub(f[1], f[2]), f[3]), f[4]);
    o[8] = A::sub(A::add(t56, f[7]), f[8]);
}

// ---------------------------------------------------------------------------------------------
// Ring production (rare path, O(perimeter) cells per step).
//
// A ring cell is a function of ONE adjacent interior cell's fresh, un-refilled state,
// so the thread that has just collided interior cell (il, j) -- the "owner" -- also produces the
// ring cells hanging off it: W/E cell if it sits in column 1 / nx-2 (ref:445-447), top/bottom cell if
// j == ny-2 / 1 (ref:448-450), and the corner through the W/E cell just produced.  One out-of-line
// function with everything it needs behind a pointer in global memory, so the hot path keeps its
// state in registers and never materialises the kernel parameters on the stack.
// ---------------------------------------------------------------------------------------------
struct RingCtx {
    Physics phys;
    float *dst;              // 9 planes of the destination buffer
    float *rho, *ux, *uy;    // macroscopic planes (EMIT steps)
    const uint8_t *code;
    long long plane;
    int nx_local, ny, pitch;
    int x_off;               // global x of local column 0
    int west_ring, east_ring;
};

// Where a produced ring cell goes: the CTA's shared-memory output tile if it lies inside the tile's
// TMA store box, else straight to global memory with scalar stores.
struct TileSink {
    float *sm_f;             // [9][bx][by] or nullptr (no tile: always global)
    float *sm_mac;           // [3][bx][by] (EMIT) or nullptr
    int il0, j0;             // tile origin (local column, row)
    int bx, by;
    int row_hi, col_lo, col_hi;  // extent of the store tensor: rows [0,row_hi), local columns [col_lo,col_hi)
};

__device__ __forceinline__ void sink_put(const RingCtx &c, const TileSink &t, bool emit, int il, int j, Cell &v,
                                         float &vmax, bool &vnan) {
    if (c.code[(long long)il * c.pitch + j] & 1) refill(v);  // ref:452-455 also hits solid ring cells
    const int tx = il - t.il0, ty = j - t.j0;
    if (t.sm_f != nullptr && tx >= 0 && tx < t.bx && ty >= 0 && ty < t.by && j < t.row_hi && il >= t.col_lo && il < t.col_hi) {
        const int o = tx * t.by + ty, n = t.bx * t.by;
#pragma unroll
        for (int k = 0; k < 9; ++k) t.sm_f[k * n + o] = v.f[k];
        if (emit) {
            t.sm_mac[o] = v.rho;
            t.sm_mac[n + o] = v.ux;
            t.sm_mac[2 * n + o] = v.uy;
        }
    } else {
        const long long o = (long long)il * c.pitch + j;
#pragma unroll
        for (int k = 0; k < 9; ++k) c.dst[k * c.plane + o] = v.f[k];
        if (emit) {
            c.rho[o] = v.rho;
            c.ux[o] = v.ux;
            c.uy[o] = v.uy;
        }
    }
    if (emit) {
        const float m2 = __fadd_rn(__fmul_rn(v.ux, v.ux), __fmul_rn(v.uy, v.uy));
        vnan |= (m2 != m2);
        vmax = fmaxf(vmax, m2);
    }
}

// `me`: fresh un-refilled state of interior cell (il, j).  Returns max |u|^2 / NaN flag of what it wrote.
__device__ __noinline__ void ring_from_owner(const RingCtx *cp, const TileSink *tp, int emit, int il, int j,
                                             const Cell *mep, float ramp, float *vmax_io, int *vnan_io) {
    const RingCtx &c = *cp;
    const TileSink t = *tp;
    const Cell &me = *mep;
    float vmax = *vmax_io;
    bool vnan = *vnan_io != 0;
    const int ny = c.ny;
    const bool bottom = (j == 1), top = (j == ny - 2);
    const int ig = c.x_off + il;
#pragma unroll 1
    for (int side = 0; side < 2; ++side) {
        const bool on = side == 0 ? (il == 1 && c.west_ring) : (il == c.nx_local - 2 && c.east_ring);
        if (!on) continue;
        const int ilr = side == 0 ? 0 : c.nx_local - 1;
        const int igr = side == 0 ? ig - 1 : ig + 1;
        Cell r;
        cell_rest(r);
        bc_core(c.phys, side == 0 ? 0 : 2, igr, ig, me, r, ramp);
        if (top) {  // corner chains through the W/E cell just produced, un-refilled
            Cell cr;
            cell_rest(cr);
            bc_core(c.phys, 1, igr, igr, r, cr, ramp);
            sink_put(c, t, emit, ilr, ny - 1, cr, vmax, vnan);
        }
        if (bottom) {
            Cell cr;
            cell_rest(cr);
            bc_core(c.phys, 3, igr, igr, r, cr, ramp);
            sink_put(c, t, emit, ilr, 0, cr, vmax, vnan);
        }
        sink_put(c, t, emit, ilr, j, r, vmax, vnan);
    }
    if (top) {
        Cell r;
        cell_rest(r);
        bc_core(c.phys, 1, ig, ig, me, r, ramp);
        sink_put(c, t, emit, il, ny - 1, r, vmax, vnan);
    }
    if (bottom) {
        Cell r;
        cell_rest(r);
        bc_core(c.phys, 3, ig, ig, me, r, ramp);
        sink_put(c, t, emit, il, 0, r, vmax, vnan);
    }
    *vmax_io = vmax;
    *vnan_io = vnan;
}

}  // namespace lbm
