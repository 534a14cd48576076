/*
 * C restatement of the reference D2Q9 MRT-LES solver  --  TEST INFRASTRUCTURE ONLY.
 *
 * This is the parity checker and the CPU baseline ("port"), not the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * It follows /root/reference/src/lbm_mrt_les/core/LBM2D_MRT_LES.py (cited as ref:LINE) with the
 * reference's structure kept on purpose: three passes per step (collide_and_stream,
 * update_macro_var, apply_bc), AoS fields f[(i*ny+j)*9+k] as Taichi's default vector-field
 * layout, dense 9x9 transforms read from memory.  Arithmetic is strict IEEE (compile with
 * -ffp-contract=off, no fast-math), left to right as written in the reference, so the fp32
 * build is bit-identical to oracle/lbm_oracle_np.py, which in turn is pinned bit-for-bit to the
 * reference source executed under the Taichi stand-in (tests/golden/ti_shim_*.npz).
 *
 * Built twice from this one file: -DREAL=float -DSUF=_f32 and -DREAL=double -DSUF=_f64.
 * OpenMP parallelises the column loops exactly where Taichi parallelises its struct-fors.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifndef REAL
#define REAL float
#define SUF _f32
#endif
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)
#define R(x) ((REAL)(x))

typedef struct {
    int32_t nx, ny;
    int32_t warmup_steps;
    int32_t les_on;          /* C_smag > 0.001, ref:342 */
    int32_t bc_type[4];      /* W, top, E, bottom  (dr = 0..3), ref:445-450 */
    int32_t frame_count;     /* in/out, ref:440 */
    REAL tau0;               /* f32(3 nu + 0.5) */
    REAL tau0_sq;            /* f32(tau0_f64 ** 2): python-scope constant, ref:348 */
    REAL cs_sq_factor;       /* f32(18 Cs^2), ref:79 */
    REAL rho_in, rho_out;
    REAL bc_value[4][2];
    REAL S_base[9];          /* ref:191-201 */
    REAL w[9];
    REAL invM[81];           /* np.linalg.inv(M_f32).astype(f32), ref:182 */
} FN(OracleParams);

static const int EX[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
static const int EY[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
static const REAL MM[81] = {
    1, 1, 1, 1, 1, 1, 1, 1, 1,
    -4, -1, -1, -1, -1, 2, 2, 2, 2,
    4, -2, -2, -2, -2, 1, 1, 1, 1,
    0, 1, 0, -1, 0, 1, -1, -1, 1,
    0, -2, 0, 2, 0, 1, -1, -1, 1,
    0, 0, 1, 0, -1, 1, 1, -1, -1,
    0, 0, -2, 0, 2, 1, 1, -1, -1,
    0, 1, -1, 1, -1, 0, 0, 0, 0,
    0, 0, 0, 0, 0, 1, -1, 1, -1};

static inline REAL rsqrt_(REAL x) { return sizeof(REAL) == 4 ? (REAL)sqrtf((float)x) : (REAL)sqrt((double)x); }
static inline REAL rmax_(REAL a, REAL b) { return a > b ? a : b; }

/* ref:214-218 */
static inline void f_eq(const FN(OracleParams) * p, REAL rho, REAL ux, REAL uy, REAL *out) {
    REAL uv = ux * ux + uy * uy;
    for (int k = 0; k < 9; ++k) {
        REAL eu = R(EX[k]) * ux + R(EY[k]) * uy;
        out[k] = (p->w[k] * rho) * (R(1) + R(3) * eu + R(4.5) * eu * eu - R(1.5) * uv);
    }
}

/* ref:243-420 */
void FN(oracle_collide_and_stream)(const FN(OracleParams) * p, const REAL *f_old, REAL *f_new,
                                   const REAL *damp_x, const REAL *damp_y) {
    const int nx = p->nx, ny = p->ny;
#pragma omp parallel for schedule(static)
    for (int i = 1; i < nx - 1; ++i) {
        for (int j = 1; j < ny - 1; ++j) {
            REAL f[9], m[9], meq[9], ms[9];
            for (int k = 0; k < 9; ++k) f[k] = f_old[((size_t)(i - EX[k]) * ny + (j - EY[k])) * 9 + k];
            for (int r = 0; r < 9; ++r) {
                REAL val = R(0);
                for (int c = 0; c < 9; ++c) val += MM[r * 9 + c] * f[c];
                m[r] = val;
            }
            REAL rho = m[0], u = R(0), v = R(0);
            if (rho > R(0)) { u = m[3] / rho; v = m[5] / rho; }
            REAL u2 = u * u + v * v;
            meq[0] = rho;
            meq[1] = rho * (R(-2.0) + R(3.0) * u2);
            meq[2] = rho * (R(1.0) - R(3.0) * u2);
            meq[3] = rho * u;
            meq[4] = -rho * u;
            meq[5] = rho * v;
            meq[6] = -rho * v;
            meq[7] = rho * (u * u - v * v);
            meq[8] = rho * u * v;
            REAL n7 = m[7] - meq[7], n8 = m[8] - meq[8];
            REAL norm = rsqrt_(R(2.0) * n7 * n7 + R(2.0) * n8 * n8);
            REAL tau_eff = p->tau0;
            if (p->les_on) {
                REAL term = p->tau0_sq + (p->cs_sq_factor * norm) / rho;
                REAL tau_eddy = R(0.5) * (rsqrt_(term) - p->tau0);
                tau_eff = p->tau0 + tau_eddy;
            }
            tau_eff += rmax_(damp_x[i], damp_y[j]);
            REAL s_eff = R(1.0) / tau_eff;
            for (int k = 0; k < 9; ++k) {
                REAL S = (k >= 7) ? s_eff : p->S_base[k];
                ms[k] = m[k] - S * (m[k] - meq[k]);
            }
            REAL *out = f_new + ((size_t)i * ny + j) * 9;
            for (int r = 0; r < 9; ++r) {
                REAL val = R(0);
                for (int c = 0; c < 9; ++c) val += p->invM[r * 9 + c] * ms[c];
                out[r] = val;
            }
        }
    }
}

/* ref:422-436 */
void FN(oracle_update_macro_var)(const FN(OracleParams) * p, REAL *f_old, const REAL *f_new, REAL *rho, REAL *vel) {
    const int nx = p->nx, ny = p->ny;
#pragma omp parallel for schedule(static)
    for (int i = 1; i < nx - 1; ++i) {
        for (int j = 1; j < ny - 1; ++j) {
            size_t c = (size_t)i * ny + j;
            REAL lr = R(0), lx = R(0), ly = R(0);
            for (int k = 0; k < 9; ++k) {
                REAL fk = f_new[c * 9 + k];
                f_old[c * 9 + k] = fk;
                lr += fk;
                lx += R(EX[k]) * fk;
                ly += R(EY[k]) * fk;
            }
            rho[c] = lr;
            if (lr > R(0)) { vel[c * 2] = lx / lr; vel[c * 2 + 1] = ly / lr; }
            else { vel[c * 2] = R(0); vel[c * 2 + 1] = R(0); }
        }
    }
}

/* ref:457-550 */
static void apply_bc_core(const FN(OracleParams) * p, REAL *f_old, REAL *rho, REAL *vel, int dr,
                          int ibc, int jbc, int inb, int jnb, REAL ramp) {
    const int ny = p->ny;
    const size_t b = (size_t)ibc * ny + jbc, n = (size_t)inb * ny + jnb;
    REAL *fb = f_old + b * 9;
    const REAL *fn = f_old + n * 9;
    REAL eb[9], en[9];
    const int t = p->bc_type[dr];
    if (t == 0) {
        if (ibc == 0) {
            REAL rc = R(1.0) + (p->rho_in - R(1.0)) * ramp;
            REAL f0 = fn[0], f2 = fn[2], f3 = fn[3], f4 = fn[4], f6 = fn[6], f7 = fn[7];
            REAL ux = R(1.0) - (f0 + f2 + f4 + R(2.0) * (f3 + f6 + f7)) / rc;
            REAL f1 = f3 + R(2.0 / 3.0) * rc * ux;
            REAL f5 = f7 - R(0.5) * (f2 - f4) + R(1.0 / 6.0) * rc * ux;
            REAL f8 = f6 + R(0.5) * (f2 - f4) + R(1.0 / 6.0) * rc * ux;
            rho[b] = rc; vel[b * 2] = ux; vel[b * 2 + 1] = R(0);
            f_eq(p, rc, ux, R(0), eb);
            eb[1] = f1; eb[5] = f5; eb[8] = f8;
            memcpy(fb, eb, sizeof eb);
        } else {
            vel[b * 2] = p->bc_value[dr][0] * ramp;
            vel[b * 2 + 1] = p->bc_value[dr][1] * ramp;
            rho[b] = rho[n];
            f_eq(p, rho[b], vel[b * 2], vel[b * 2 + 1], eb);
            f_eq(p, rho[n], vel[n * 2], vel[n * 2 + 1], en);
            for (int k = 0; k < 9; ++k) fb[k] = eb[k] - en[k] + fn[k];
        }
    } else if (t == 1) {
        if (ibc == p->nx - 1) {
            REAL ro = p->rho_out;
            REAL f0 = fn[0], f1 = fn[1], f2 = fn[2], f4 = fn[4], f5 = fn[5], f8 = fn[8];
            REAL ux = R(-1.0) + (f0 + f2 + f4 + R(2.0) * (f1 + f5 + f8)) / ro;
            if (ux < R(0.0)) {
                vel[b * 2] = vel[n * 2]; vel[b * 2 + 1] = vel[n * 2 + 1];
                rho[b] = ro;
                f_eq(p, rho[b], vel[b * 2], vel[b * 2 + 1], eb);
                f_eq(p, rho[n], vel[n * 2], vel[n * 2 + 1], en);
                for (int k = 0; k < 9; ++k) fb[k] = eb[k] - en[k] + fn[k];
            } else {
                REAL f3 = f1 - R(2.0 / 3.0) * ro * ux;
                REAL f6 = f8 - R(0.5) * (f2 - f4) - R(1.0 / 6.0) * ro * ux;
                REAL f7 = f5 + R(0.5) * (f2 - f4) - R(1.0 / 6.0) * ro * ux;
                rho[b] = ro; vel[b * 2] = ux; vel[b * 2 + 1] = R(0);
                f_eq(p, ro, ux, R(0), eb);
                eb[3] = f3; eb[6] = f6; eb[7] = f7;
                memcpy(fb, eb, sizeof eb);
            }
        }
    } else if (t == 2) {
        if (ibc == inb) { vel[b * 2] = vel[n * 2]; vel[b * 2 + 1] = R(0); }
        else { vel[b * 2] = R(0); vel[b * 2 + 1] = vel[n * 2 + 1]; }
        rho[b] = rho[n];
        f_eq(p, rho[b], vel[b * 2], vel[b * 2 + 1], eb);
        f_eq(p, rho[n], vel[n * 2], vel[n * 2 + 1], en);
        for (int k = 0; k < 9; ++k) fb[k] = eb[k] - en[k] + fn[k];
    }
}

/* ref:442-443; the cosine is the correctly rounded f32 of the double cosine (see the numpy oracle) */
REAL FN(oracle_ramp)(int frame_count, int warmup_steps) {
    REAL progress = R(frame_count) / R(warmup_steps);
    if (progress > R(1.0)) progress = R(1.0);
    REAL arg = R(0.5 * 3.14159265) * progress;
    REAL c = (REAL)cos((double)arg);
    return R(1.0) - c;
}

/* ref:438-455 */
void FN(oracle_apply_bc)(FN(OracleParams) * p, REAL *f_old, REAL *rho, REAL *vel, const float *mask) {
    const int nx = p->nx, ny = p->ny;
    p->frame_count += 1;
    const REAL ramp = FN(oracle_ramp)(p->frame_count, p->warmup_steps);
#pragma omp parallel for schedule(static)
    for (int j = 1; j < ny - 1; ++j) {
        apply_bc_core(p, f_old, rho, vel, 0, 0, j, 1, j, ramp);
        apply_bc_core(p, f_old, rho, vel, 2, nx - 1, j, nx - 2, j, ramp);
    }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nx; ++i) {
        apply_bc_core(p, f_old, rho, vel, 1, i, ny - 1, i, ny - 2, ramp);
        apply_bc_core(p, f_old, rho, vel, 3, i, 0, i, 1, ramp);
    }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nx; ++i) {
        for (int j = 0; j < ny; ++j) {
            size_t c = (size_t)i * ny + j;
            if (mask[c] == 1.0f) {
                vel[c * 2] = R(0); vel[c * 2 + 1] = R(0);
                f_eq(p, rho[c], R(0), R(0), f_old + c * 9);
            }
        }
    }
}

/* ref:552-573 */
void FN(oracle_run_step)(FN(OracleParams) * p, REAL *f_old, REAL *f_new, REAL *rho, REAL *vel,
                         const float *mask, const REAL *damp_x, const REAL *damp_y, int steps) {
    for (int s = 0; s < steps; ++s) {
        FN(oracle_collide_and_stream)(p, f_old, f_new, damp_x, damp_y);
        FN(oracle_update_macro_var)(p, f_old, f_new, rho, vel);
        FN(oracle_apply_bc)(p, f_old, rho, vel, mask);
    }
}

/* ref:235-241 */
void FN(oracle_init)(FN(OracleParams) * p, REAL *f_old, REAL *f_new, REAL *rho, REAL *vel) {
    const size_t n = (size_t)p->nx * p->ny;
    p->frame_count = 0;
    for (size_t c = 0; c < n; ++c) {
        vel[c * 2] = R(0); vel[c * 2 + 1] = R(0);
        rho[c] = R(1);
        f_eq(p, rho[c], R(0), R(0), f_old + c * 9);
        memcpy(f_new + c * 9, f_old + c * 9, 9 * sizeof(REAL));
    }
}

/* ref:588-641; sequential (i, j, k) accumulation */
void FN(oracle_force)(const FN(OracleParams) * p, const REAL *f_new, const float *mask, REAL *out2) {
    static const int INV[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
    const int nx = p->nx, ny = p->ny;
    REAL fx = R(0), fy = R(0);
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < ny; ++j) {
            if (mask[(size_t)i * ny + j] != 1.0f) continue;
            for (int k = 0; k < 9; ++k) {
                int ni = i + EX[k], nj = j + EY[k];
                if (ni < 0 || ni >= nx || nj < 0 || nj >= ny) continue;
                if (mask[(size_t)ni * ny + nj] != 0.0f) continue;
                REAL fv = R(2.0) * f_new[((size_t)ni * ny + nj) * 9 + INV[k]];
                fx += fv * R(-EX[k]);
                fy += fv * R(-EY[k]);
            }
        }
    out2[0] = fx; out2[1] = fy;
}

/* ref:648-654; NaN propagates */
REAL FN(oracle_max_velocity)(const FN(OracleParams) * p, const REAL *vel) {
    const size_t n = (size_t)p->nx * p->ny;
    REAL mx = R(0);
    int nan = 0;
    for (size_t c = 0; c < n; ++c) {
        REAL v = rsqrt_(vel[c * 2] * vel[c * 2] + vel[c * 2 + 1] * vel[c * 2 + 1]);
        if (v != v) nan = 1;
        if (v > mx) mx = v;
    }
    return nan ? (REAL)NAN : mx;
}

/* ref:667-737 */
void FN(oracle_moments)(const FN(OracleParams) * p, const REAL *f_new, REAL *mom) {
    const size_t n = (size_t)p->nx * p->ny;
#pragma omp parallel for schedule(static)
    for (size_t c = 0; c < n; ++c) {
        const REAL *f = f_new + c * 9;
        REAL *o = mom + c * 9;
        REAL rho = R(0);
        for (int k = 0; k < 9; ++k) rho += f[k];
        REAL s14 = f[1] + f[2] + f[3] + f[4];
        REAL s58 = f[5] + f[6] + f[7] + f[8];
        o[0] = rho;
        o[1] = R(-4.0) * f[0] - s14 + R(2.0) * s58;
        o[2] = R(4.0) * f[0] - R(2.0) * s14 + s58;
        o[3] = f[1] - f[3] + f[5] - f[6] - f[7] + f[8];
        o[4] = R(-2.0) * f[1] + R(2.0) * f[3] + f[5] - f[6] - f[7] + f[8];
        o[5] = f[2] - f[4] + f[5] + f[6] - f[7] - f[8];
        o[6] = R(-2.0) * f[2] + R(2.0) * f[4] + f[5] + f[6] - f[7] - f[8];
        o[7] = f[1] - f[2] + f[3] - f[4];
        o[8] = f[5] - f[6] + f[7] - f[8];
    }
}
