// Device-side building blocks of the engine: the fexp restatement, the
// per-gaussian evaluation constants, model fill / convolve / norms, and the
// block reductions.  Built with -fmad=false: every fused multiply-add in the
// hot loops is an explicit fma(), everything else rounds exactly like the
// reference's (numba, no contraction) scalar code.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/ngx.h"

#define NGX_LOW_DETVAL 1.0e-200       // GMIX_LOW_DETVAL, gmix_nb.py:11
#define NGX_MAX_GAUSS 160             // 16 model x 10 psf gaussians
#define NGX_TWO_PI 6.283185307179586476925286766559

namespace ngx {

// exp(i), i=-15..0 (fastexp_nb.py:5-16); correctly rounded doubles
__constant__ double c_exp_lookup[16] = {
    3.059023205018258e-07, 8.315287191035679e-07, 2.2603294069810542e-06,
    6.14421235332821e-06, 1.670170079024566e-05, 4.5399929762484854e-05,
    0.00012340980408667956, 0.00033546262790251185, 0.0009118819655545162,
    0.0024787521766663585, 0.006737946999085467, 0.01831563888873418,
    0.049787068367863944, 0.1353352832366127, 0.36787944117144233, 1.0};

// model tables, gmix_nb.py:243-304
__constant__ double c_pvals_exp[6] = {
    0.00061601229677880041, 0.0079461395724623237, 0.053280454055540001,
    0.21797364640726541, 0.45496740582554868, 0.26521634184240478};
__constant__ double c_fvals_exp[6] = {
    0.002467115141477932, 0.018147435573256168, 0.07944063151366336,
    0.27137669897479122, 0.79782256866993773, 2.1623306025075739};
__constant__ double c_pvals_dev[10] = {
    6.5288960012625658e-05, 0.00044199216814302695, 0.0020859587871659754,
    0.0075913681418996841, 0.02260266219257237, 0.056532254390212859,
    0.11939049233042602, 0.20969545753234975, 0.29254151133139222,
    0.28905301416582552};
__constant__ double c_fvals_dev[10] = {
    2.9934935706271918e-07, 3.4651596338231207e-06, 2.4807910570562753e-05,
    1.4307404300535354e-04, 7.2753169298239500e-04, 3.4582464394427260e-03,
    1.6086645440719100e-02, 7.7006776775654429e-02, 4.1012562102501476e-01,
    2.9812509778548648e00};
__constant__ double c_pvals_turb[3] = {
    0.596510042804182, 0.4034898268889178, 1.303069003078001e-07};
__constant__ double c_fvals_turb[3] = {
    0.5793612389470884, 1.621860687127999, 7.019347162356363};

// exp5_smooth coefficients, fastexp_nb.py:252-260
#define NGX_C0 1.0000011318561302
#define NGX_C1 0.999993601071577
#define NGX_C2 0.49992478810274166
#define NGX_C3 0.16674612720799442
#define NGX_C4 0.042330947141114836
#define NGX_C5 0.008197933236258961

// 1.5 * 2^52: adding it with round-up leaves ceil(y) in the low mantissa bits
#define NGX_MAGIC 6755399441055744.0

__host__ __device__ inline int model_ngauss(int model) {
    switch (model) {
    case NGX_GAUSS: return 1;
    case NGX_TURB: return 3;
    case NGX_EXP: return 6;
    case NGX_DEV: return 10;
    case NGX_BDF: case NGX_BD: case NGX_CM: return 16;
    default: return -1;
    }
}
__host__ __device__ inline int model_npars(int model) {
    switch (model) {
    case NGX_GAUSS: case NGX_TURB: case NGX_EXP: case NGX_DEV: case NGX_CM:
        return 6;
    case NGX_BDF: return 7;
    case NGX_BD: return 8;
    default: return -1;
    }
}

// ---------------------------------------------------------------------
// fexp = exp5_smooth (fastexp_nb.py:223-262), split into the range
// reduction and the polynomial.  ival = int(x-0.5) truncates toward zero;
// for x <= 0 that is ceil(x-0.5), obtained without a float->int conversion
// by a round-up add of NGX_MAGIC.  f = x - ival is exact.
// Returns exp(ival) index (ival+15) through *idx.
__device__ __forceinline__ double fexp_reduce(double x, int *idx) {
    double y = __dadd_rn(x, -0.5);
    double t = __dadd_ru(y, NGX_MAGIC);
    *idx = __double2loint(t) + 15;
    double r = __dadd_rn(t, -NGX_MAGIC);
    return __dadd_rn(x, -r);
}

// the quintic, Horner with fused multiply-adds (<= 1 ulp from the
// reference's unfused Horner)
__device__ __forceinline__ double fexp_poly(double f) {
    double p = fma(f, NGX_C5, NGX_C4);
    p = fma(f, p, NGX_C3);
    p = fma(f, p, NGX_C2);
    p = fma(f, p, NGX_C1);
    p = fma(f, p, NGX_C0);
    return p;
}

// bit-faithful fexp (unfused, reference operation order); used where the
// cost does not matter
__device__ __forceinline__ double fexp_ref(double x) {
    int idx;
    double f = fexp_reduce(x, &idx);
    double p = NGX_C0 + f * (NGX_C1 + f * (NGX_C2 + f * (NGX_C3 + f * (
        NGX_C4 + f * NGX_C5))));
    return c_exp_lookup[idx] * p;
}

// the same with the exp(i) table in shared memory (per-lane index)
__device__ __forceinline__ double fexp_ref_tab(double x, const double *tab) {
    int idx;
    double f = fexp_reduce(x, &idx);
    double p = NGX_C0 + f * (NGX_C1 + f * (NGX_C2 + f * (NGX_C3 + f * (
        NGX_C4 + f * NGX_C5))));
    return tab[idx] * p;
}

// apod_window (fastexp_nb.py:97-117) in terms of x = -chi2/2;
// 25 - chi2 == fma(2, x, 25) exactly as the reference rounds it
__device__ __forceinline__ double apod_window_x(double x) {
    double u = fma(2.0, x, 25.0) * 0.2;
    return u * u * u * (10.0 + u * (-15.0 + 6.0 * u));
}
// the same window with fused multiply-adds (hot loops; within an ulp or two
// of apod_window_x)
__constant__ double c_apod[5] = {0.4, 5.0, 6.0, -15.0, 10.0};
__device__ __forceinline__ double apod_window_fast(double x) {
    double u = fma(c_apod[0], x, c_apod[1]);
    double t = fma(fma(c_apod[2], u, c_apod[3]), u, c_apod[4]);
    return (u * u) * (u * t);
}
// apod_window_deriv (fastexp_nb.py:120-135)
__device__ __forceinline__ double apod_window_deriv_x(double x) {
    double u = fma(2.0, x, 25.0) * 0.2;
    double umu = u * (1.0 - u);
    return -30.0 * umu * umu * 0.2;
}

// range tests on x = -chi2/2 done on the integer pipe:
//   in_range : chi2 >= 0 && chi2 < 25   <=>  x <= 0 && x > -12.5
//   in_apod  : chi2 > 20 (window == 1 exactly at chi2 == 20)
__device__ __forceinline__ bool x_in_range(double x) {
    unsigned hi = (unsigned)__double2hiint(x);
    unsigned lo = (unsigned)__double2loint(x);
    return ((hi - 0x80000000u) < 0x40290000u) || ((hi | lo) == 0u);
}
__device__ __forceinline__ bool x_in_apod(double x) {
    unsigned hi = (unsigned)__double2hiint(x) & 0x7fffffffu;
    return hi >= 0x40240000u;
}

// ---------------------------------------------------------------------
// The hot-loop form of fexp.  Everything that is not fp64 arithmetic is
// squeezed out of the per-evaluation path:
//  * the quintic's coefficients are constant-bank operands of the DFMAs
//    (a literal costs two 32-bit moves per use once registers run short);
//  * the range tests work on the high word of y = x - 1/2, which the range
//    reduction needs anyway: with hy = hi(y) - hi(-1/2) (unsigned)
//        0 <= chi2 < 25  <=>  -13 < y <= -1/2  <=>  hy < NGX_HY_RANGE
//        chi2 >= 20      <=>  y <= -10.5       <=>  hy >= NGX_HY_APOD
//    (x - 1/2 is exact on these boundaries; chi2 == 20 has window 1 and slope
//    0 exactly, so including it in the band changes nothing), and x <= 0 is
//    one signed compare of hi(x);
//  * the per-gaussian table pn*exp(i) is stored ROTATED, slot = ival & 15
//    (ival in [-12, 0] uses slots 0 and 4..15), and slot 1 holds 0.0: a lane
//    outside the footprint reads slot 1, so its value is an exact zero without
//    selects on the 64-bit result.
__constant__ double c_fexp_c[6] = {NGX_C0, NGX_C1, NGX_C2,
                                   NGX_C3, NGX_C4, NGX_C5};
#define NGX_HY_BASE 0xBFE00000u
#define NGX_HY_RANGE 0x004A0000u      // hi(-13) - hi(-1/2)
#define NGX_HY_APOD 0x00450000u       // hi(-10.5) - hi(-1/2)
#define NGX_TAB_ZERO 1u

__device__ __forceinline__ double fexp_poly_c(double f) {
    double p = fma(f, c_fexp_c[5], c_fexp_c[4]);
    p = fma(f, p, c_fexp_c[3]);
    p = fma(f, p, c_fexp_c[2]);
    p = fma(f, p, c_fexp_c[1]);
    p = fma(f, p, c_fexp_c[0]);
    return p;
}

// range reduction + range tests; returns f, the table slot, and the
// in-range / in-apodization-band predicates
__device__ __forceinline__ double fexp_split(double x, unsigned *slot,
                                             bool *inr, bool *ap) {
    double y = __dadd_rn(x, -0.5);
    unsigned hy = (unsigned)__double2hiint(y) - NGX_HY_BASE;
    bool in = (hy < NGX_HY_RANGE) && (__double2hiint(x) <= 0);
    *inr = in;
    *ap = in && (hy >= NGX_HY_APOD);
    double t = __dadd_ru(y, NGX_MAGIC);
    *slot = in ? ((unsigned)__double2loint(t) & 15u) : NGX_TAB_ZERO;
    double r = __dadd_rn(t, -NGX_MAGIC);
    return __dadd_rn(x, -r);
}

// fill one gaussian's rotated table (16 doubles)
__device__ __forceinline__ void fill_rot_table(double *tab, double pn) {
#pragma unroll
    for (int t = 0; t < 16; t++)
        tab[(t + 1) & 15] = (t >= 3) ? c_exp_lookup[t] * pn : 0.0;
}

// apodization window and the combination w - 2 w' the derivative images
// need (derivs_nb.py:104-108), sharing u = (25 - chi2) / 5:
//   w = u^3 (10 - 15 u + 6 u^2),  w' = -6 (u (1 - u))^2
__device__ __forceinline__ void apod_pair(double x, double *w, double *wc) {
    double u = fma(0.4, x, 5.0);
    double t = fma(fma(6.0, u, -15.0), u, 10.0);
    double ww = (u * u) * (u * t);
    double umu = fma(-u, u, u);
    *w = ww;
    *wc = fma(12.0 * umu, umu, ww);
}

// tile -> (tile row, tile column) without the ~20-instruction integer
// division: (tile + 1/2) / tpr in fp32 truncates to the exact quotient for
// any tile count below 2^20
__device__ __forceinline__ void tile_rowcol(int tile, int tpr, int *trow,
                                            int *tcol) {
    int r = __float2int_rz(__fdividef((float)tile + 0.5f, (float)tpr));
    *trow = r;
    *tcol = tile - r * tpr;
}

// 4 x 8 pixel tile -> centre and half extents in (v, u)
__device__ __forceinline__ void tile_extent(const ngx_jacobian &j, int tile,
                                            int tpr, int nrow, int ncol,
                                            double *vc, double *uc,
                                            double *hv, double *hu) {
    int trow, tcol;
    tile_rowcol(tile, tpr, &trow, &tcol);
    int r0 = trow << 2, c0 = tcol << 3;
    int r1 = min(r0 + 3, nrow - 1), c1 = min(c0 + 7, ncol - 1);
    double rc = 0.5 * (r0 + r1) - j.row0, cc = 0.5 * (c0 + c1) - j.col0;
    double hr = 0.5 * (r1 - r0), hc = 0.5 * (c1 - c0);
    *vc = j.dvdrow * rc + j.dvdcol * cc;
    *uc = j.dudrow * rc + j.dudcol * cc;
    *hv = fabs(j.dvdrow) * hr + fabs(j.dvdcol) * hc;
    *hu = fabs(j.dudrow) * hr + fabs(j.dudcol) * hc;
}

// does the chi2 < 25 footprint of a gaussian at (row, col) with
// rv = 5 sqrt(irr), ru = 5 sqrt(icc) possibly meet the tile?  chi2 < 25
// implies |dv| < rv and |du| < ru; the margins cover rounding, and anything
// not finite counts as a hit.
__device__ __forceinline__ bool tile_hit(double vc, double uc, double hv,
                                         double hu, double row, double col,
                                         double rv, double ru) {
    double dv = fabs(vc - row), du = fabs(uc - col);
    double tv = (rv + hv) * (1.0 + 1.0e-9) + 1.0e-9 * (fabs(vc) + fabs(row));
    double tu = (ru + hu) * (1.0 + 1.0e-9) + 1.0e-9 * (fabs(uc) + fabs(col));
    return !(dv > tv) && !(du > tu);
}

// ---------------------------------------------------------------------
// per-gaussian constants for the value pass:
//   x = -chi2/2 = a dv^2 + b du^2 + c dv du,  value = pn * fexp(x)
struct GaussV {
    double row, col, a, b, c, pn;
};

// one apodized fast evaluation, gauss2d_eval_pixel_fast (gmix_nb.py:28-63).
// tab is the 16-entry exp(i) table in SHARED memory (a __constant__ table
// would serialise on the per-lane index).
__device__ __forceinline__ double eval_fast(const GaussV &g, double v,
                                            double u, const double *tab) {
    double dv = v - g.row;
    double du = u - g.col;
    double t = fma(g.c, du, g.a * dv);
    double x = fma(g.b * du, du, dv * t);
    double val = 0.0;
    if (x_in_range(x)) {
        int idx;
        double f = fexp_reduce(x, &idx);
        val = (tab[idx] * g.pn) * fexp_poly(f);
        if (x_in_apod(x)) val *= apod_window_x(x);
    }
    return val;
}

// GaussV from a normalised ngx_gauss record and the pixel area
__device__ __forceinline__ GaussV make_gaussv(const ngx_gauss &g,
                                              double area) {
    GaussV c;
    c.row = g.row; c.col = g.col;
    c.a = -0.5 * g.dcc; c.b = -0.5 * g.drr; c.c = g.drc;
    c.pn = g.pnorm * area;
    return c;
}

// copy the exp(i) table into shared memory (call from all threads, then sync)
__device__ __forceinline__ void load_exp_table(double *tab) {
    if (threadIdx.x < 16) tab[threadIdx.x] = c_exp_lookup[threadIdx.x];
}

// exact evaluation, gauss2d_eval_pixel (gmix_nb.py:66-92): libm exp, no cut
__device__ __forceinline__ double eval_exact(const ngx_gauss &g, double v,
                                             double u, double area) {
    double vdiff = v - g.row;
    double udiff = u - g.col;
    double chi2 = g.dcc * vdiff * vdiff + g.drr * udiff * udiff
        - 2.0 * g.drc * vdiff * udiff;
    return g.pnorm * exp(-0.5 * chi2) * area;
}

// tiled index q -> (row, col); tiles are 4 rows x 8 cols, row-band major
__device__ __forceinline__ void q_to_rowcol(int q, int tpr, int *row,
                                            int *col) {
    int tile = q >> 5, w = q & 31;
    int trow, tcol;
    tile_rowcol(tile, tpr, &trow, &tcol);
    *row = (trow << 2) + (w >> 3);
    *col = (tcol << 3) + (w & 7);
}

// exp part of one evaluation for a lane that may be out of range: the table
// index is clamped so the shared-memory read stays in bounds and the result
// is only used under the in-range predicate.
__device__ __forceinline__ double fexp_tab(double x, const double *tab) {
    int idx;
    double f = fexp_reduce(x, &idx);
    idx = min((unsigned)idx, 15u);
    return tab[idx] * fexp_poly(f);
}

// ---------------------------------------------------------------------
// gauss2d_set (gmix_nb.py:221-240)
__device__ __forceinline__ void gauss_set(ngx_gauss &g, double p, double row,
                                          double col, double irr, double irc,
                                          double icc) {
    g.norm_set = 0;
    g.drr = nan(""); g.drc = nan(""); g.dcc = nan("");
    g.norm = nan(""); g.pnorm = nan("");
    g.p = p; g.row = row; g.col = col;
    g.irr = irr; g.irc = irc; g.icc = icc;
    g.det = irr * icc - irc * irc;
}

// gauss2d_set_norm (gmix_nb.py:190-218); returns NGX_ST_*
__device__ __forceinline__ int gauss_set_norm(ngx_gauss &g) {
    if (g.det < NGX_LOW_DETVAL) return NGX_ST_RANGE_ERROR;
    double T = g.irr + g.icc;
    if (T <= NGX_LOW_DETVAL) return NGX_ST_RANGE_ERROR;
    double idet = 1.0 / g.det;
    g.drr = g.irr * idet;
    g.drc = g.irc * idet;
    g.dcc = g.icc * idet;
    g.norm = 1.0 / (2 * 3.141592653589793 * sqrt(g.det));
    g.pnorm = g.p * g.norm;
    g.norm_set = 1;
    return NGX_ST_OK;
}

// g1g2_to_e1e2 (gmix_nb.py:652-678)
__device__ __forceinline__ int g1g2_to_e1e2(double g1, double g2, double *e1,
                                            double *e2) {
    double g = sqrt(g1 * g1 + g2 * g2);
    if (!(g < 1.0)) return NGX_ST_RANGE_ERROR;
    if (g == 0.0) {
        *e1 = 0.0; *e2 = 0.0;
    } else {
        double eta = 2 * atanh(g);
        double e = tanh(eta);
        if (e >= 1.0) e = 0.99999999;
        double fac = e / g;
        *e1 = fac * g1;
        *e2 = fac * g2;
    }
    return NGX_ST_OK;
}

// get_cm_Tfactor (gmix_nb.py:561-593)
__device__ __forceinline__ double cm_Tfactor(double fracdev, double TdByTe) {
    double ifracdev = 1.0 - fracdev;
    double Tfactor = 0.0;
    for (int i = 0; i < 6; i++)
        Tfactor += (c_pvals_exp[i] * ifracdev) * c_fvals_exp[i];
    for (int i = 0; i < 10; i++)
        Tfactor += (c_pvals_dev[i] * fracdev) * (c_fvals_dev[i] * TdByTe);
    return 1.0 / Tfactor;
}

// Shape/size/flux of the model, resolved once per parameter vector
struct ModelHead {
    double row, col, e1, e2, T, flux, fracdev, TdByTe;
    int status;
};

// the part of gmix_fill_<model> that does not depend on the component
__device__ inline ModelHead model_head(int model, const double *pars) {
    ModelHead h;
    h.row = pars[0]; h.col = pars[1];
    h.fracdev = 0.0; h.TdByTe = 1.0;
    h.T = pars[4];
    if (model == NGX_BDF) {
        h.fracdev = pars[5];
        h.flux = pars[6];
        h.T = pars[4] * cm_Tfactor(h.fracdev, 1.0);
    } else if (model == NGX_BD) {
        h.TdByTe = pow(10.0, pars[5]);
        h.fracdev = pars[6];
        h.flux = pars[7];
        h.T = pars[4] * cm_Tfactor(h.fracdev, h.TdByTe);
    } else {
        h.flux = pars[5];
    }
    h.status = g1g2_to_e1e2(pars[2], pars[3], &h.e1, &h.e2);
    return h;
}

// component i of the unconvolved model: p, irr, irc, icc
// (gmix_fill_simple :307-340, composite :430-558)
__device__ __forceinline__ void model_component(int model, const ModelHead &h,
                                                int i, double *p, double *irr,
                                                double *irc, double *icc) {
    double f, pv;
    switch (model) {
    case NGX_GAUSS: f = 1.0; pv = 1.0; break;
    case NGX_TURB: f = c_fvals_turb[i]; pv = c_pvals_turb[i]; break;
    case NGX_EXP: f = c_fvals_exp[i]; pv = c_pvals_exp[i]; break;
    case NGX_DEV: f = c_fvals_dev[i]; pv = c_pvals_dev[i]; break;
    default:   // bdf, bd, cm
        if (i < 6) {
            pv = c_pvals_exp[i] * (1.0 - h.fracdev);
            f = c_fvals_exp[i];
        } else {
            pv = c_pvals_dev[i - 6] * h.fracdev;
            f = c_fvals_dev[i - 6] * h.TdByTe;
        }
    }
    double T_i_2 = 0.5 * h.T * f;
    *p = h.flux * pv;
    *irr = T_i_2 * (1 - h.e1);
    *irc = T_i_2 * h.e2;
    *icc = T_i_2 * (1 + h.e1);
}

// ---------------------------------------------------------------------
// reductions.  Deterministic: fixed shuffle tree, then a fixed-order sum of
// the per-warp partials by every thread (no atomics).
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of NV values per thread; result valid in ALL threads.
// scratch must hold NV * nwarps doubles.  Contains two __syncthreads.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *scratch) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; i++) {
        double s = warp_sum(v[i]);
        if (lane == 0) scratch[i * nwarps + warp] = s;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; i++) {
        double s = 0.0;
        for (int w = 0; w < nwarps; w++) s += scratch[i * nwarps + w];
        v[i] = s;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------
// Warp reduction of up to 32 values per lane with a halving butterfly: at
// each of the five steps a lane hands half of its values to its partner and
// adds the partner's half of the others, so all NV sums cost NV-1 shuffles
// in total instead of 5 NV.  On return lane L holds the warp total of v[L]
// (L < NV) in v[0].  Fixed order, deterministic.
template <int NV>
__device__ __forceinline__ void warp_sum_multi(double (&v)[NV]) {
    static_assert(NV <= 32, "at most one value per lane");
    const int lane = threadIdx.x & 31;
    double w[32];
#pragma unroll
    for (int i = 0; i < 32; i++) w[i] = i < NV ? v[i] : 0.0;
#pragma unroll
    for (int h = 16; h >= 1; h >>= 1) {
        const bool up = (lane & h) != 0;
#pragma unroll
        for (int i = 0; i < h; i++) {
            // values that are structurally zero on both sides stay zero
            if (i >= NV && i + h >= NV) continue;
            double send = up ? w[i] : w[i + h];
            double keep = up ? w[i + h] : w[i];
            w[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
        }
    }
    v[0] = w[0];
}

// Block totals of NV <= 32 per-thread values, left in shared memory:
// out[i] = sum over the block of v[i].  scratch holds 32 doubles per warp.
// Contains two __syncthreads.
template <int NV>
__device__ __forceinline__ void block_sum_to(double (&v)[NV], double *scratch,
                                             double *out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = (blockDim.x + 31) >> 5;
    warp_sum_multi<NV>(v);
    if (lane < NV) scratch[warp * 32 + lane] = v[0];
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
        for (int w = 0; w < nwarps; w++) s += scratch[w * 32 + threadIdx.x];
        out[threadIdx.x] = s;
    }
    __syncthreads();
}

}  // namespace ngx
