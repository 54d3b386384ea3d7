// ngx_lm_fit: the batched Levenberg-Marquardt fit, one CTA per stamp.
//
// Replaces Fitter.go (ngmix/fitting/fitters.py:64-112) = FitModel.calc_fdiff
// / calc_jacobian (fitting/results.py:439-625) + deriv_images
// (fitting/derivs_nb.py:40-127) + scipy's MINPACK lmder/lmdif
// (fitting/leastsqbound.py:445-447) + run_leastsq's epilogue (:70-126) +
// FitModel.set_fit_result (results.py:45-72).
//
// Data movement: the stamp (pixel values and weights, 16 B/pixel) is read
// from HBM once.  The pixel values live in shared memory (tiled order) for
// the whole fit; the weights stay in L2 and each pass re-reads them there
// (ierr = sqrt(w) is recomputed where it is used), which keeps a stamp's
// shared-memory footprint at 8 B/pixel.  The next stamp of the CTA is
// prefetched into L2 while the current one is fitted.  Per fit only the
// 1.2 KB result record goes out; the reference's 48 B/pixel AoS and the
// m x n Jacobian never exist.
//
// Pixel order: 4 rows x 8 columns per 32-pixel tile, C consecutive tiles per
// warp and chunk.  Before every pass each tile gets a bit mask of the
// gaussians whose chi^2 < 25 bounding box can reach it, so a warp only
// visits gaussians that may contribute; the narrow components of the
// exp/dev mixtures touch only a few tiles.
//
// Work per fit: nfev value passes (18 fp64 instr per gaussian-pixel) and
// njev derivative passes (30 per gaussian-pixel) over pixels kept in
// registers C at a time with the gaussians in the outer loop; every
// derivative pass also accumulates J^T J (21) and J^T f (6) per thread; the
// 31 sums are reduced with a halving butterfly (31 shuffles per warp).  One
// thread then advances the MINPACK state machine (ngx_lm_solve.cuh) while
// the other CTAs resident on the SM keep the fp64 pipe busy.
#include <cuda_pipeline.h>
#include <chrono>
#include <functional>
#include <mutex>
#include <vector>
#include "ngx_core.cuh"
#include "ngx_host.cuh"
#include "ngx_lm_solve.cuh"
#include "ngx_prior.cuh"

namespace ngx {

#define NGX_PDEF (-9.999e9)
#define NGX_CDEF (9.999e9)
#define NGX_BIGVAL (9999.0e47)

// per-gaussian constants in shared memory + 16-entry table pn*exp(i)
//   0 row, 1 col, 2 a=-dcc/2, 3 b=-drr/2, 4 c2=drc/2, 5 c=drc, 6 h=Tk/2,
//   7 qc=c2/2, 8 ha=a/2, 9 hb=b/2, 10 hb-ha, 11 hb+ha,
//   12 rv=5 sqrt(irr), 13 ru=5 sqrt(icc)
// (pairs are read as 16-byte shared-memory loads).  One record per gaussian:
// the 16-entry rotated pn*exp(i) table followed by the constants.
#define GC_STRIDE 14
#define REC (16 + GC_STRIDE)
// The forward-difference pass keeps np + 1 mixtures resident; they use a
// compact record (row, col, a, b, c, pn) and one shared rotated exp(i) table.
#define FDREC 6

struct LMArgs {
    int n, model, np, ngauss0, npsf, ngauss, nrow, ncol, ntot;
    int ignore_zero_weight, analytic, nq, tpr;
    // per round a warp moves on by rot_v / rot_d chunk slots (value pass /
    // derivative and forward-difference passes): see pick_rotation
    int rot_v, rot_d;
    ngx_prior prior;
    ngx_lm_pars lm;
    const double *images, *weights;
    const ngx_jacobian *jac;
    const double *psf, *guess;
    void *res;              // ngx_lm_result[] (or _wide[] for the multi kernel)
    int *counter;
    // objects with several stamps (ngx_lm_fit_multi)
    int nband;
    const int32_t *stamp_start, *stamp_band;
    // stamps too large for shared memory (GS kernels): per-CTA planes
    double *gscratch;
    // retry rounds of the fitting chain: fit only objects with todo != 0
    const int32_t *todo;
    // streamed host input (lm_fit_host_stream): the stamps arrive by DMA while
    // the kernel runs.  ready[c] becomes 1 by a copy ordered after the stamps
    // of chunk c (wait_ready has the chunk boundaries) and their jacobians,
    // psfs and guesses.  src 0: images / weights are double planes; 1: float planes
    // (im32, wt32); 2: float pixels, one weight per stamp (wscalar) and
    // optional mask bits (numpy.packbits order, nmb bytes per stamp).
    // done_count[c] counts the finished records of [c * done_chunk, ...); the
    // CTA that completes a chunk raises host_done[c] (mapped pinned memory) so
    // the host can start that chunk's device-to-host copy at once.
    const int *ready;
    int ready_chunk, src, nmb, done_chunk;
    const float *im32, *wt32;
    const double *wscalar;
    const unsigned char *maskbits;
    int *done_count, *host_done, *err;
    // the error flag the waiters poll: in device memory (polling the mapped
    // host flag from every waiting thread floods PCIe and starves the very
    // copies being waited for); `err` above is written once, on a timeout
    int *err_dev;
};

__device__ __forceinline__ int ld_acquire_sys(const int *p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];"
                 : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// wait until the chunk of `stamp` has arrived (a few seconds at most: a feed
// that stalls marks the call as failed instead of hanging the device)
// chunks: [0,S) [S,2S) [2S,4S) [4S,8S) and then 8S = ready_chunk stamps each,
// S = ready_chunk / 8 (short first chunks: the first CTAs start after ~50 us
// of copying; long later ones: few copies to set up)
__device__ __forceinline__ const int *ready_flag(const LMArgs &a, int stamp) {
    const int C = a.ready_chunk, S = C >> 3;
    const int idx = stamp >= C ? 3 + stamp / C
        : (stamp < S ? 0 : stamp < 2 * S ? 1 : stamp < 4 * S ? 2 : 3);
    return a.ready + idx;
}
__device__ __forceinline__ bool chunk_arrived(const LMArgs &a, int stamp) {
    return *(const volatile int *)ready_flag(a, stamp) != 0;
}
static __device__ __noinline__ void wait_ready(const LMArgs &a, int stamp) {
    const int *f = ready_flag(a, stamp);
    if (ld_acquire_sys(f)) return;
    const unsigned long long t0 = global_ns();
    for (;;) {
        __nanosleep(400);
        if (ld_acquire_sys(f)) return;
        // a failed call (the host raises err when it cannot queue the rest
        // of the feed) releases every waiter at once
        if (*(volatile int *)a.err_dev) return;
        if (global_ns() - t0 > 4000000000ull) {
            *(volatile int *)a.err_dev = 1;
            *(volatile int *)a.err = 1;
            return;
        }
    }
}
// one thread, after the record of `stamp` is complete
__device__ __forceinline__ void mark_done(const LMArgs &a, int stamp) {
    if (!a.done_count) return;
    __threadfence();
    const int c = stamp / a.done_chunk;
    int size = a.n - c * a.done_chunk;
    if (size > a.done_chunk) size = a.done_chunk;
    if (atomicAdd(a.done_count + c, 1) + 1 == size) {
        __threadfence_system();
        *(volatile int *)(a.host_done + c) = 1;
    }
}

// bounded fits: leastsqbound's parameter transforms (leastsqbound.py:183-263).
// MINPACK iterates on internal (unconstrained) parameters, the model and the
// priors are evaluated at the external ones.
__device__ __forceinline__ double b_i2e(const ngx_lm_pars &lm, int i,
                                        double x) {
    if (lm.nbounds == 0) return x;
    switch (lm.btype[i]) {
    case NGX_B_LOWER: return lm.blo[i] - 1.0 + sqrt(x * x + 1.0);
    case NGX_B_UPPER: return lm.bhi[i] + 1.0 - sqrt(x * x + 1.0);
    case NGX_B_BOTH:
        return lm.blo[i] + ((lm.bhi[i] - lm.blo[i]) / 2.0) * (sin(x) + 1.0);
    default: return x;
    }
}
__device__ __forceinline__ double b_e2i(const ngx_lm_pars &lm, int i,
                                        double x) {
    if (lm.nbounds == 0) return x;
    double t;
    switch (lm.btype[i]) {
    case NGX_B_LOWER: t = x - lm.blo[i] + 1.0; return sqrt(t * t - 1.0);
    case NGX_B_UPPER: t = lm.bhi[i] - x + 1.0; return sqrt(t * t - 1.0);
    case NGX_B_BOTH:
        return asin((2.0 * (x - lm.blo[i]) / (lm.bhi[i] - lm.blo[i])) - 1.0);
    default: return x;
    }
}
// d(external)/d(internal), _internal2external_grad
__device__ __forceinline__ double b_grad(const ngx_lm_pars &lm, int i,
                                         double v) {
    if (lm.nbounds == 0) return 1.0;
    switch (lm.btype[i]) {
    case NGX_B_LOWER: return v / sqrt(v * v + 1.0);
    case NGX_B_UPPER: return -v / sqrt(v * v + 1.0);
    case NGX_B_BOTH: return (lm.bhi[i] - lm.blo[i]) * cos(v) / 2.0;
    default: return 1.0;
    }
}

struct LMShared {
    LMState st;
    double xeval[LM_MAXP];     // external parameters of the trial point
    double xcur[LM_MAXP];      // external parameters of the accepted point
    double psf[16 * 6];
    double psf_row, psf_col, psf_ipsum;
    double cur_ssq, cur_rfirst2, cur_num, cur_den;   // at the accepted x
    ngx_jacobian jac;
    double tot[128];           // block totals of the pass sums
    int stamp, next, npix, first_q, status, action, spec;
    int warp_i[32], warp_j[32], warp_k[32];
};

// ---------------------------------------------------------------------
// build the convolved mixture for parameters x: thread k < ngauss owns
// composed gaussian k = iobj*npsf + ipsf (gmix_convolve_fill's order).
// sh->status is OR-ed with range errors.
static __device__ __noinline__ void build_component(const LMArgs &a,
                                             LMShared *sh, const double *x,
                                             int k, double *mix, double area,
                                             double *compact) {
    if (k >= a.ngauss) return;
    ModelHead h = model_head(a.model, x);
    if (h.status != NGX_ST_OK) {
        atomicOr(&sh->status, 1);
        return;
    }
    int npsf = a.npsf > 0 ? a.npsf : 1;
    int iobj = k / npsf, ipsf = k - iobj * npsf;
    double p, irr0, irc0, icc0;
    if (a.model == NGX_COELLIP) {
        // gmix_fill_coellip (gmix_nb.py:375-405): pars = (cen1, cen2, g1, g2,
        // T_1..T_n, F_1..F_n), every gaussian with the shared shape
        double Thalf = 0.5 * x[4 + iobj];
        p = x[4 + a.ngauss0 + iobj];
        irr0 = Thalf * (1 - h.e1);
        irc0 = Thalf * h.e2;
        icc0 = Thalf * (1 + h.e1);
    } else {
        model_component(a.model, h, iobj, &p, &irr0, &irc0, &icc0);
    }
    double row = h.row, col = h.col, irr = irr0, irc = irc0, icc = icc0;
    if (a.npsf > 0) {
        const double *q = sh->psf + 6 * ipsf;
        p = p * q[0] * sh->psf_ipsum;
        row = h.row + (q[1] - sh->psf_row);
        col = h.col + (q[2] - sh->psf_col);
        irr = irr0 + q[3];
        irc = irc0 + q[4];
        icc = icc0 + q[5];
    }
    double det = irr * icc - irc * irc;
    // gauss2d_set_norm :203-208
    if (det < NGX_LOW_DETVAL || (irr + icc) <= NGX_LOW_DETVAL
        || !(det == det)) {
        atomicOr(&sh->status, 1);
        return;
    }
    double idet = 1.0 / det;
    double drr = irr * idet, drc = irc * idet, dcc = icc * idet;
    double norm = 1.0 / (2 * 3.141592653589793 * sqrt(det));
    double pn = (p * norm) * area;
    if (compact) {
        // forward-difference record: row, col, a, b, c, pn
        double *c = compact + FDREC * k;
        c[0] = row; c[1] = col;
        c[2] = -0.5 * dcc; c[3] = -0.5 * drr;
        c[4] = drc; c[5] = pn;
    }
    if (!mix) return;
    double *g = mix + REC * k + 16;
    double ha = -0.25 * dcc, hb = -0.25 * drr;
    g[0] = row; g[1] = col;
    g[2] = -0.5 * dcc; g[3] = -0.5 * drr;
    g[4] = 0.5 * drc; g[5] = drc;
    g[6] = 0.5 * (irr0 + icc0); g[7] = 0.25 * drc;
    g[8] = ha; g[9] = hb;
    g[10] = hb - ha; g[11] = hb + ha;
    // chi^2 < 25 implies |dv| < 5 sqrt(irr), |du| < 5 sqrt(icc)
    g[12] = 5.0 * sqrt(irr); g[13] = 5.0 * sqrt(icc);
    fill_rot_table(mix + REC * k, pn);
}

// thread k < ngauss owns composed gaussian k
static __device__ __forceinline__ void build_mixture(const LMArgs &a,
                                                     LMShared *sh,
                                                     const double *x,
                                                     double *mix, double area,
                                                     double *compact = nullptr) {
    build_component(a, sh, x, threadIdx.x, mix, area, compact);
}

// per tile, the bit mask of the gaussians of mixture gc whose footprint can
// reach it (slack widens the test when one mask also serves the
// forward-difference mixtures)
__device__ __forceinline__ void build_tile_masks(const LMArgs &a,
                                                 const LMShared *sh,
                                                 const double *mix,
                                                 unsigned *tmask,
                                                 double slack) {
    const int ntile = a.nq >> 5;
    const int nwords = (a.ngauss + 31) >> 5;
    for (int t = threadIdx.x; t < ntile; t += blockDim.x) {
        double vc, uc, hv, hu;
        tile_extent(sh->jac, t, a.tpr, a.nrow, a.ncol, &vc, &uc, &hv, &hu);
        hv += slack * (fabs(vc) + hv);
        hu += slack * (fabs(uc) + hu);
        for (int w = 0; w < nwords; w++) {
            unsigned m = 0;
            int kend = min(32, a.ngauss - 32 * w);
            for (int kk = 0; kk < kend; kk++) {
                const double *g = mix + REC * (32 * w + kk) + 16;
                double rv = g[12], ru = g[13];
                rv += slack * (rv + fabs(g[0]));
                ru += slack * (ru + fabs(g[1]));
                if (tile_hit(vc, uc, hv, hu, g[0], g[1], rv, ru))
                    m |= 1u << kk;
            }
            tmask[(size_t)t * nwords + w] = m;
        }
    }
}

// pixel coordinates of tiled index q from the row/col offset tables.
// Padding slots (beyond the stamp's edge) take the coordinates of the
// nearest real pixel; their val/ierr are zero, so they add exact zeros.
__device__ __forceinline__ void coords_of(const LMArgs &a, const LMShared *sh,
                                          const double *rowd,
                                          const double *cold, int q,
                                          double *v, double *u) {
    int row, col;
    q_to_rowcol(q, a.tpr, &row, &col);
    double rd = rowd[min(row, a.nrow - 1)], cd = cold[min(col, a.ncol - 1)];
    *v = sh->jac.dvdrow * rd + sh->jac.dvdcol * cd;
    *u = sh->jac.dudrow * rd + sh->jac.dudcol * cd;
}

// weight of tiled slot q straight from the caller's weight map (L2-resident:
// the stamp was prefetched and read once when it was staged); 0 for padding.
// Only the pixel values live in shared memory: ierr = sqrt(weight) is
// recomputed where it is used, which halves the footprint of a stamp and
// lets a fifth CTA fit on the SM.
__device__ __forceinline__ double weight_of(const LMArgs &a, const double *wt,
                                            int q) {
    int row, col;
    q_to_rowcol(q, a.tpr, &row, &col);
    double w = 0.0;
    if (q < a.nq && row < a.nrow && col < a.ncol)
        w = __ldg(&wt[row * a.ncol + col]);
    return w;
}
__device__ __forceinline__ double ierr_of(double w) {
    return w > 0.0 ? sqrt(w) : 0.0;
}
// pixel value of tiled slot q from the caller's image (objects with several
// stamps keep nothing resident); excluded pixels read as exact zeros, like
// the staged copy
__device__ __forceinline__ double pixel_of(const LMArgs &a, const double *im,
                                           int q, double w) {
    int row, col;
    q_to_rowcol(q, a.tpr, &row, &col);
    double d = 0.0;
    if (q < a.nq && row < a.nrow && col < a.ncol
        && !(a.ignore_zero_weight && w <= 0.0))
        d = __ldg(&im[row * a.ncol + col]);
    return d;
}

// union of the masks of the C consecutive tiles starting at tile t0
template <int C>
__device__ __forceinline__ unsigned tiles_mask(const unsigned *tmask, int t0,
                                               int ntile, int nwords, int w) {
    unsigned m = 0;
#pragma unroll
    for (int c = 0; c < C; c++)
        if (t0 + c < ntile) m |= tmask[(size_t)(t0 + c) * nwords + w];
    return m;
}

// ---------------------------------------------------------------------
// value pass: sums = {sum r^2, s2n_numer, s2n_denom, r_first^2}.
// Each warp takes CV consecutive tiles per chunk.
template <int CV, int SRC>
__device__ __forceinline__ void value_pass(const LMArgs &a,
                                           const LMShared *sh,
                                           const double *mix,
                                           const unsigned *tmask,
                                           const double *val,
                                           const double *ierr,
                                           const double *wt,
                                           const double *im,
                                           const double *rowd,
                                           const double *cold,
                                           double (&sums)[4],
                                           double *mstore = nullptr) {
    const int T = blockDim.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ntile = a.nq >> 5;
    const int nwords = (a.ngauss + 31) >> 5;
    sums[0] = sums[1] = sums[2] = sums[3] = 0.0;
    const int nwarps_ = T >> 5;
    int wslot = warp;
    for (int base = 0; base < a.nq; base += T * CV) {
        const int q0 = base + wslot * CV * 32 + lane;
        wslot += a.rot_v;
        if (wslot >= nwarps_) wslot -= nwarps_;
        double v[CV], u[CV], acc[CV], wgt[CV], dval[CV];
#pragma unroll
        for (int c = 0; c < CV; c++) {
            coords_of(a, sh, rowd, cold, q0 + 32 * c, &v[c], &u[c]);
            wgt[c] = SRC >= 1 ? weight_of(a, wt, q0 + 32 * c) : 0.0;
            dval[c] = SRC == 2 ? pixel_of(a, im, q0 + 32 * c, wgt[c]) : 0.0;
            acc[c] = 0.0;
        }
        for (int w = 0; w < nwords; w++) {
            unsigned m = tiles_mask<CV>(tmask, q0 >> 5, ntile, nwords, w);
#pragma unroll 1
            while (m) {
                const int k = 32 * w + __ffs((int)m) - 1;
                m &= m - 1;
                const double *tab = mix + REC * k;
                const double2 *g2 = (const double2 *)(tab + 16);
                const double2 rc = g2[0], ab = g2[1], cc = g2[2];
                double x[CV], tv[CV], pv[CV];
                bool ap[CV], anyap = false;
#pragma unroll
                for (int c = 0; c < CV; c++) {
                    double dv = v[c] - rc.x;
                    double du = u[c] - rc.y;
                    double t = fma(cc.y, du, ab.x * dv);
                    x[c] = fma(ab.y * du, du, dv * t);
                    unsigned slot;
                    bool inr;
                    double f = fexp_split(x[c], &slot, &inr, &ap[c]);
                    tv[c] = tab[slot];
                    pv[c] = fexp_poly_c(f);
                    anyap |= ap[c];
                }
                if (__any_sync(0xffffffffu, anyap)) {
                    // some lane of some pixel sits in the apodization band
#pragma unroll
                    for (int c = 0; c < CV; c++) {
                        double e = tv[c] * pv[c];
                        double w = apod_window_fast(x[c]);
                        if (ap[c]) e *= w;
                        acc[c] += e;
                    }
                } else {
                    // table entry times polynomial folded into the sum
#pragma unroll
                    for (int c = 0; c < CV; c++)
                        acc[c] = fma(tv[c], pv[c], acc[c]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < CV; c++) {
            int q = q0 + 32 * c;
            if (q < a.nq) {
                // forward-difference mode keeps the model of the last
                // evaluated point: MINPACK differences against it
                if (mstore) mstore[q] = acc[c];
                double ie = SRC >= 1 ? ierr_of(wgt[c]) : ierr[q];
                double d = SRC == 2 ? dval[c] : val[q];
                double ivar = ie * ie;
                double diff = acc[c] - d;
                double r = diff * ie;
                sums[0] = fma(r, r, sums[0]);
                sums[1] = fma(d * acc[c], ivar, sums[1]);
                sums[2] = fma(acc[c] * acc[c], ivar, sums[2]);
                if (q == sh->first_q) sums[3] = r * r;
            }
        }
    }
}

// ---------------------------------------------------------------------
// derivative pass for the simple models (gauss/exp/dev): per pixel the six
// images of deriv_images, restructured so that the three covariance
// derivatives share the moment sums MA, MB, MS.  With
// hq = -Q delta / 2 the per-gaussian work is
//   S0 += val            S1 += valc hqv        S2 += valc hqu
//   MA += h valc ((hqu^2 + b/2) - (hqv^2 + a/2))
//   MS += h valc ((hqu^2 + b/2) + (hqv^2 + a/2))
//   MB += h valc (hqv hqu + c/4)
// sums = {A (21, packed lower triangle row by row), b (6), sum r^2,
//         s2n_numer, s2n_denom, r_first^2}
#define NDSUM 31
template <int CD, int SRC>
__device__ __forceinline__ void deriv_pass(const LMArgs &a,
                                           const LMShared *sh,
                                           const double *x,
                                           const double *mix,
                                           const unsigned *tmask,
                                           const double *val,
                                           const double *ierr,
                                           const double *wt,
                                           const double *im,
                                           const double *rowd,
                                           const double *cold,
                                           double (&sums)[NDSUM]) {
    const int T = blockDim.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ntile = a.nq >> 5;
    const int nwords = (a.ngauss + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NDSUM; i++) sums[i] = 0.0;
    // d(e1,e2)/d(g1,g2), results.py:988-995
    double g1 = x[2], g2 = x[3], Tpar = x[4], flux = x[5];
    double gsq = g1 * g1 + g2 * g2;
    double ff = 2.0 / (1.0 + gsq);
    double dfac = -ff / (1.0 + gsq);
    double de1dg1 = ff + 2.0 * g1 * g1 * dfac;
    double de1dg2 = 2.0 * g1 * g2 * dfac;
    double de2dg1 = de1dg2;
    double de2dg2 = ff + 2.0 * g2 * g2 * dfac;
    double e1, e2;
    g1g2_to_e1e2(g1, g2, &e1, &e2);
    double iT = 1.0 / Tpar, iflux = 1.0 / flux;

    const int nwarps_ = T >> 5;
    int wslot = warp;
    for (int base = 0; base < a.nq; base += T * CD) {
        const int q0 = base + wslot * CD * 32 + lane;
        wslot += a.rot_d;
        if (wslot >= nwarps_) wslot -= nwarps_;
        double v[CD], u[CD], wgt[CD], dval[CD];
        double S0[CD], S1[CD], S2[CD], MA[CD], MB[CD], MS[CD];
#pragma unroll
        for (int c = 0; c < CD; c++) {
            coords_of(a, sh, rowd, cold, q0 + 32 * c, &v[c], &u[c]);
            wgt[c] = SRC >= 1 ? weight_of(a, wt, q0 + 32 * c) : 0.0;
            dval[c] = SRC == 2 ? pixel_of(a, im, q0 + 32 * c, wgt[c]) : 0.0;
            S0[c] = S1[c] = S2[c] = MA[c] = MB[c] = MS[c] = 0.0;
        }
        for (int w = 0; w < nwords; w++) {
            unsigned m = tiles_mask<CD>(tmask, q0 >> 5, ntile, nwords, w);
#pragma unroll 1
            while (m) {
                const int k = 32 * w + __ffs((int)m) - 1;
                m &= m - 1;
                const double *tab = mix + REC * k;
                const double2 *gg = (const double2 *)(tab + 16);
                const double2 rc = gg[0], ab = gg[1], cc = gg[2];
                double xx[CD], hqv[CD], hqu[CD], vl[CD];
                bool ap[CD], anyap = false;
#pragma unroll
                for (int c = 0; c < CD; c++) {
                    double dv = v[c] - rc.x;
                    double du = u[c] - rc.y;
                    hqv[c] = fma(cc.x, du, ab.x * dv);     // -qv/2
                    hqu[c] = fma(ab.y, du, cc.x * dv);     // -qu/2
                    xx[c] = fma(du, hqu[c], dv * hqv[c]);  // -chi2/2
                    unsigned slot;
                    bool inr;
                    double f = fexp_split(xx[c], &slot, &inr, &ap[c]);
                    // exact zero outside the footprint (table slot 1)
                    vl[c] = tab[slot] * fexp_poly_c(f);
                    anyap |= ap[c];
                }
                const double2 hq = gg[3];           // h, qc
                const bool warp_apod = __any_sync(0xffffffffu, anyap);
#pragma unroll
                for (int c = 0; c < CD; c++) {
                    if (!(warp_apod && __any_sync(0xffffffffu, ap[c]))) {
                        const double2 hab = gg[4];  // ha, hb
                        S0[c] += vl[c];
                        S1[c] = fma(vl[c], hqv[c], S1[c]);
                        S2[c] = fma(vl[c], hqu[c], S2[c]);
                        double cv = vl[c] * hq.x;
                        double ea = fma(hqu[c], hqu[c], hab.y);
                        double eb = fma(hqv[c], hqv[c], hab.x);
                        // MA = EA - EB and MS = EA + EB are formed once per
                        // pixel after the loop (two adds less per gaussian)
                        MA[c] = fma(cv, ea, MA[c]);
                        MS[c] = fma(cv, eb, MS[c]);
                        MB[c] = fma(cv, fma(hqv[c], hqu[c], hq.y), MB[c]);
                    } else {
                        // a lane of this tile sits in the apodization
                        // band: the window slope enters the terms that flow
                        // through chi^2, derivs_nb.py:104-114
                        const double2 hab = gg[4];  // ha, hb
                        double wv = 1.0, wc = 1.0;
                        if (ap[c]) apod_pair(xx[c], &wv, &wc);
                        double vc = vl[c] * wc;
                        double vw = vl[c] * wv;
                        S0[c] += vw;
                        S1[c] = fma(vc, hqv[c], S1[c]);
                        S2[c] = fma(vc, hqu[c], S2[c]);
                        double cv = vc * hq.x, cw = vw * hq.x;
                        double qu2 = hqu[c] * hqu[c], qv2 = hqv[c] * hqv[c];
                        MA[c] = fma(cv, qu2, fma(cw, hab.y, MA[c]));
                        MS[c] = fma(cv, qv2, fma(cw, hab.x, MS[c]));
                        MB[c] = fma(cv, hqv[c] * hqu[c],
                                    fma(cw, hq.y, MB[c]));
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < CD; c++) {
            int q = q0 + 32 * c;
            if (q < a.nq) {
                double ie = SRC >= 1 ? ierr_of(wgt[c]) : ierr[q];
                double d = SRC == 2 ? dval[c] : val[q];
                // undo the factors folded into the accumulators:
                // qv = -2 hqv; (q q^T - W) terms carry 4, the cross term 8
                // (MA, MS hold EA = sum cv (hqu^2 + b/2), EB = sum cv
                // (hqv^2 + a/2) at this point)
                double ma = 4.0 * (MA[c] - MS[c]), ms = 4.0 * (MA[c] + MS[c]),
                       mb = 8.0 * MB[c];
                double j[6];
                j[0] = (-2.0 * S1[c]) * ie;
                j[1] = (-2.0 * S2[c]) * ie;
                j[2] = (0.5 * (de1dg1 * ma + de2dg1 * mb)) * ie;
                j[3] = (0.5 * (de1dg2 * ma + de2dg2 * mb)) * ie;
                j[4] = (0.5 * iT * (ms + e1 * ma + e2 * mb)) * ie;
                j[5] = S0[c] * (ie * iflux);
                double r = (S0[c] - d) * ie;
                int qq = 0;
#pragma unroll
                for (int aa = 0; aa < 6; aa++)
#pragma unroll
                    for (int bb = 0; bb <= aa; bb++) {
                        sums[qq] = fma(j[aa], j[bb], sums[qq]);
                        qq++;
                    }
#pragma unroll
                for (int aa = 0; aa < 6; aa++)
                    sums[21 + aa] = fma(j[aa], r, sums[21 + aa]);
                sums[27] = fma(r, r, sums[27]);
                double ivar = ie * ie;
                sums[28] = fma(d * S0[c], ivar, sums[28]);
                sums[29] = fma(S0[c] * S0[c], ivar, sums[29]);
                if (q == sh->first_q) sums[30] = r * r;
            }
        }
    }
}

// ---------------------------------------------------------------------
// forward-difference Jacobian (MINPACK fdjac2) for the models without
// analytic derivatives (bdf, bd) or analytic_jacobian = 0: evaluates the
// base mixture and the np perturbed mixtures per pixel.  The mixtures are
// built by build_mixture as compact records fdm[(np+1)][ngauss][FDREC]; the
// tile masks come from the base mixture (with slack for the 1.5e-8 relative
// perturbations).
// sums = {A packed (np(np+1)/2), b (np), sum r^2}
#define CF 2
// the per-thread sums of the forward-difference pass are sized for the
// single-stamp models (np <= 8), not for the solver's workspace
#define FD_MAXP 8
#define FD_B0_(P) ((P) * ((P) + 1) / 2)
#define FD_NSUM_(P) (FD_B0_(P) + (P) + 1)
template <int NPMAX, int SRC>
__device__ __forceinline__ void fd_pass(const LMArgs &a, const LMShared *sh,
                                        const double *hs, const double *fdm,
                                        const double *etab,
                                        const unsigned *tmask,
                                        const double *val, const double *ierr,
                                        const double *wt, const double *im,
                                        const double *rowd,
                                        const double *cold,
                                        double (&sums)[FD_NSUM_(NPMAX)],
                                        const double *mbase, int npl = -1) {
    const int T = blockDim.x;
    // npl: parameters perturbed in this pass when they are not the fit's
    // (a stamp of a multi-band object: the shared ones and its band's flux)
    const int np = npl >= 0 ? npl : a.np;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ntile = a.nq >> 5;
    const int nwords = (a.ngauss + 31) >> 5;
#pragma unroll
    for (int i = 0; i < FD_NSUM_(NPMAX); i++) sums[i] = 0.0;
    double hstep[NPMAX];
#pragma unroll
    for (int s = 0; s < NPMAX; s++) hstep[s] = s < np ? hs[s] : 1.0;
    const int nwarps_ = T >> 5;
    int wslot = warp;
    for (int base = 0; base < a.nq; base += T * CF) {
        const int q0 = base + wslot * CF * 32 + lane;
        wslot += a.rot_d;
        if (wslot >= nwarps_) wslot -= nwarps_;
        double v[CF], u[CF], wgt[CF], dval[CF];
#pragma unroll
        for (int c = 0; c < CF; c++) {
            coords_of(a, sh, rowd, cold, q0 + 32 * c, &v[c], &u[c]);
            wgt[c] = SRC >= 1 ? weight_of(a, wt, q0 + 32 * c) : 0.0;
            dval[c] = SRC == 2 ? pixel_of(a, im, q0 + 32 * c, wgt[c]) : 0.0;
        }
        double mod[NPMAX + 1][CF];
#pragma unroll
        for (int s = 0; s <= NPMAX; s++) {
#pragma unroll
            for (int c = 0; c < CF; c++) mod[s][c] = 0.0;
            if (s > np) continue;
            if (s == 0 && mbase) {
                // the base model is the last value pass's (the accepted
                // point: lmdif calls fdjac2 with its fvec)
#pragma unroll
                for (int c = 0; c < CF; c++) {
                    int q = q0 + 32 * c;
                    if (q < a.nq) mod[0][c] = mbase[q];
                }
                continue;
            }
            const double *recs = fdm + (size_t)s * a.ngauss * FDREC;
            for (int w = 0; w < nwords; w++) {
                unsigned m = tiles_mask<CF>(tmask, q0 >> 5, ntile, nwords, w);
#pragma unroll 1
                while (m) {
                    const int k = 32 * w + __ffs((int)m) - 1;
                    m &= m - 1;
                    const double2 *g2 =
                        (const double2 *)(recs + FDREC * k);
                    const double2 rc = g2[0], ab = g2[1], cp = g2[2];
                    double x[CF], tv[CF], pv[CF];
                    bool ap[CF], anyap = false;
#pragma unroll
                    for (int c = 0; c < CF; c++) {
                        double dv = v[c] - rc.x;
                        double du = u[c] - rc.y;
                        double t = fma(cp.x, du, ab.x * dv);
                        x[c] = fma(ab.y * du, du, dv * t);
                        unsigned slot;
                        bool inr;
                        double f = fexp_split(x[c], &slot, &inr, &ap[c]);
                        // (exp(i) pn) P(f): the product the per-gaussian
                        // tables of the other passes hold precomputed
                        tv[c] = etab[slot] * cp.y;
                        pv[c] = fexp_poly_c(f);
                        anyap |= ap[c];
                    }
                    if (__any_sync(0xffffffffu, anyap)) {
#pragma unroll
                        for (int c = 0; c < CF; c++) {
                            double e = tv[c] * pv[c];
                            double wv = apod_window_x(x[c]);
                            if (ap[c]) e *= wv;
                            mod[s][c] += e;
                        }
                    } else {
                        // as in the value pass, whose model the differences
                        // are taken against
#pragma unroll
                        for (int c = 0; c < CF; c++)
                            mod[s][c] = fma(tv[c], pv[c], mod[s][c]);
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < CF; c++) {
            int q = q0 + 32 * c;
            if (q < a.nq) {
                double ie = SRC >= 1 ? ierr_of(wgt[c]) : ierr[q];
                double d = SRC == 2 ? dval[c] : val[q];
                double r = (mod[0][c] - d) * ie;
                double j[NPMAX];
#pragma unroll
                for (int s = 0; s < NPMAX; s++) {
                    double rs = (mod[s + 1][c] - d) * ie;
                    j[s] = s < np ? (rs - r) / hstep[s] : 0.0;
                }
                // static layout: packed triangle of NPMAX, then b, then r^2
                // (the first np rows of the triangle are those of np)
                int qq = 0;
#pragma unroll
                for (int aa = 0; aa < NPMAX; aa++)
#pragma unroll
                    for (int bb = 0; bb <= aa; bb++) {
                        sums[qq] = fma(j[aa], j[bb], sums[qq]);
                        qq++;
                    }
#pragma unroll
                for (int aa = 0; aa < NPMAX; aa++)
                    sums[FD_B0_(NPMAX) + aa] =
                        fma(j[aa], r, sums[FD_B0_(NPMAX) + aa]);
                sums[FD_B0_(NPMAX) + NPMAX] =
                    fma(r, r, sums[FD_B0_(NPMAX) + NPMAX]);
            }
        }
    }
}

// generic block reduction of nsums values held in a local array
__device__ __forceinline__ void block_sum_n(double *v, int nv,
                                            double *scratch) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int nwarps = (blockDim.x + 31) >> 5;
    for (int i = 0; i < nv; i++) {
        double s = warp_sum(v[i]);
        if (lane == 0) scratch[i * nwarps + warp] = s;
    }
    __syncthreads();
    for (int i = 0; i < nv; i++) {
        double s = 0.0;
        for (int w = 0; w < nwarps; w++) s += scratch[i * nwarps + w];
        v[i] = s;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------
// prior contributions, thread 0
static __device__ __noinline__ double prior_fnorm2(const LMArgs &a, const double *x,
                                      int *status, int *ninf) {
    *ninf = 0;
    if (!a.prior.present) return 0.0;
    double f[NGX_MAX_PRIOR_ROWS];
    if (prior_fill_fdiff(a.model, a.prior, x, f)) {
        *status = 1;
        return 0.0;
    }
    int nr = prior_nrows(a.model, a.prior);
    double s = 0.0;
    for (int r = 0; r < nr; r++) {
        if (isinf(f[r])) *ninf += 1;
        else s += f[r] * f[r];
    }
    return s;
}

// the epilogue: run_leastsq (leastsqbound.py:70-126) + set_fit_result.
// Entered by the whole warp 0.  Writes straight into the global result
// record (lane-strided, coalesced); matrices are staged in the shared-memory
// workspace of the state (no per-thread local arrays).
template <typename RES>
__device__ __noinline__ void lm_finish(const LMArgs &a, LMShared *sh,
                                       RES *out) {
    LMState *s = &sh->st;
    const int np = a.np;
    const int lane = threadIdx.x & 31;
    RES &r = *out;
    constexpr int RP = (int)(sizeof(r.pars) / sizeof(double));
    const double qnan = nan("");
    double *cov0 = s->A;        // J^T J is no longer needed
    double *pcov = s->T1;       // after lm_cov0 is done with it
    int flags = 0;
    int have_cov = 0, have_pcov = 0, have_err = 0;
    const int info = s->info;
    if (info > 4) {
        flags |= 1 << (info - 5);
    } else {
        have_cov = lm_cov0(s, cov0);
        if (have_cov && a.lm.nbounds) {
            // fjac from internal to external parameters (leastsqbound.py:
            // 531-540): (J_ext^T J_ext)^-1 = G (J_int^T J_int)^-1 G
            __syncwarp();
#pragma unroll 1
            for (int i = lane; i < np * np; i += 32) {
                const int ra = i / np, rb = i - ra * np;
                cov0[i] *= b_grad(a.lm, ra, s->x[ra])
                    * b_grad(a.lm, rb, s->x[rb]);
            }
            __syncwarp();
        }
        if (!have_cov) {
            flags |= NGX_LM_SINGULAR_MATRIX;
        } else {
            int dof = sh->npix - np;
            if (dof == 0) {
                flags |= NGX_ZERO_DOF;
            } else {
                // fdiff[n_prior_pars:]: the simple priors fill one row fewer
                // than they reserve, which drops the first pixel row
                double ssq = sh->cur_ssq;
                if (prior_nslots(a.model, a.prior)
                    > prior_nrows(a.model, a.prior))
                    ssq -= sh->cur_rfirst2;
                double s_sq = ssq / dof;
                __syncwarp();
#pragma unroll 1
                for (int i = lane; i < np * np; i += 32)
                    pcov[i] = cov0[i] * s_sq;
                __syncwarp();
                have_pcov = 1;
                int cflags = lm_test_cov(np, pcov, s->T2);
                if (cflags) flags |= cflags;
                else have_err = 1;
            }
        }
    }
    __syncwarp();
#pragma unroll 1
    for (int i = lane; i < RP; i += 32) {
        double pv = 0.0, pe = 0.0;
        if (i < np) {
            pv = (info > 4) ? NGX_PDEF : sh->xcur[i];
            pe = have_err ? sqrt(pcov[i * np + i]) : NGX_CDEF;
        }
        r.pars[i] = pv;
        r.pars_err[i] = pe;
    }
#pragma unroll 1
    for (int i = lane; i < RP * RP; i += 32) {
        double c0 = 0.0, c1 = 0.0;
        if (i < np * np) {
            c0 = have_cov ? cov0[i] : qnan;
            c1 = have_pcov ? pcov[i] : NGX_CDEF;
        }
        r.pars_cov0[i] = c0;
        r.pars_cov[i] = c1;
    }
    if (lane != 0) return;
    r.flags = flags;
    r.nfev = s->nfev;
    r.njev = s->njev;
    r.ier = info;
    int npix = 0, dof = 0;
    double lnprob = qnan, num = qnan, den = qnan, chi2per = qnan, s2n = qnan;
    double Tv = qnan, Terr = qnan, flux = qnan, fluxerr = qnan;
    if (flags == 0) {
        double lnp_prior = 0.0;
        int bad = prior_lnprob(a.model, a.prior, sh->xcur, &lnp_prior);
        if (bad) {
            lnprob = -INFINITY;
            num = 0.0;
            den = NGX_BIGVAL;
            npix = 0;
        } else {
            lnprob = -0.5 * sh->cur_ssq + lnp_prior;
            num = sh->cur_num;
            den = sh->cur_den;
            npix = sh->npix;
        }
        s2n = 0.0;
        if (den > 0) s2n = num / sqrt(den);
        double chi2 = lnprob / (-0.5);
        dof = npix - np;
        chi2per = chi2 / dof;
        Tv = sh->xcur[4];
        Terr = sqrt(pcov[4 * np + 4]);
        // band 0 of a multi-band fit; the last parameter otherwise
        const int fi = np - (a.nband > 1 ? a.nband : 1);
        flux = sh->xcur[fi];
        fluxerr = sqrt(pcov[fi * np + fi]);
    }
    r.npix = npix;
    r.dof = dof;
    r.lnprob = lnprob; r.s2n_numer = num; r.s2n_denom = den;
    r.chi2per = chi2per; r.s2n = s2n;
    r.T = Tv; r.T_err = Terr; r.flux = flux; r.flux_err = fluxerr;
}

// ---------------------------------------------------------------------
// ---------------------------------------------------------------------
// assemble J^T J, J^T f from the pixel sums and the prior rows and feed the
// state machine (thread 0).  packed = lower triangle row by row, then b.
// pre: A and b already hold the pixel sums (assemble_normal_eq, run by the
// whole CTA); only the prior rows are added here.
static __device__ __noinline__ void feed_jacobian(const LMArgs &a, LMShared *sh,
                                           const double *sums, int status,
                                           int analytic, const double *hs,
                                           int pre = 0) {
    // entered by the whole warp 0: lane 0 adds the prior rows, then all
    // lanes advance the state machine (ngx_lm_solve.cuh)
    double *A = sh->st.A, *b = sh->st.b;
    const int np = a.np;
    if ((threadIdx.x & 31) == 0) {
    // the accepted point in external parameters (st.x is internal)
    double *x = sh->st.v[6];
    for (int i = 0; i < np; i++) x[i] = b_i2e(a.lm, i, sh->st.x[i]);
    int nr = prior_nrows(a.model, a.prior);
    const int status_in = status;
    double *f0 = sh->st.v[7];
    double *jp = sh->st.T1;
    if (!status && nr > 0) {
        if (prior_fill_fdiff(a.model, a.prior, x, f0)) {
            status = 1;
        } else if (analytic) {
            if (prior_jacobian(a.model, a.prior, x, np, f0, jp)) status = 1;
        } else {
            // the prior rows are part of fvec: MINPACK's own forward
            // difference with its step hs
            double *xp = sh->st.v[8], *fs = sh->st.v[9];
            for (int s = 0; s < np && !status; s++) {
                for (int i = 0; i < np; i++) xp[i] = x[i];
                // lmdif steps the internal parameter
                xp[s] = b_i2e(a.lm, s, sh->st.x[s] + hs[s]);
                if (prior_fill_fdiff(a.model, a.prior, xp, fs)) {
                    status = 1;
                    break;
                }
                for (int r = 0; r < nr; r++)
                    jp[r * np + s] = (fs[r] - f0[r]) / hs[s];
            }
        }
    }
    // with the sums in place and nothing to add (no prior, or a status that
    // was already known when they were placed) the assembly is done
    if (!(pre && status == status_in && (nr == 0 || status))) {
        int q = 0;
        for (int aa = 0; aa < np; aa++)
            for (int bb = 0; bb <= aa; bb++) {
                double sv = status ? 0.0 : sums[q];
                if (!status)
                    for (int r = 0; r < nr; r++)
                        sv += jp[r * np + aa] * jp[r * np + bb];
                A[aa + LM_MAXP * bb] = sv;
                A[bb + LM_MAXP * aa] = sv;
                q++;
            }
        for (int aa = 0; aa < np; aa++) {
            double sv = status ? 0.0 : sums[q + aa];
            if (!status)
                for (int r = 0; r < nr; r++)
                    if (isfinite(f0[r])) sv += jp[r * np + aa] * f0[r];
            b[aa] = sv;
        }
    }
    if (analytic && a.lm.nbounds) {
        // wDfun = Dfun(i2e(x)) * d(external)/d(internal), leastsqbound.py:
        // 481-486: J^T J and J^T f pick up the gradients of their columns
        double *g = sh->st.v[8];
        for (int i = 0; i < np; i++) g[i] = b_grad(a.lm, i, sh->st.x[i]);
        for (int aa = 0; aa < np; aa++) {
            for (int bb = 0; bb < np; bb++)
                A[aa + LM_MAXP * bb] *= g[aa] * g[bb];
            b[aa] *= g[aa];
        }
    }
    }
    __syncwarp();
    lm_feed_j(&sh->st, A, b, analytic ? 0 : np);
}

// J^T J (symmetric, ld LM_MAXP) and J^T f of a six-parameter fit from the
// packed block totals, one entry per thread; zeros when the Jacobian is
// undefined.  The solver thread then only adds the prior rows.
__device__ __forceinline__ void assemble_normal_eq(LMShared *sh,
                                                   const double *tot,
                                                   int jstatus) {
    const int t = threadIdx.x;
    if (t < 21) {
        const int aa = t >= 15 ? 5 : t >= 10 ? 4 : t >= 6 ? 3 : t >= 3 ? 2
            : t >= 1 ? 1 : 0;
        const int bb = t - aa * (aa + 1) / 2;
        const double sv = jstatus ? 0.0 : tot[t];
        sh->st.A[aa + LM_MAXP * bb] = sv;
        sh->st.A[bb + LM_MAXP * aa] = sv;
    } else if (t < 27) {
        sh->st.b[t - 21] = jstatus ? 0.0 : tot[t];
    }
}

// feed |f| of a residual pass (whole warp 0; lane 0 evaluates the prior);
// returns 1 on every lane if the point is adopted
static __device__ __noinline__ int feed_residual(const LMArgs &a, LMShared *sh,
                                          const double *xeval, double ssq,
                                          int status) {
    double fnorm = 0.0;
    if ((threadIdx.x & 31) == 0) {
        int ninf = 0;
        double ps = prior_fnorm2(a, xeval, &status, &ninf);
        if (status) fnorm = nan("");
        else if (ninf == 1) fnorm = INFINITY;
        else if (ninf > 1) fnorm = nan("");
        else fnorm = sqrt(ssq + ps);
    }
    fnorm = __shfl_sync(0xffffffffu, fnorm, 0);
    return lm_feed_f(&sh->st, fnorm);
}

// ---------------------------------------------------------------------
// MODE 0: analytic Jacobian (gauss/exp/dev), np = 6
// MODE 1: forward differences (any model), np <= 8
// WG: the weights stay in global memory / L2 (stamps too large to keep both
// planes in shared memory at four CTAs per SM); otherwise ierr = sqrt(w) is
// staged next to the pixel values.
// GS: the pixel-value plane (and the base-model plane of the forward-
// difference mode) live in a per-CTA slice of a global scratch buffer
// instead of shared memory, for stamps too large to be resident (WG only).
// FDP: parameter capacity of the forward-difference pass (8 for the galaxy
// models, 12 for coelliptical mixtures); RES: the result record type.
template <int MODE, int CV, int CD, int MINB, bool WG, int NT = 128,
          bool GS = false, int FDP = FD_MAXP, typename RES = ngx_lm_result>
__global__ void __launch_bounds__(NT, MINB)
k_lm_fit(LMArgs a) {
    static_assert(!GS || WG, "GS keeps the weights in global memory");
    extern __shared__ double smem_d[];
    __shared__ LMShared sh;
    const int T = blockDim.x;
    const int nwarps = T >> 5;
    const int nfd = (MODE == 0) ? 0 : (a.np + 1);
    // dynamic shared memory carve-up
    double *mix = smem_d;                                    // ngauss*REC
    double *fdm = mix + (size_t)a.ngauss * REC;              // nfd*ngauss*FDREC
    double *etab = fdm + (size_t)nfd * a.ngauss * FDREC;     // 16 (MODE 1)
    double *planes = etab + (nfd ? 16 : 0);
    double *gplanes = GS ? a.gscratch
        + (size_t)blockIdx.x * ((nfd ? 2 : 1) * (size_t)a.nq) : planes;
    double *mplane = gplanes;                                // nq (MODE 1)
    double *val = mplane + (nfd ? a.nq : 0);                 // nq
    double *ierr = val + a.nq;                               // nq (!WG)
    double *rowd = GS ? planes : ierr + (WG ? 0 : a.nq);
    double *cold = rowd + a.nrow;
    double *scratch = cold + a.ncol;                         // 48 * nwarps
    double *hs = scratch + 48 * nwarps;                      // 8
    unsigned *tmask = (unsigned *)(hs + LM_MAXP);            // ntile*nwords

    if (nfd) fill_rot_table(etab, 1.0);     // every thread, same values
    // Work queue: stamps are handed out by an atomic counter, one ahead, so
    // the stamp after the current one is known early and can be pulled into
    // L2 while this one is being fitted.
    if (threadIdx.x == 0) sh.stamp = atomicAdd(a.counter, 1);
    __syncthreads();
    int stamp = sh.stamp;
    int pending = -1;       // stamp whose epilogue (lm_finish) is still owed
    const int warp_id = threadIdx.x >> 5;

    for (;;) {
        const bool have = stamp < a.n;
        const bool skip = have && a.todo && a.todo[stamp] == 0;
#ifdef NGX_PROF
        long long pt0 = clock64(), pt_mix = 0, pt_pass = 0, pt_solve = 0,
                  pt_stage = 0, pt_fin = 0, ptx;
#define PROF_T(x) do { ptx = clock64(); x += ptx - pt0; pt0 = ptx; } while (0)
#else
#define PROF_T(x)
#endif
        // ---- warp 0: the epilogue of the previous stamp (one thread);
        //      the other warps meanwhile stage the next stamp, which only
        //      touches val/ierr and a few words the epilogue does not read
        int cnt = 0, first = a.ntot, firstq = 0;
        if (warp_id == 0) {
            if (pending >= 0) {
                lm_finish(a, &sh, &((RES *)a.res)[pending]);
                if (a.done_count) {
                    __threadfence();
                    __syncwarp();
                    if (threadIdx.x == 0) mark_done(a, pending);
                }
            }
            if (threadIdx.x == 0 && have) {
                sh.next = atomicAdd(a.counter, 1);
                if (a.ready) wait_ready(a, stamp);
                sh.jac = a.jac[stamp];
                sh.status = 0;
            }
        } else if (have && !skip) {
            if (a.ready) wait_ready(a, stamp);
            const double *im = a.images + (size_t)stamp * a.ntot;
            const double *wt = a.weights + (size_t)stamp * a.ntot;
            const int t0 = threadIdx.x - 32, TS = T - 32;
            if (WG) {
                // the pixel values go straight into their tiled shared-memory
                // slots as asynchronous 8-byte copies (LDGSTS); the weights are
                // only inspected here (valid-pixel count, first valid pixel)
                for (int q = t0; q < a.nq; q += TS) {
                    int row, col;
                    q_to_rowcol(q, a.tpr, &row, &col);
                    if (row < a.nrow && col < a.ncol) {
                        if (GS)
                            val[q] = im[row * a.ncol + col];
                        else
                            __pipeline_memcpy_async(
                                &val[q], &im[row * a.ncol + col],
                                sizeof(double));
                    }
                }
                __pipeline_commit();
                if (t0 < 6 * a.npsf)
                    sh.psf[t0] = a.psf[(size_t)stamp * 6 * a.npsf + t0];
                unsigned long long invalid = 0;     // per-thread slots, <= 64
                int slot = 0;
    #pragma unroll 4
                for (int q = t0; q < a.nq; q += TS, slot++) {
                    int row, col;
                    q_to_rowcol(q, a.tpr, &row, &col);
                    bool valid = false;
                    if (row < a.nrow && col < a.ncol) {
                        int p = row * a.ncol + col;
                        double w = __ldg(&wt[p]);
                        valid = !(a.ignore_zero_weight && w <= 0.0);
                        if (valid) {
                            cnt++;
                            if (p < first) { first = p; firstq = q; }
                        }
                    }
                    if (!valid && slot < 64) invalid |= 1ull << slot;
                }
                __pipeline_wait_prior(0);
                // pixels that do not take part hold an exact zero
                slot = 0;
                for (int q = t0; q < a.nq; q += TS, slot++) {
                    if (slot >= 64 || ((invalid >> slot) & 1ull)) {
                        int row, col;
                        q_to_rowcol(q, a.tpr, &row, &col);
                        bool valid = false;
                        if (slot >= 64 && row < a.nrow && col < a.ncol) {
                            double w = __ldg(&wt[row * a.ncol + col]);
                            valid = !(a.ignore_zero_weight && w <= 0.0);
                        }
                        if (!valid) val[q] = 0.0;
                    }
                }
            } else if (a.src != 0) {
                // float stamps straight from the host's format, widened here
                // (exactly): asynchronous 4-byte copies into the low half of
                // each pixel's shared-memory slot, converted in place by the
                // thread that fetched them - no intermediate double planes
                // in HBM
                const float *im32 = a.im32 + (size_t)stamp * a.ntot;
                const float *wt32 = a.wt32 + (size_t)stamp * a.ntot;
                const unsigned char *mb = (a.src == 2 && a.maskbits)
                    ? a.maskbits + (size_t)stamp * a.nmb : nullptr;
                for (int q = t0; q < a.nq; q += TS) {
                    int row, col;
                    q_to_rowcol(q, a.tpr, &row, &col);
                    if (row < a.nrow && col < a.ncol) {
                        int p = row * a.ncol + col;
                        __pipeline_memcpy_async(&val[q], &im32[p],
                                                sizeof(float));
                        if (a.src == 1)
                            __pipeline_memcpy_async(&ierr[q], &wt32[p],
                                                    sizeof(float));
                    }
                }
                __pipeline_commit();
                const double ws = a.src == 2 ? a.wscalar[stamp] : 0.0;
                if (t0 < 6 * a.npsf)
                    sh.psf[t0] = a.psf[(size_t)stamp * 6 * a.npsf + t0];
                __pipeline_wait_prior(0);
                for (int q = t0; q < a.nq; q += TS) {
                    int row, col;
                    q_to_rowcol(q, a.tpr, &row, &col);
                    double vv = 0.0, ie = 0.0;
                    if (row < a.nrow && col < a.ncol) {
                        const int p = row * a.ncol + col;
                        double w;
                        if (a.src == 1)
                            w = (double)*(const float *)&ierr[q];
                        else
                            w = (!mb || ((mb[p >> 3] >> (7 - (p & 7))) & 1))
                                ? ws : 0.0;
                        const bool valid = !(a.ignore_zero_weight && w <= 0.0);
                        if (valid) {
                            vv = (double)*(const float *)&val[q];
                            ie = sqrt(w < 0.0 ? 0.0 : w);
                            cnt++;
                            if (p < first) { first = p; firstq = q; }
                        }
                    }
                    val[q] = vv;
                    ierr[q] = ie;
                }
            } else {
                // both planes as asynchronous 8-byte copies (LDGSTS) straight
                // into their tiled shared-memory slots, so the loads of a
                // thread overlap instead of each waiting on DRAM
                for (int q = t0; q < a.nq; q += TS) {
                    int row, col;
                    q_to_rowcol(q, a.tpr, &row, &col);
                    if (row < a.nrow && col < a.ncol) {
                        int p = row * a.ncol + col;
                        __pipeline_memcpy_async(&val[q], &im[p],
                                                sizeof(double));
                        __pipeline_memcpy_async(&ierr[q], &wt[p],
                                                sizeof(double));
                    }
                }
                __pipeline_commit();
                if (t0 < 6 * a.npsf)
                    sh.psf[t0] = a.psf[(size_t)stamp * 6 * a.npsf + t0];
                __pipeline_wait_prior(0);
                // each thread converts the entries it fetched itself
                for (int q = t0; q < a.nq; q += TS) {
                    int row, col;
                    q_to_rowcol(q, a.tpr, &row, &col);
                    double vv = 0.0, ie = 0.0;
                    if (row < a.nrow && col < a.ncol) {
                        int p = row * a.ncol + col;
                        double w = ierr[q];
                        bool valid = !(a.ignore_zero_weight && w <= 0.0);
                        if (valid) {
                            vv = val[q];
                            ie = sqrt(w < 0.0 ? 0.0 : w);
                            cnt++;
                            if (p < first) { first = p; firstq = q; }
                        }
                    }
                    val[q] = vv;
                    ierr[q] = ie;
                }
            }
            const int lane = threadIdx.x & 31;
            for (int o = 16; o > 0; o >>= 1) {
                cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
                int of = __shfl_xor_sync(0xffffffffu, first, o);
                int oq = __shfl_xor_sync(0xffffffffu, firstq, o);
                if (of < first) { first = of; firstq = oq; }
            }
            if (lane == 0) {
                sh.warp_i[warp_id] = cnt;
                sh.warp_j[warp_id] = first;
                sh.warp_k[warp_id] = firstq;
            }
        }
        if (!have) break;
        pending = -1;
        __syncthreads();
        const int next = sh.next;
        if (skip) {
            // nothing to do for this stamp (its record stays as it is)
            if (threadIdx.x == 0) mark_done(a, stamp);
            __syncthreads();
            stamp = next;
            continue;
        }
        if (next < a.n && (!a.ready || chunk_arrived(a, next))) {
            // pull the next stamp into L2 while this one is fitted
            const char *pi = (const char *)(a.images + (size_t)next * a.ntot);
            const char *pw = (const char *)(a.weights + (size_t)next * a.ntot);
            size_t bytes = (size_t)a.ntot * sizeof(double);
            if (a.src != 0) {
                pi = (const char *)(a.im32 + (size_t)next * a.ntot);
                pw = a.src == 1 ? (const char *)(a.wt32 + (size_t)next * a.ntot)
                                : pi;
                bytes = (size_t)a.ntot * sizeof(float);
            }
            for (size_t off = (size_t)threadIdx.x * 128; off < bytes;
                 off += (size_t)T * 128) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pi + off));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pw + off));
            }
        }
        for (int p = threadIdx.x; p < a.nrow; p += T)
            rowd[p] = p - sh.jac.row0;
        for (int p = threadIdx.x; p < a.ncol; p += T)
            cold[p] = p - sh.jac.col0;
        if (threadIdx.x == 0) {
            int c = 0, f = a.ntot, fq = -1;
            for (int w = 1; w < nwarps; w++) {
                c += sh.warp_i[w];
                if (sh.warp_j[w] < f) { f = sh.warp_j[w]; fq = sh.warp_k[w]; }
            }
            sh.npix = c;
            sh.first_q = fq;
            // psf centroid and flux, gmix_get_cen :108-130
            double prow = 0.0, pcol = 0.0, psum = 0.0;
            for (int k = 0; k < a.npsf; k++) {
                double p = sh.psf[6 * k];
                prow += p * sh.psf[6 * k + 1];
                pcol += p * sh.psf[6 * k + 2];
                psum += p;
            }
            if (a.npsf > 0) {
                sh.psf_row = prow / psum;
                sh.psf_col = pcol / psum;
                sh.psf_ipsum = 1.0 / psum;
            }
            lm_init(&sh.st, a.np, a.guess + (size_t)stamp * a.np,
                    a.lm.maxfev, a.lm.ftol, a.lm.xtol);
            sh.action = sh.st.action;
            sh.spec = 1;
            for (int i = 0; i < a.np; i++) {
                const double xe = sh.st.xtrial[i];
                sh.xeval[i] = sh.xcur[i] = xe;
                sh.st.x[i] = sh.st.xtrial[i] = b_e2i(a.lm, i, xe);
            }
            sh.cur_ssq = sh.cur_rfirst2 = sh.cur_num = sh.cur_den = 0.0;
        }
        __syncthreads();
        const double area = sh.jac.scale * sh.jac.scale;
        const double *wts = a.weights + (size_t)stamp * a.ntot;

        if (sh.npix == 0) {
            // GMixFatalError("no weights > 0") at Observation construction
            if (threadIdx.x == 0) {
                RES *r = &((RES *)a.res)[stamp];
                memset(r, 0, sizeof(*r));
                r->flags = NGX_NO_ATTEMPT;
                for (int i = 0; i < a.np; i++) {
                    r->pars[i] = NGX_PDEF;
                    r->pars_err[i] = NGX_CDEF;
                }
                mark_done(a, stamp);
            }
            __syncthreads();
            stamp = next;
            continue;
        }

        PROF_T(pt_stage);
        // ---- the fit
        while (sh.action != LM_ACT_DONE) {
            const int action = sh.action;
            const int spec = sh.spec;
            __syncthreads();
            if (threadIdx.x == 0) sh.status = 0;
            __syncthreads();
            // the point this pass evaluates: the trial point for a residual
            // evaluation, the accepted point for a Jacobian
            const bool want_f = (action == LM_ACT_EVAL_F);
            const double *xp = want_f ? sh.xeval : sh.xcur;
            // Speculation: when the previous trial was accepted (and at the
            // very first evaluation) the residual pass also produces the
            // Jacobian at the trial point, which MINPACK asks for next if the
            // step is accepted.  One fused pass costs 30 fp64 instr per
            // gaussian-pixel instead of 16 + 30, halves the serial sections
            // per iteration, and leaves nfev/njev/iterates unchanged; a
            // rejected step just discards the extra sums.
            // No speculation when the step's predicted reduction is already
            // below ftol: that evaluation almost always ends the fit, and a
            // value pass costs 18 instead of 30 fp64 instr per gaussian-pixel.
            const bool fused = want_f && spec && xp[4] != 0.0 && xp[5] != 0.0
                && !(sh.st.prered <= sh.st.ftol);
            if (MODE == 0 && (fused || action == LM_ACT_EVAL_J)) {
                build_mixture(a, &sh, xp, mix, area);
                __syncthreads();
                const int mix_status = sh.status;
                if (!mix_status) build_tile_masks(a, &sh, mix, tmask, 1.0e-9);
                __syncthreads();
                PROF_T(pt_mix);
                double sums[NDSUM];
#pragma unroll
                for (int i = 0; i < NDSUM; i++) sums[i] = 0.0;
                // calc_jacobian raises for T == 0 or flux == 0
                int jstatus = mix_status || xp[4] == 0.0 || xp[5] == 0.0;
                if (!jstatus)
                    deriv_pass<CD, WG ? 1 : 0>(a, &sh, xp, mix, tmask, val, ierr, wts,
                                               nullptr, rowd,
                                   cold, sums);
                block_sum_to<NDSUM>(sums, scratch, sh.tot);
                // J^T J of the previous Jacobian is dead by now (its
                // factor R is what the solver keeps), so the new sums can
                // go in place before the step is known to be accepted
                assemble_normal_eq(&sh, sh.tot, jstatus);
                __syncthreads();
                PROF_T(pt_pass);
                if (warp_id == 0) {
                    // the whole warp advances the state machine
                    const double *tot = sh.tot;
                    const bool lane0 = threadIdx.x == 0;
                    if (want_f) {
                        int acc = feed_residual(a, &sh, xp, tot[27],
                                                mix_status);
                        if (acc) {
                            if (lane0) {
                                sh.cur_ssq = tot[27];
                                sh.cur_num = tot[28];
                                sh.cur_den = tot[29];
                                sh.cur_rfirst2 = tot[30];
                            }
                            if (sh.st.action == LM_ACT_EVAL_J)
                                feed_jacobian(a, &sh, tot, jstatus, 1, hs, 1);
                        }
                        if (lane0) sh.spec = acc;
                    } else {
                        feed_jacobian(a, &sh, tot, jstatus, 1, hs, 1);
                    }
                }
            } else if (action == LM_ACT_EVAL_F) {
                build_mixture(a, &sh, sh.xeval, mix, area);
                __syncthreads();
                if (!sh.status) build_tile_masks(a, &sh, mix, tmask, 1.0e-9);
                __syncthreads();
                PROF_T(pt_mix);
                double sums[4] = {0, 0, 0, 0};
                if (!sh.status)
                    value_pass<CV, WG ? 1 : 0>(a, &sh, mix, tmask, val, ierr, wts,
                                               nullptr, rowd, cold, sums,
                                               MODE == 1 ? mplane : nullptr);
                block_sum_to<4>(sums, scratch, sh.tot);
                PROF_T(pt_pass);
                if (warp_id == 0) {
                    const double *tot = sh.tot;
                    int acc = feed_residual(a, &sh, sh.xeval, tot[0],
                                            sh.status);
                    if (threadIdx.x == 0) {
                        if (acc) {
                            sh.cur_ssq = tot[0];
                            sh.cur_num = tot[1];
                            sh.cur_den = tot[2];
                            sh.cur_rfirst2 = tot[3];
                        }
                        sh.spec = acc;
                    }
                }
            } else {
                // forward differences: base + np perturbed mixtures
                const int np = a.np;
                if (threadIdx.x == 0) {
                    const double eps = 1.4901161193847656e-08;  // sqrt(epsmch)
                    for (int s = 0; s < np; s++) {
                        double h = eps * fabs(sh.st.x[s]);
                        if (h == 0.0) h = eps;
                        hs[s] = h;
                    }
                }
                __syncthreads();
                // the base mixture and the np perturbed ones, one composed
                // gaussian per thread: (np + 1) * ngauss independent jobs
                // (built one mixture at a time by ngauss threads this was
                // 3 (np + 1) barriers with most of the CTA idle)
                {
                    const int njobs = (np + 1) * a.ngauss;
                    for (int t = threadIdx.x; t < njobs; t += T) {
                        const int s = t / a.ngauss, k = t - s * a.ngauss;
                        double xl[LM_MAXP];
#pragma unroll 1
                        for (int i = 0; i < np; i++) {
                            double xi = sh.xcur[i];
                            if (i == s - 1)
                                xi = b_i2e(a.lm, i, sh.st.x[i] + hs[i]);
                            xl[i] = xi;
                        }
                        build_component(a, &sh, xl, k,
                                        s == 0 ? mix : nullptr, area,
                                        fdm + (size_t)s * a.ngauss * FDREC);
                    }
                }
                __syncthreads();
                const int bad_any = sh.status;
                __syncthreads();
                if (!bad_any) build_tile_masks(a, &sh, mix, tmask, 1.0e-6);
                __syncthreads();
                constexpr int FDN = FD_NSUM_(FDP);
                double sums[FDN];
#pragma unroll
                for (int i = 0; i < FDN; i++) sums[i] = 0.0;
                if (!bad_any)
                    fd_pass<FDP, WG ? 1 : 0>(a, &sh, hs, fdm, etab, tmask, val,
                                                 ierr, wts, nullptr,
                                     rowd, cold, sums, mplane);
                {
                    // block totals, 32 sums at a time
                    constexpr int NCH = (FDN + 31) / 32;
#pragma unroll
                    for (int ch = 0; ch < NCH; ch++) {
                        double s1[32];
#pragma unroll
                        for (int i = 0; i < 32; i++)
                            s1[i] = 32 * ch + i < FDN ? sums[32 * ch + i] : 0.0;
                        block_sum_to<32>(s1, scratch, sh.tot + 32 * ch);
                    }
                }
                if (warp_id == 0) {
                    // feed_jacobian wants {triangle of np, b, r^2} contiguous
                    double *pk = sh.st.T2;
                    const int tri = np * (np + 1) / 2;
                    const int lane = threadIdx.x;
                    for (int i = lane; i < tri; i += 32) pk[i] = sh.tot[i];
                    for (int i = lane; i <= np; i += 32)
                        pk[tri + i] = sh.tot[FD_B0_(FDP) + i];
                    __syncwarp();
                    feed_jacobian(a, &sh, pk, bad_any, 0, hs);
                }
            }
            __syncthreads();
            if (threadIdx.x < a.np) {
                const int i = threadIdx.x;
                sh.xeval[i] = b_i2e(a.lm, i, sh.st.xtrial[i]);
                sh.xcur[i] = b_i2e(a.lm, i, sh.st.x[i]);
            }
            if (threadIdx.x == 0) sh.action = sh.st.action;
            __syncthreads();
            PROF_T(pt_solve);
        }
        PROF_T(pt_solve);
#ifdef NGX_PROF
        // the epilogue is timed with the next stamp's staging
        if (warp_id == 0) lm_finish(a, &sh, &((RES *)a.res)[stamp]);
        if (threadIdx.x == 0) {
            PROF_T(pt_fin);
            double *o = ((RES *)a.res)[stamp].pars_cov0 + 56;
            o[0] = (double)pt_stage; o[1] = (double)pt_mix;
            o[2] = (double)pt_pass; o[3] = (double)pt_solve;
            o[4] = (double)pt_fin;
        }
        __syncthreads();
#else
        pending = stamp;
#endif
        stamp = next;
    }
}

// ---------------------------------------------------------------------
// Objects with several stamps: Fitter.go on an ObsList / MultiBandObsList
// (fitting/results.py:212-347, 439-570).  One CTA per object; the parameters
// are (cen1, cen2, g1, g2, T, F_0 .. F_{nband-1}).  Every pass walks the
// object's stamps: per stamp the band's mixture is composed with that stamp's
// psf and jacobian, the stamp's pixels come straight from global memory / L2
// (nothing is resident: an object may have many epochs), and the per-thread
// sums of the stamp are reduced and scattered into the object's J^T J
// (flux column 5 + band).  Solver, priors, speculation and epilogue are the
// single-stamp ones.
#define MACC_X (LM_MAXP * (LM_MAXP + 1) / 2 + LM_MAXP)
struct LMMulti {
    double macc[MACC_X + 4];    // object sums: packed J^T J (np), J^T f, then
                                // [MACC_X] sum r^2, [+1] s2n numer, [+2]
                                // denom, [+3] r_first^2
    double bp[LM_MAXP];     // band parameters of the current stamp
    double hsl[8];          // forward-difference steps of its parameters
    int first_q_obj, mstatus;
};

__device__ __forceinline__ bool multi_pars_nonzero(const double *x, int nband) {
    bool ok = x[4] != 0.0;
    for (int b = 0; b < nband; b++) ok = ok && x[5 + b] != 0.0;
    return ok;
}

// MODE 0: gauss / exp / dev with the analytic Jacobian.  MODE 1: MINPACK's
// forward differences (bdf, bd, analytic_jacobian = 0): per stamp the base
// mixture and one per local parameter (the shared ones and the band's flux,
// stepped by the global parameter's h) are evaluated in one pass and the
// stamp's (nloc x nloc) block is scattered into the object's J^T J.
template <int MODE, int CV, int CD, typename RES>
__global__ void __launch_bounds__(128, 4)
k_lm_fit_multi(LMArgs a) {
    extern __shared__ double smem_d[];
    __shared__ LMShared sh;
    __shared__ LMMulti mu;
    const int T = blockDim.x;
    const int nwarps = T >> 5;
    const int np = a.np, nband = a.nband;
    const int nsh = np - nband, nloc = nsh + 1;   // shared / per-stamp pars
    const int nfd = MODE == 1 ? nloc + 1 : 0;
    double *mix = smem_d;                                    // ngauss*REC
    double *fdm = mix + (size_t)a.ngauss * REC;      // nfd*ngauss*FDREC
    double *etab = fdm + (size_t)nfd * a.ngauss * FDREC;     // 16 (MODE 1)
    double *rowd = etab + (nfd ? 16 : 0);
    double *cold = rowd + a.nrow;
    double *scratch = cold + a.ncol;                         // 48 * nwarps
    double *hs = scratch + 48 * nwarps;                      // LM_MAXP
    unsigned *tmask = (unsigned *)(hs + LM_MAXP);
    const int tri = np * (np + 1) / 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (nfd) fill_rot_table(etab, 1.0);

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) sh.stamp = atomicAdd(a.counter, 1);
        __syncthreads();
        const int obj = sh.stamp;
        if (obj >= a.n) break;
        if (a.todo && a.todo[obj] == 0) continue;
        const int s0 = a.stamp_start[obj], s1 = a.stamp_start[obj + 1];

        // ---- valid pixels of the object, first valid pixel of its first
        //      stamp (the row run_leastsq drops when a prior is present)
        int cnt = 0, first = a.ntot, firstq = 0;
        for (int e = s0; e < s1; e++) {
            const double *wt = a.weights + (size_t)e * a.ntot;
            for (int q = threadIdx.x; q < a.nq; q += T) {
                int row, col;
                q_to_rowcol(q, a.tpr, &row, &col);
                if (row < a.nrow && col < a.ncol) {
                    int p = row * a.ncol + col;
                    double w = __ldg(&wt[p]);
                    if (!(a.ignore_zero_weight && w <= 0.0)) {
                        cnt++;
                        if (e == s0 && p < first) { first = p; firstq = q; }
                    }
                }
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            int of = __shfl_xor_sync(0xffffffffu, first, o);
            int oq = __shfl_xor_sync(0xffffffffu, firstq, o);
            if (of < first) { first = of; firstq = oq; }
        }
        if (lane == 0) {
            sh.warp_i[warp] = cnt;
            sh.warp_j[warp] = first;
            sh.warp_k[warp] = firstq;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int c = 0, f = a.ntot, fq = -1;
            for (int w = 0; w < nwarps; w++) {
                c += sh.warp_i[w];
                if (sh.warp_j[w] < f) { f = sh.warp_j[w]; fq = sh.warp_k[w]; }
            }
            sh.npix = c;
            mu.first_q_obj = fq;
            lm_init(&sh.st, np, a.guess + (size_t)obj * np, a.lm.maxfev,
                    a.lm.ftol, a.lm.xtol);
            sh.action = sh.st.action;
            sh.spec = 1;
            for (int i = 0; i < np; i++) {
                const double xe = sh.st.xtrial[i];
                sh.xeval[i] = sh.xcur[i] = xe;
                sh.st.x[i] = sh.st.xtrial[i] = b_e2i(a.lm, i, xe);
            }
            sh.cur_ssq = sh.cur_rfirst2 = sh.cur_num = sh.cur_den = 0.0;
        }
        __syncthreads();
        if (sh.npix == 0 || s1 <= s0) {
            if (threadIdx.x == 0) {
                RES *r = &((RES *)a.res)[obj];
                memset(r, 0, sizeof(*r));
                r->flags = NGX_NO_ATTEMPT;
                for (int i = 0; i < np; i++) {
                    r->pars[i] = NGX_PDEF;
                    r->pars_err[i] = NGX_CDEF;
                }
            }
            continue;
        }

        while (sh.action != LM_ACT_DONE) {
            const int action = sh.action;
            const int spec = sh.spec;
            __syncthreads();
            const bool want_f = (action == LM_ACT_EVAL_F);
            const double *xp = want_f ? sh.xeval : sh.xcur;
            const bool nonzero = MODE == 1 || multi_pars_nonzero(xp, nband);
            const bool fused = MODE == 0 && want_f && spec && nonzero
                && !(sh.st.prered <= sh.st.ftol);
            const bool do_deriv = fused || action == LM_ACT_EVAL_J;
            for (int i = threadIdx.x; i < MACC_X + 4; i += blockDim.x)
                mu.macc[i] = 0.0;
            if (threadIdx.x == 0) {
                mu.mstatus = 0;
                if (MODE == 1 && do_deriv) {
                    // fdjac2's steps on the (internal) parameters
                    const double eps = 1.4901161193847656e-08;
                    for (int i = 0; i < np; i++) {
                        double h = eps * fabs(sh.st.x[i]);
                        if (h == 0.0) h = eps;
                        hs[i] = h;
                    }
                }
            }
            __syncthreads();
            // calc_jacobian raises for T == 0 or a zero flux
            const int jstatus0 = do_deriv && !nonzero;
            for (int e = s0; e < s1; e++) {
                const int band = a.stamp_band[e];
                if (threadIdx.x == 0) {
                    sh.jac = a.jac[e];
                    sh.status = 0;
                    for (int i = 0; i < nsh; i++) mu.bp[i] = xp[i];
                    mu.bp[nsh] = xp[nsh + band];
                    if (MODE == 1)
                        for (int l = 0; l < nloc; l++)
                            mu.hsl[l] = hs[l < nsh ? l : nsh + band];
                    sh.first_q = (e == s0) ? mu.first_q_obj : -1;
                }
                if (threadIdx.x < 6 * a.npsf)
                    sh.psf[threadIdx.x] = a.psf[(size_t)e * 6 * a.npsf
                                                + threadIdx.x];
                __syncthreads();
                for (int p = threadIdx.x; p < a.nrow; p += T)
                    rowd[p] = p - sh.jac.row0;
                for (int p = threadIdx.x; p < a.ncol; p += T)
                    cold[p] = p - sh.jac.col0;
                if (threadIdx.x == 0 && a.npsf > 0) {
                    // psf centroid and flux, gmix_get_cen :108-130
                    double prow = 0.0, pcol = 0.0, psum = 0.0;
                    for (int k = 0; k < a.npsf; k++) {
                        double p = sh.psf[6 * k];
                        prow += p * sh.psf[6 * k + 1];
                        pcol += p * sh.psf[6 * k + 2];
                        psum += p;
                    }
                    sh.psf_row = prow / psum;
                    sh.psf_col = pcol / psum;
                    sh.psf_ipsum = 1.0 / psum;
                }
                __syncthreads();
                const double area = sh.jac.scale * sh.jac.scale;
                if (MODE == 1 && do_deriv) {
                    // base + nloc perturbed mixtures, a gaussian per thread
                    const int njobs = (nloc + 1) * a.ngauss;
                    for (int t = threadIdx.x; t < njobs; t += T) {
                        const int s = t / a.ngauss, k = t - s * a.ngauss;
                        double xl[LM_MAXP];
#pragma unroll 1
                        for (int i = 0; i < nloc; i++) {
                            double xi = mu.bp[i];
                            if (i == s - 1) {
                                const int gi = i < nsh ? i : nsh + band;
                                xi = b_i2e(a.lm, gi, sh.st.x[gi] + hs[gi]);
                            }
                            xl[i] = xi;
                        }
                        build_component(a, &sh, xl, k,
                                        s == 0 ? mix : nullptr, area,
                                        fdm + (size_t)s * a.ngauss * FDREC);
                    }
                } else {
                    build_mixture(a, &sh, mu.bp, mix, area);
                }
                __syncthreads();
                if (sh.status) {            // uniform: read after the barrier
                    if (threadIdx.x == 0) mu.mstatus = 1;
                    break;
                }
                build_tile_masks(a, &sh, mix, tmask,
                                 MODE == 1 && do_deriv ? 1.0e-6 : 1.0e-9);
                __syncthreads();
                const double *im = a.images + (size_t)e * a.ntot;
                const double *wt = a.weights + (size_t)e * a.ntot;
                if (MODE == 1 && do_deriv) {
                    constexpr int FDN = FD_NSUM_(FD_MAXP);
                    double sums[FDN];
#pragma unroll
                    for (int i = 0; i < FDN; i++) sums[i] = 0.0;
                    fd_pass<FD_MAXP, 2>(a, &sh, mu.hsl, fdm, etab, tmask,
                                        nullptr, nullptr, wt, im, rowd, cold,
                                        sums, nullptr, nloc);
                    {
                        constexpr int NCH = (FDN + 31) / 32;
#pragma unroll
                        for (int ch = 0; ch < NCH; ch++) {
                            double s1[32];
#pragma unroll
                            for (int i = 0; i < 32; i++)
                                s1[i] = 32 * ch + i < FDN ? sums[32 * ch + i]
                                                          : 0.0;
                            block_sum_to<32>(s1, scratch, sh.tot + 32 * ch);
                        }
                    }
                    if (threadIdx.x == 0) {
                        // scatter the stamp's nloc x nloc block: local
                        // parameter l < nsh is global l, local nsh is the
                        // band's flux
                        const double *t = sh.tot;
                        for (int la = 0; la < nloc; la++) {
                            int ga = la < nsh ? la : nsh + band;
                            for (int lb = 0; lb <= la; lb++) {
                                int gb = lb < nsh ? lb : nsh + band;
                                mu.macc[ga * (ga + 1) / 2 + gb]
                                    += t[la * (la + 1) / 2 + lb];
                            }
                            mu.macc[tri + ga] += t[FD_B0_(FD_MAXP) + la];
                        }
                    }
                } else if (do_deriv) {
                    double sums[NDSUM];
#pragma unroll
                    for (int i = 0; i < NDSUM; i++) sums[i] = 0.0;
                    if (!jstatus0)
                        deriv_pass<CD, 2>(a, &sh, mu.bp, mix, tmask, nullptr,
                                          nullptr, wt, im, rowd, cold, sums);
                    block_sum_to<NDSUM>(sums, scratch, sh.tot);
                    if (threadIdx.x == 0) {
                        // scatter the stamp's 6 x 6 block: local parameter
                        // l < 5 is global l, local 5 is the band's flux
                        const double *t = sh.tot;
                        int q = 0;
                        for (int la = 0; la < 6; la++) {
                            int ga = la < 5 ? la : 5 + band;
                            for (int lb = 0; lb <= la; lb++) {
                                int gb = lb < 5 ? lb : 5 + band;
                                mu.macc[ga * (ga + 1) / 2 + gb] += t[q];
                                q++;
                            }
                            mu.macc[tri + ga] += t[21 + la];
                        }
                        mu.macc[MACC_X] += t[27];
                        mu.macc[MACC_X + 1] += t[28];
                        mu.macc[MACC_X + 2] += t[29];
                        if (e == s0) mu.macc[MACC_X + 3] = t[30];
                    }
                } else {
                    double sums[4] = {0, 0, 0, 0};
                    value_pass<CV, 2>(a, &sh, mix, tmask, nullptr, nullptr, wt,
                                      im, rowd, cold, sums);
                    block_sum_to<4>(sums, scratch, sh.tot);
                    if (threadIdx.x == 0) {
                        mu.macc[MACC_X] += sh.tot[0];
                        mu.macc[MACC_X + 1] += sh.tot[1];
                        mu.macc[MACC_X + 2] += sh.tot[2];
                        if (e == s0) mu.macc[MACC_X + 3] = sh.tot[3];
                    }
                }
                __syncthreads();
            }
            __syncthreads();
            if (warp == 0) {
                const int mstatus = mu.mstatus;
                const int jstatus = mstatus || jstatus0;
                if (want_f) {
                    int acc = feed_residual(a, &sh, xp, mu.macc[MACC_X],
                                                mstatus);
                    if (acc) {
                        if (lane == 0) {
                            sh.cur_ssq = mu.macc[MACC_X];
                            sh.cur_num = mu.macc[MACC_X + 1];
                            sh.cur_den = mu.macc[MACC_X + 2];
                            sh.cur_rfirst2 = mu.macc[MACC_X + 3];
                        }
                        if (fused && sh.st.action == LM_ACT_EVAL_J)
                            feed_jacobian(a, &sh, mu.macc, jstatus, 1, hs);
                    }
                    if (lane == 0) sh.spec = acc;
                } else {
                    feed_jacobian(a, &sh, mu.macc, jstatus, MODE == 0, hs);
                }
                __syncwarp();
                if (lane < np) {
                    sh.xeval[lane] = b_i2e(a.lm, lane, sh.st.xtrial[lane]);
                    sh.xcur[lane] = b_i2e(a.lm, lane, sh.st.x[lane]);
                }
                if (lane == 0) sh.action = sh.st.action;
            }
            __syncthreads();
        }
        if (warp == 0) lm_finish(a, &sh, &((RES *)a.res)[obj]);
    }
}

// ---------------------------------------------------------------------
// Work-queue counter of one launch.  Each device has a ring of counters
// (allocated once); a launch takes the next slot whose previous user has
// finished - known from an event recorded behind that launch on its stream -
// zeroes it on its own stream and records the slot's event behind its kernel.
// Launches on different streams, devices or host threads therefore never
// share a live counter, no launch waits for another stream, and nothing is
// allocated per call.  (Stream-ordered allocations were tried for this:
// with the default pool a counter freed on one stream and reused on another
// makes the second stream wait for the first one's kernel, which serialises
// the two compute streams of the host pipeline; a private pool without
// cross-stream reuse maps fresh memory on nearly every call.)
#define NGX_RING 256
struct CounterRing {
    int *base = nullptr;
    cudaEvent_t ev[NGX_RING];
    bool used[NGX_RING];
    unsigned next = 0;
};
static std::mutex g_ring_mutex;
static CounterRing g_rings[64];

static int get_counter(int **out, int *slot_out) {
    int dev = 0;
    NGX_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64)
        return fail(NGX_ERR_ARG, "ngx_lm_fit: device index out of range%s%s");
    std::lock_guard<std::mutex> lock(g_ring_mutex);
    CounterRing &r = g_rings[dev];
    if (!r.base) {
        NGX_CUDA(cudaMalloc((void **)&r.base, NGX_RING * 32 * sizeof(int)));
        for (int i = 0; i < NGX_RING; i++) {
            NGX_CUDA(cudaEventCreateWithFlags(&r.ev[i],
                                              cudaEventDisableTiming));
            r.used[i] = false;
        }
    }
    int slot = -1;
    for (int tries = 0; tries < NGX_RING; tries++) {
        int i = (int)(r.next++ % NGX_RING);
        if (!r.used[i] || cudaEventQuery(r.ev[i]) == cudaSuccess) {
            slot = i;
            break;
        }
    }
    cudaGetLastError();             // cudaErrorNotReady from the queries
    if (slot < 0) {
        // every counter is still in flight: wait for the oldest
        slot = (int)(r.next++ % NGX_RING);
        NGX_CUDA(cudaEventSynchronize(r.ev[slot]));
    }
    r.used[slot] = true;
    *out = r.base + 32 * slot;
    *slot_out = slot;
    return NGX_OK;
}

// zero the counter ahead of the kernel and mark the slot busy until the
// work queued on `st` so far (the caller records after its launch) is done
static int counter_begin(int **out, int *slot, cudaStream_t st) {
    int rc = get_counter(out, slot);
    if (rc) return rc;
    NGX_CUDA(cudaMemsetAsync(*out, 0, sizeof(int), st));
    return NGX_OK;
}
static void counter_end(int slot, cudaStream_t st) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    // the slot is ours until this record: no lock needed for ev[slot]
    cudaEventRecord(g_rings[dev].ev[slot], st);
}

struct LMLaunch {
    int threads, blocks;
    size_t smem;
};

#ifndef NGX_LM_WIDE_TU
// float -> double, four elements per thread (the f32 host entry point)
__global__ void k_widen(const float *__restrict__ in, double *__restrict__ out,
                        size_t n) {
    size_t i = 4 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (i + 3 < n && (((uintptr_t)(in + i)) & 15) == 0) {
        float4 v = *(const float4 *)(in + i);
        out[i] = v.x; out[i + 1] = v.y; out[i + 2] = v.z; out[i + 3] = v.w;
    } else {
        for (size_t k = i; k < n && k < i + 4; k++) out[k] = in[k];
    }
}
#endif

#ifndef NGX_LM_WIDE_TU
// compact host format -> the engine's planes: float32 pixel values widened,
// weight = the stamp's scalar weight where the stamp's mask bit is set (bit
// k of the stamp, most significant bit of each byte first, as numpy.packbits
// writes it; no mask = all pixels good), 0 elsewhere.  One thread per pixel
// pair; 4.1 bytes per pixel came over PCIe instead of 16.
__global__ void k_expand_compact(const float *__restrict__ im32,
                                 const double *__restrict__ wscalar,
                                 const unsigned char *__restrict__ mask,
                                 double *__restrict__ im,
                                 double *__restrict__ wt, int n, int ntot,
                                 int nmb) {
    const size_t total = (size_t)n * ntot;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int s = (int)(i / ntot), k = (int)(i - (size_t)s * ntot);
        bool good = true;
        if (mask) good = (mask[(size_t)s * nmb + (k >> 3)] >> (7 - (k & 7))) & 1;
        im[i] = (double)im32[i];
        wt[i] = good ? wscalar[s] : 0.0;
    }
}
#endif

#ifndef NGX_LM_WIDE_TU
// stamp_start = 0..n and stamp_band = 0 for n one-stamp objects
__global__ void k_iota_zero(int n, int32_t *start, int32_t *band) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) start[i] = i;
    if (i < n) band[i] = 0;
}
#endif

// Balance of the static tile assignment.  In every round of a pass warp w
// takes chunk slot (w + rot * round) mod nwarps (C consecutive tiles).  With
// rot = 0 a warp can sit on the same tile columns in every round - on a 64x64
// stamp (8 tiles per row, 8 warps) four warps would always get the centre
// columns, where every gaussian of the mixture is in reach, and the others
// the edges: 1.5x the mean work on the slowest warp, the rest waiting at the
// reduction barrier.  The stride is chosen per (stamp size, warps, C) on a
// proxy cost: gaussians of geometrically spaced radii centred on the stamp,
// a tile costing one unit per gaussian whose 5-sigma box reaches it.
static int pick_rotation(int nrow, int ncol, int nwarps, int C) {
    struct Key { int nrow, ncol, nwarps, C, rot; };
    static thread_local Key cache[8];
    static thread_local int ncache = 0;
    for (int i = 0; i < ncache; i++)
        if (cache[i].nrow == nrow && cache[i].ncol == ncol
            && cache[i].nwarps == nwarps && cache[i].C == C)
            return cache[i].rot;
    const int tr = (nrow + 3) / 4, tc = (ncol + 7) / 8, ntile = tr * tc;
    std::vector<float> cost(ntile, 0.0f);
    const float cr = 0.5f * (nrow - 1), cc = 0.5f * (ncol - 1);
    for (int k = 0; k < 14; k++) {
        const float rad = 0.5f * powf(2.0f, 0.5f * k);
        for (int i = 0; i < tr; i++)
            for (int j = 0; j < tc; j++) {
                float r0 = 4.0f * i, r1 = fminf(r0 + 3, nrow - 1.0f);
                float c0 = 8.0f * j, c1 = fminf(c0 + 7, ncol - 1.0f);
                float dr = fmaxf(0.0f, fmaxf(r0 - cr, cr - r1));
                float dc = fmaxf(0.0f, fmaxf(c0 - cc, cc - c1));
                if (dr <= rad && dc <= rad) cost[i * tc + j] += 1.0f;
            }
    }
    int best = 0;
    float best_max = 1e30f;
    for (int rot = 0; rot < nwarps; rot++) {
        float tot[32] = {0};
        int rnd = 0;
        for (int base = 0; base < ntile; base += nwarps * C, rnd++)
            for (int w = 0; w < nwarps; w++) {
                int ws = (w + rot * rnd) % nwarps;
                for (int c = 0; c < C; c++) {
                    int t = base + ws * C + c;
                    if (t < ntile) tot[w] += cost[t];
                }
            }
        float mx = 0.0f;
        for (int w = 0; w < nwarps; w++) mx = fmaxf(mx, tot[w]);
        if (mx < best_max - 1e-3f) { best_max = mx; best = rot; }
    }
    Key k = {nrow, ncol, nwarps, C, best};
    cache[ncache < 8 ? ncache++ : 7] = k;
    return best;
}

template <typename K>
static int config_and_launch(K kernel, const LMArgs &a_in, int threads,
                             size_t smem, cudaStream_t stream) {
    const LMArgs &a0 = a_in;
    if (smem > (size_t)smem_optin_bytes())
        return fail(NGX_ERR_ARG,
                    "ngx_lm_fit: stamp + mixture do not fit shared memory%s%s");
    NGX_CUDA(cudaFuncSetAttribute(kernel,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    int per_sm = 0;
    NGX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel,
                                                           threads, smem));
    if (per_sm < 1) per_sm = 1;
    int blocks = sm_count() * per_sm;
    if (blocks > a0.n) blocks = a0.n;
    LMArgs a = a_in;
    // CV = 3 tiles per warp and round in the value pass, CD = CF = 2 in the
    // derivative / forward-difference passes (template arguments below)
    a.rot_v = pick_rotation(a.nrow, a.ncol, threads / 32, 3);
    a.rot_d = pick_rotation(a.nrow, a.ncol, threads / 32, 2);
    int slot = 0;
    int rcc = counter_begin(&a.counter, &slot, stream);
    if (rcc) return rcc;
    kernel<<<blocks, threads, smem, stream>>>(a);
    cudaError_t e = cudaGetLastError();
    counter_end(slot, stream);
    NGX_CUDA(e);
    return NGX_OK;
}

}  // namespace ngx

using namespace ngx;

extern "C" {

#ifndef NGX_LM_WIDE_TU
// shared memory of k_lm_fit without the ierr plane (*smem); returns whether
// the weights stay in global memory: ierr = sqrt(w) is kept in shared memory
// too when four CTAs (the register limit) still fit on an SM, larger stamps
// leave the weights in L2
static bool lm_weights_global(int ngauss, int nfd, int nq, int nrow, int ncol,
                              int nwarps, size_t *smem) {
    *smem = sizeof(double)
        * ((size_t)ngauss * REC + (size_t)nfd * ngauss * FDREC
           + (nfd ? 16 + (size_t)nq : 0) + (size_t)nq + nrow + ncol
           + 48 * nwarps + LM_MAXP
           + ((size_t)(nq / 32) * ((ngauss + 31) / 32) + 1) / 2);
    size_t per_cta = *smem + sizeof(double) * (size_t)nq + sizeof(LMShared)
        + 1024;
    return 4 * per_cta > (size_t)smem_optin_bytes() + 1024;
}
#endif

// objects to fit in the current call (ngx_lm_fit_todo); NULL = all
static thread_local const int32_t *g_lm_todo = nullptr;

// streamed input of the current call (lm_fit_host_stream); NULL = resident
struct LMFeed {
    const int *ready;
    int ready_chunk, src, nmb, done_chunk;
    const float *im32, *wt32;
    const double *wscalar;
    const unsigned char *maskbits;
    int *done_count, *host_done, *err, *err_dev;
};
static thread_local const LMFeed *g_lm_feed = nullptr;

#ifndef NGX_LM_WIDE_TU
int ngx_lm_fit_todo(int n, int model, const double *images,
                    const double *weights, const ngx_jacobian *jac, int nrow,
                    int ncol, const double *psf_pars, int npsf,
                    const double *guess, const ngx_prior *prior,
                    const ngx_lm_pars *lm_pars, int ignore_zero_weight,
                    const int32_t *todo, ngx_lm_result *res, void *stream) {
    g_lm_todo = todo;
    int rc = ngx_lm_fit(n, model, images, weights, jac, nrow, ncol, psf_pars,
                        npsf, guess, prior, lm_pars, ignore_zero_weight, res,
                        stream);
    g_lm_todo = nullptr;
    return rc;
}

int ngx_lm_fit(int n, int model, const double *images, const double *weights,
               const ngx_jacobian *jac, int nrow, int ncol,
               const double *psf_pars, int npsf, const double *guess,
               const ngx_prior *prior, const ngx_lm_pars *lm_pars,
               int ignore_zero_weight, ngx_lm_result *res, void *stream) {
    NGX_REQUIRE(n >= 0 && images && weights && jac && guess && res);
    NGX_REQUIRE(nrow >= 1 && ncol >= 1 && npsf >= 0 && npsf <= 16);
    NGX_REQUIRE(npsf == 0 || psf_pars);
    NGX_REQUIRE(lm_pars && lm_pars->maxfev > 0);
    int np = model_npars(model), ng0 = model_ngauss(model);
    NGX_REQUIRE(np > 0 && ng0 > 0 && model != NGX_CM);
    if (n == 0) return NGX_OK;

    LMArgs a;
    memset(&a, 0, sizeof(a));
    a.n = n; a.model = model; a.np = np; a.ngauss0 = ng0; a.npsf = npsf;
    a.ngauss = ng0 * (npsf > 0 ? npsf : 1);
    a.nrow = nrow; a.ncol = ncol; a.ntot = nrow * ncol;
    a.ignore_zero_weight = ignore_zero_weight;
    if (prior) a.prior = *prior;
    a.lm = *lm_pars;
    a.images = images; a.weights = weights; a.jac = jac;
    a.psf = psf_pars; a.guess = guess; a.res = res;
    a.todo = g_lm_todo;
    const LMFeed *feed = g_lm_feed;
    if (feed) {
        a.ready = feed->ready; a.ready_chunk = feed->ready_chunk;
        a.src = feed->src; a.nmb = feed->nmb; a.done_chunk = feed->done_chunk;
        a.im32 = feed->im32; a.wt32 = feed->wt32; a.wscalar = feed->wscalar;
        a.maskbits = feed->maskbits; a.done_count = feed->done_count;
        a.host_done = feed->host_done; a.err = feed->err;
        a.err_dev = feed->err_dev;
    }
    NGX_REQUIRE(a.ngauss <= NGX_MAX_GAUSS);
    if (a.prior.present) {
        NGX_REQUIRE(a.prior.g_kind == 1);
    }

    bool analytic = lm_pars->analytic_jacobian
        && (model == NGX_GAUSS || model == NGX_EXP || model == NGX_DEV);
    a.analytic = analytic;
    int threads = 128;
    if (a.ntot <= 1024) threads = 64;
    NGX_REQUIRE(threads >= a.ngauss && threads >= 6 * npsf);
    // tiled pixel order: 4 x 8 tiles, one per warp
    a.tpr = (ncol + 7) / 8;
    a.nq = 32 * a.tpr * ((nrow + 3) / 4);
    int nfd = analytic ? 0 : np + 1;
    int nwarps = threads / 32;
    size_t smem = 0;
    bool wg = lm_weights_global(a.ngauss, nfd, a.nq, nrow, ncol, nwarps, &smem);
    size_t smem_both = smem + sizeof(double) * (size_t)a.nq;
    cudaStream_t st = as_stream(stream);
    // a streamed feed is staged by the kernels that keep both planes resident
    NGX_REQUIRE(!feed || !wg);
    if (analytic && smem + sizeof(LMShared) + 1024
            > (size_t)smem_optin_bytes()) {
        // the stamp does not fit shared memory even without the weights:
        // fit each stamp as a one-epoch object of the multi-epoch kernel,
        // which reads its pixels from global memory on every pass
        int32_t *ss = nullptr;
        NGX_CUDA(cudaMallocAsync((void **)&ss, sizeof(int32_t) * (2 * n + 1),
                                 st));
        k_iota_zero<<<(n + 256) / 256, 256, 0, st>>>(n, ss, ss + n + 1);
        int rc2 = ngx_lm_fit_multi(n, model, ss, ss + n + 1, 1, images,
                                   weights, jac, nrow, ncol, psf_pars, npsf,
                                   guess, prior, lm_pars, ignore_zero_weight,
                                   res, stream);
        cudaFreeAsync(ss, st);
        return rc2;
    }
    if (analytic) {
        if (wg)
            return config_and_launch(k_lm_fit<0, 3, 2, 4, true>, a, threads,
                                     smem, st);
        return config_and_launch(k_lm_fit<0, 3, 2, 4, false>, a, threads,
                                 smem_both, st);
    }
    if (smem + sizeof(LMShared) + 1024 > (size_t)smem_optin_bytes()) {
        // forward-difference fit of a stamp that does not fit shared memory:
        // the value and base-model planes go to a per-CTA global scratch
        size_t smem_g = smem - sizeof(double) * (size_t)a.nq * (nfd ? 2 : 1)
            + sizeof(double) * 48 * 4;
        if (smem_g > (size_t)smem_optin_bytes())
            return fail(NGX_ERR_ARG,
                        "ngx_lm_fit: mixture does not fit shared memory%s%s");
        auto kern = k_lm_fit<1, 3, 2, 2, true, 256, true>;
        NGX_CUDA(cudaFuncSetAttribute(
            kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
        int per_sm = 0;
        NGX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &per_sm, kern, 256, smem_g));
        if (per_sm < 1) per_sm = 1;
        int blocks = sm_count() * per_sm;
        if (blocks > n) blocks = n;
        NGX_CUDA(cudaMallocAsync(
            (void **)&a.gscratch,
            sizeof(double) * (size_t)blocks * (nfd ? 2 : 1) * a.nq, st));
        int slot = 0;
        int rcc = counter_begin(&a.counter, &slot, st);
        if (rcc) {
            cudaFreeAsync(a.gscratch, st);
            return rcc;
        }
        kern<<<blocks, 256, smem_g, st>>>(a);
        cudaError_t e = cudaGetLastError();
        counter_end(slot, st);
        cudaFreeAsync(a.gscratch, st);
        NGX_CUDA(e);
        return NGX_OK;
    }
    if (wg) {
        // large stamps with the np + 1 forward-difference mixtures resident
        // leave room for few CTAs: give each one eight warps
        if (threads == 128 && 4 * (smem + sizeof(LMShared) + 1024)
                > (size_t)smem_optin_bytes() + 1024) {
            size_t smem8 = smem + sizeof(double) * 48 * 4;
            return config_and_launch(k_lm_fit<1, 3, 2, 2, true, 256>, a, 256,
                                     smem8, st);
        }
        return config_and_launch(k_lm_fit<1, 3, 2, 4, true>, a, threads, smem,
                                 st);
    }
    return config_and_launch(k_lm_fit<1, 3, 2, 4, false>, a, threads,
                             smem_both, st);
}
#endif

static int lm_fit_multi_impl(int nobj, int model, const int32_t *stamp_start,
                             const int32_t *stamp_band, int nband,
                             const double *images, const double *weights,
                             const ngx_jacobian *jac, int nrow, int ncol,
                             const double *psf_pars, int npsf,
                             const double *guess, const ngx_prior *prior,
                             const ngx_lm_pars *lm_pars,
                             int ignore_zero_weight, void *res, int wide,
                             void *stream) {
    NGX_REQUIRE(nobj >= 0 && stamp_start && stamp_band && images && weights);
    NGX_REQUIRE(jac && guess && res && lm_pars && lm_pars->maxfev > 0);
    NGX_REQUIRE(nrow >= 1 && ncol >= 1 && npsf >= 0 && npsf <= 16);
    NGX_REQUIRE(npsf == 0 || psf_pars);
    const bool simple = model == NGX_GAUSS || model == NGX_EXP
        || model == NGX_DEV;
    NGX_REQUIRE(simple || model == NGX_TURB || model == NGX_BDF
                || model == NGX_BD);
    // shared parameters: 5 (simple), 6 (bdf), 7 (bd), then a flux per band
    const int nsh = model_npars(model) - 1;
    NGX_REQUIRE(nband >= 1 && nband <= (LM_MAXP >= 20 ? NGX_MAX_BANDS : 7)
                && nsh + nband <= LM_MAXP);
    // forward differences (Fitter.go, fitters.py:93-98): always for the
    // models without analytic derivatives
    const bool analytic = simple && lm_pars->analytic_jacobian;
    if (nobj == 0) return NGX_OK;
    LMArgs a;
    memset(&a, 0, sizeof(a));
    a.n = nobj; a.model = model; a.np = nsh + nband; a.nband = nband;
    a.ngauss0 = model_ngauss(model); a.npsf = npsf;
    a.ngauss = a.ngauss0 * (npsf > 0 ? npsf : 1);
    a.nrow = nrow; a.ncol = ncol; a.ntot = nrow * ncol;
    a.ignore_zero_weight = ignore_zero_weight;
    a.analytic = analytic;
    if (prior) a.prior = *prior;
    if (a.prior.present) {
        NGX_REQUIRE(a.prior.g_kind == 1);
        a.prior.nband = nband;
    }
    a.lm = *lm_pars;
    a.images = images; a.weights = weights; a.jac = jac;
    a.psf = psf_pars; a.guess = guess; a.res = res;
    a.stamp_start = stamp_start; a.stamp_band = stamp_band;
    a.todo = g_lm_todo;
    NGX_REQUIRE(a.ngauss <= NGX_MAX_GAUSS);
    int threads = 128;
    NGX_REQUIRE(threads >= a.ngauss && threads >= 6 * npsf);
    a.tpr = (ncol + 7) / 8;
    a.nq = 32 * a.tpr * ((nrow + 3) / 4);
    int nwarps = threads / 32;
    const int nfd = analytic ? 0 : nsh + 2;     // base + nloc mixtures
    size_t smem = sizeof(double)
        * ((size_t)a.ngauss * REC + (size_t)nfd * a.ngauss * FDREC
           + (nfd ? 16 : 0) + nrow + ncol + 48 * nwarps + LM_MAXP
           + ((size_t)(a.nq / 32) * ((a.ngauss + 31) / 32) + 1) / 2);
    // the 8-parameter and the 14-parameter solver workspaces are compiled in
    // separate translation units (ngx_lm.cu / ngx_lm_wide.cu)
#ifdef NGX_LM_WIDE_TU
    NGX_REQUIRE(wide);
    if (!analytic)
        return config_and_launch(k_lm_fit_multi<1, 3, 2, ngx_lm_result_wide>,
                                 a, threads, smem, as_stream(stream));
    return config_and_launch(k_lm_fit_multi<0, 3, 2, ngx_lm_result_wide>, a,
                             threads, smem, as_stream(stream));
#else
    NGX_REQUIRE(!wide);
    if (!analytic)
        return config_and_launch(k_lm_fit_multi<1, 3, 2, ngx_lm_result>, a,
                                 threads, smem, as_stream(stream));
    return config_and_launch(k_lm_fit_multi<0, 3, 2, ngx_lm_result>, a,
                             threads, smem, as_stream(stream));
#endif
}

#if defined(NGX_LM_WIDE_TU) && !defined(NGX_LM_XWIDE_TU)
// CoellipFitter.go (fitting/fitters.py:120-141 -> CoellipFitModel,
// results.py:628-674): n coelliptical gaussians, parameters (cen1, cen2, g1,
// g2, T_1..T_n, F_1..F_n), forward-difference Jacobian (MINPACK lmdif),
// optional PriorCoellipSame (prior->ncoellip = ngauss).  ngauss <= 5 (14
// parameters); the 256-thread kernel with the weights in L2.
int ngx_lm_fit_coellip(int n, int ngauss, const double *images,
                       const double *weights, const ngx_jacobian *jac,
                       int nrow, int ncol, const double *psf_pars, int npsf,
                       const double *guess, const ngx_prior *prior,
                       const ngx_lm_pars *lm_pars, int ignore_zero_weight,
                       ngx_lm_result_wide *res, void *stream) {
    NGX_REQUIRE(n >= 0 && images && weights && jac && guess && res);
    NGX_REQUIRE(nrow >= 1 && ncol >= 1 && npsf >= 0 && npsf <= 16);
    NGX_REQUIRE(npsf == 0 || psf_pars);
    NGX_REQUIRE(lm_pars && lm_pars->maxfev > 0);
    NGX_REQUIRE(ngauss >= 1 && 4 + 2 * ngauss <= LM_MAXP);
    if (n == 0) return NGX_OK;
    LMArgs a;
    memset(&a, 0, sizeof(a));
    a.n = n; a.model = NGX_COELLIP; a.np = 4 + 2 * ngauss;
    a.ngauss0 = ngauss; a.npsf = npsf;
    a.ngauss = ngauss * (npsf > 0 ? npsf : 1);
    a.nrow = nrow; a.ncol = ncol; a.ntot = nrow * ncol;
    a.ignore_zero_weight = ignore_zero_weight;
    if (prior) a.prior = *prior;
    if (a.prior.present) {
        NGX_REQUIRE(a.prior.g_kind == 1);
        a.prior.ncoellip = ngauss;
    }
    a.lm = *lm_pars;
    a.images = images; a.weights = weights; a.jac = jac;
    a.psf = psf_pars; a.guess = guess; a.res = res;
    a.analytic = 0;
    NGX_REQUIRE(a.ngauss <= NGX_MAX_GAUSS);
    const int threads = 256, nwarps = 8;
    NGX_REQUIRE(threads >= a.ngauss && threads >= 6 * npsf);
    a.tpr = (ncol + 7) / 8;
    a.nq = 32 * a.tpr * ((nrow + 3) / 4);
    const int nfd = a.np + 1;
    size_t smem = sizeof(double)
        * ((size_t)a.ngauss * REC + (size_t)nfd * a.ngauss * FDREC
           + 16 + 2 * (size_t)a.nq + nrow + ncol + 48 * nwarps + LM_MAXP
           + ((size_t)(a.nq / 32) * ((a.ngauss + 31) / 32) + 1) / 2);
    // the per-thread J^T J sums are sized for the fit's own parameter
    // count (45 for 8 parameters ... 120 for 14): two gaussians run without
    // the register spills of the 14-parameter instantiation
    cudaStream_t st = as_stream(stream);
    if (a.ntot <= 1024 && a.ngauss <= 64 && 6 * npsf <= 64
        && !getenv("NGX_COELLIP_256")) {
        // psf-sized stamps (the usual input of CoellipFitter): 28 tiles keep
        // two warps busy, not eight (at 256 threads 6.5 of every 7.5 warp
        // cycles waited at a barrier, profiles/ncu_kernels_r2b.txt); both
        // planes resident, eight CTAs per SM
        const int t64 = 64;
        size_t smem64 = smem - sizeof(double) * 48 * (nwarps - 2)
            + sizeof(double) * (size_t)a.nq;
        if (a.np <= 8)
            return config_and_launch(
                k_lm_fit<1, 3, 2, 8, false, 64, false, 8, ngx_lm_result_wide>,
                a, t64, smem64, st);
        if (a.np <= 10)
            return config_and_launch(
                k_lm_fit<1, 3, 2, 8, false, 64, false, 10,
                         ngx_lm_result_wide>,
                a, t64, smem64, st);
        if (a.np <= 12)
            return config_and_launch(
                k_lm_fit<1, 3, 2, 8, false, 64, false, 12,
                         ngx_lm_result_wide>,
                a, t64, smem64, st);
        return config_and_launch(
            k_lm_fit<1, 3, 2, 8, false, 64, false, LM_MAXP,
                     ngx_lm_result_wide>,
            a, t64, smem64, st);
    }
    if (a.np <= 8)
        return config_and_launch(
            k_lm_fit<1, 3, 2, 2, true, 256, false, 8, ngx_lm_result_wide>,
            a, threads, smem, st);
    if (a.np <= 10)
        return config_and_launch(
            k_lm_fit<1, 3, 2, 2, true, 256, false, 10, ngx_lm_result_wide>,
            a, threads, smem, st);
    if (a.np <= 12)
        return config_and_launch(
            k_lm_fit<1, 3, 2, 2, true, 256, false, 12, ngx_lm_result_wide>,
            a, threads, smem, st);
    return config_and_launch(
        k_lm_fit<1, 3, 2, 2, true, 256, false, LM_MAXP, ngx_lm_result_wide>,
        a, threads, smem, st);
}
#endif

#if defined(NGX_LM_WIDE_TU) && !defined(NGX_LM_XWIDE_TU)
int ngx_lm_fit_coellip_host(int n, int ngauss, const double *images,
                            const double *weights, const ngx_jacobian *jac,
                            int nrow, int ncol, const double *psf_pars,
                            int npsf, const double *guess,
                            const ngx_prior *prior,
                            const ngx_lm_pars *lm_pars,
                            int ignore_zero_weight,
                            ngx_lm_result_wide *res) {
    NGX_REQUIRE(n >= 0 && images && weights && jac && guess && res);
    NGX_REQUIRE(ngauss >= 1 && 4 + 2 * ngauss <= LM_MAXP);
    if (n == 0) return NGX_OK;
    DevBuf<ngx_jacobian> dj(n, jac);
    DevBuf<double> dp(npsf > 0 ? (size_t)n * npsf * 6 : 0, psf_pars);
    DevBuf<double> dg((size_t)n * (4 + 2 * ngauss), guess);
    DevBuf<ngx_lm_result_wide> dr(n);
    NGX_CUDA(dj.err); NGX_CUDA(dp.err); NGX_CUDA(dg.err); NGX_CUDA(dr.err);
    const int npg = 4 + 2 * ngauss;
    int rc = planes_pipeline(
        n, (size_t)nrow * ncol, images, weights,
        [&](int lo, int cn, const double *di, const double *dw,
            cudaStream_t q) {
            return ngx_lm_fit_coellip(
                cn, ngauss, di, dw, dj.d + lo, nrow, ncol,
                npsf > 0 ? dp.d + (size_t)lo * npsf * 6 : nullptr, npsf,
                dg.d + (size_t)lo * npg, prior, lm_pars, ignore_zero_weight,
                dr.d + lo, q);
        });
    if (rc) return rc;
    NGX_CUDA(dr.to_host(res));
    return NGX_OK;
}
#endif

#ifndef NGX_LM_WIDE_TU
int ngx_lm_fit_multi(int nobj, int model, const int32_t *stamp_start,
                     const int32_t *stamp_band, int nband,
                     const double *images, const double *weights,
                     const ngx_jacobian *jac, int nrow, int ncol,
                     const double *psf_pars, int npsf, const double *guess,
                     const ngx_prior *prior, const ngx_lm_pars *lm_pars,
                     int ignore_zero_weight, ngx_lm_result *res,
                     void *stream) {
    return lm_fit_multi_impl(nobj, model, stamp_start, stamp_band, nband,
                             images, weights, jac, nrow, ncol, psf_pars, npsf,
                             guess, prior, lm_pars, ignore_zero_weight, res, 0,
                             stream);
}
#endif

#ifdef NGX_LM_WIDE_TU
int ngx_lm_fit_multi_wide(int nobj, int model, const int32_t *stamp_start,
                          const int32_t *stamp_band, int nband,
                          const double *images, const double *weights,
                          const ngx_jacobian *jac, int nrow, int ncol,
                          const double *psf_pars, int npsf,
                          const double *guess, const ngx_prior *prior,
                          const ngx_lm_pars *lm_pars, int ignore_zero_weight,
                          ngx_lm_result_wide *res, void *stream) {
    return lm_fit_multi_impl(nobj, model, stamp_start, stamp_band, nband,
                             images, weights, jac, nrow, ncol, psf_pars, npsf,
                             guess, prior, lm_pars, ignore_zero_weight, res, 1,
                             stream);
}
#endif

}  // extern "C"

template <typename RES>
static int lm_fit_multi_host_impl(int nobj, int model,
                                  const int32_t *stamp_start,
                                  const int32_t *stamp_band, int nband,
                                  const double *images, const double *weights,
                                  const ngx_jacobian *jac, int nrow, int ncol,
                                  const double *psf_pars, int npsf,
                                  const double *guess, const ngx_prior *prior,
                                  const ngx_lm_pars *lm_pars,
                                  int ignore_zero_weight, RES *res,
                                  int wide) {
    NGX_REQUIRE(nobj >= 0 && stamp_start && stamp_band && images && weights);
    NGX_REQUIRE(jac && guess && res && nband >= 1 && nband <= NGX_MAX_BANDS);
    if (nobj == 0) return NGX_OK;
    const int ns = stamp_start[nobj];
    NGX_REQUIRE(ns >= 0 && stamp_start[0] == 0);
    const size_t ntot = (size_t)nrow * ncol;
    const int np = model_npars(model) - 1 + nband;
    // chunks of objects, double buffered: the copy of chunk c + 1 runs while
    // chunk c is fitted (one unpipelined copy of the whole batch was 8x
    // slower than the kernel: 38k vs 314k objects/s on 2 bands x 2 epochs)
    int chunk = 4 * sm_count();
    if (chunk > nobj) chunk = nobj;
    size_t max_st = 0;          // most stamps in any chunk
    for (int lo = 0; lo < nobj; lo += chunk) {
        int hi = lo + chunk < nobj ? lo + chunk : nobj;
        size_t m = (size_t)(stamp_start[hi] - stamp_start[lo]);
        if (m > max_st) max_st = m;
    }
    // streams and device buffers persist per calling thread (allocation and
    // release per call cost more than the fits of a small batch, and
    // cudaFree synchronises the device)
    static thread_local cudaStream_t st[2] = {nullptr, nullptr};
    static thread_local char *arena[2] = {nullptr, nullptr};
    static thread_local size_t arena_cap = 0;
    if (!st[0]) {
        NGX_CUDA(cudaStreamCreate(&st[0]));
        NGX_CUDA(cudaStreamCreate(&st[1]));
    }
    const size_t npf = npsf > 0 ? (size_t)npsf * 6 : 0;
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t o_ss = 0;
    const size_t o_sb = o_ss + al(sizeof(int32_t) * (chunk + 1));
    const size_t o_im = o_sb + al(sizeof(int32_t) * max_st);
    const size_t o_wt = o_im + al(sizeof(double) * max_st * ntot);
    const size_t o_jac = o_wt + al(sizeof(double) * max_st * ntot);
    const size_t o_psf = o_jac + al(sizeof(ngx_jacobian) * max_st);
    const size_t o_g = o_psf + al(sizeof(double) * max_st * npf);
    const size_t o_res = o_g + al(sizeof(double) * (size_t)chunk * np);
    const size_t need = o_res + al(sizeof(RES) * (size_t)chunk);
    if (need > arena_cap) {
        NGX_CUDA(cudaDeviceSynchronize());
        for (int b = 0; b < 2; b++) {
            cudaFree(arena[b]);
            arena[b] = nullptr;
        }
        arena_cap = 0;
        for (int b = 0; b < 2; b++)
            NGX_CUDA(cudaMalloc((void **)&arena[b], need));
        arena_cap = need;
    }
    int rc = NGX_OK;
    {
        std::vector<int32_t> rebased[2];
        // records come back asynchronously: into the caller's array when it
        // is pinned, else into a pinned staging array (a device-to-host copy
        // into pageable memory would block the host until the chunk's fits
        // are done and keep the next chunk's copies from being queued)
        static thread_local RES *h_stage = nullptr;
        static thread_local size_t cap_stage = 0;
        cudaPointerAttributes pr;
        const bool res_pinned =
            cudaPointerGetAttributes(&pr, res) == cudaSuccess
            && pr.type == cudaMemoryTypeHost;
        cudaGetLastError();
        if (!res_pinned && (size_t)nobj > cap_stage) {
            if (h_stage) cudaFreeHost(h_stage);
            h_stage = nullptr;
            cap_stage = 0;
            NGX_CUDA(cudaMallocHost((void **)&h_stage,
                                    sizeof(RES) * (size_t)nobj));
            cap_stage = nobj;
        }
        RES *res_dst = res_pinned ? res : h_stage;
        cudaPointerAttributes pa, pb;
        const bool planes_pinned =
            cudaPointerGetAttributes(&pa, images) == cudaSuccess
            && pa.type == cudaMemoryTypeHost
            && cudaPointerGetAttributes(&pb, weights) == cudaSuccess
            && pb.type == cudaMemoryTypeHost;
        cudaGetLastError();
        for (int lo = 0, c = 0; lo < nobj && rc == NGX_OK; lo += chunk, c++) {
            const int b = c & 1;
            const int hi = lo + chunk < nobj ? lo + chunk : nobj;
            const int cn = hi - lo;
            const int s0 = stamp_start[lo];
            const size_t m = (size_t)(stamp_start[hi] - s0);
            cudaStream_t q = st[b];
            char *A = arena[b];
            int32_t *dss = (int32_t *)(A + o_ss), *dsb = (int32_t *)(A + o_sb);
            double *di = (double *)(A + o_im), *dw = (double *)(A + o_wt);
            ngx_jacobian *dj = (ngx_jacobian *)(A + o_jac);
            double *dp = (double *)(A + o_psf), *dg = (double *)(A + o_g);
            RES *dr = (RES *)(A + o_res);
            // the rebased starts live in a host vector reused every second
            // chunk: wait for its previous copy before overwriting
            // (buffer set b was last used by chunk c - 2 on the same stream)
            cudaStreamSynchronize(q);
            rebased[b].resize(cn + 1);
            for (int i = 0; i <= cn; i++)
                rebased[b][i] = stamp_start[lo + i] - s0;
            cudaMemcpyAsync(dss, rebased[b].data(), sizeof(int32_t) * (cn + 1),
                            cudaMemcpyHostToDevice, q);
            cudaMemcpyAsync(dsb, stamp_band + s0, sizeof(int32_t) * m,
                            cudaMemcpyHostToDevice, q);
            if (planes_pinned) {
                cudaMemcpyAsync(di, images + (size_t)s0 * ntot,
                                sizeof(double) * m * ntot,
                                cudaMemcpyHostToDevice, q);
                cudaMemcpyAsync(dw, weights + (size_t)s0 * ntot,
                                sizeof(double) * m * ntot,
                                cudaMemcpyHostToDevice, q);
            } else {
                // pageable planes: through pinned slots filled by the copy
                // pool (the driver's own staging copy runs on one thread)
                StageRing::get().h2d(di, images + (size_t)s0 * ntot,
                                     sizeof(double) * m * ntot, q);
                StageRing::get().h2d(dw, weights + (size_t)s0 * ntot,
                                     sizeof(double) * m * ntot, q);
            }
            cudaMemcpyAsync(dj, jac + s0, sizeof(ngx_jacobian) * m,
                            cudaMemcpyHostToDevice, q);
            if (npf)
                cudaMemcpyAsync(dp, psf_pars + (size_t)s0 * npf,
                                sizeof(double) * m * npf,
                                cudaMemcpyHostToDevice, q);
            cudaMemcpyAsync(dg, guess + (size_t)lo * np,
                            sizeof(double) * (size_t)cn * np,
                            cudaMemcpyHostToDevice, q);
            rc = lm_fit_multi_impl(cn, model, dss, dsb, nband, di, dw, dj,
                                   nrow, ncol, npf ? dp : nullptr, npsf, dg,
                                   prior, lm_pars, ignore_zero_weight, dr,
                                   wide, q);
            if (rc == NGX_OK)
                cudaMemcpyAsync(res_dst + lo, dr, sizeof(RES) * cn,
                                cudaMemcpyDeviceToHost, q);
        }
        cudaError_t e0 = cudaStreamSynchronize(st[0]);
        cudaError_t e1 = cudaStreamSynchronize(st[1]);
        if (rc) return rc;
        NGX_CUDA(e0);
        NGX_CUDA(e1);
        NGX_CUDA(cudaGetLastError());
        if (res_dst != res)
            memcpy(res, res_dst, sizeof(RES) * (size_t)nobj);
    }
    return NGX_OK;
}

extern "C" {

#ifndef NGX_LM_WIDE_TU
int ngx_lm_fit_multi_host(int nobj, int model, const int32_t *stamp_start,
                          const int32_t *stamp_band, int nband,
                          const double *images, const double *weights,
                          const ngx_jacobian *jac, int nrow, int ncol,
                          const double *psf_pars, int npsf,
                          const double *guess, const ngx_prior *prior,
                          const ngx_lm_pars *lm_pars, int ignore_zero_weight,
                          ngx_lm_result *res) {
    NGX_REQUIRE(nband >= 1);
    return lm_fit_multi_host_impl(nobj, model, stamp_start, stamp_band, nband,
                                  images, weights, jac, nrow, ncol, psf_pars,
                                  npsf, guess, prior, lm_pars,
                                  ignore_zero_weight, res, 0);
}
#endif

#ifdef NGX_LM_WIDE_TU
int ngx_lm_fit_multi_wide_host(int nobj, int model,
                               const int32_t *stamp_start,
                               const int32_t *stamp_band, int nband,
                               const double *images, const double *weights,
                               const ngx_jacobian *jac, int nrow, int ncol,
                               const double *psf_pars, int npsf,
                               const double *guess, const ngx_prior *prior,
                               const ngx_lm_pars *lm_pars,
                               int ignore_zero_weight,
                               ngx_lm_result_wide *res) {
    return lm_fit_multi_host_impl(nobj, model, stamp_start, stamp_band, nband,
                                  images, weights, jac, nrow, ncol, psf_pars,
                                  npsf, guess, prior, lm_pars,
                                  ignore_zero_weight, res, 1);
}
#endif

#ifndef NGX_LM_WIDE_TU
// ---------------------------------------------------------------------
// Streamed host-buffer pipeline (pinned stamps that fit the resident
// kernels).  ONE persistent launch per batch of up to NGX_STREAM_BATCH
// stamps: the copy engine writes the stamps in the host's own format
// (double planes, float planes or the compact format) chunk by chunk, each
// chunk followed by a 4-byte copy that raises its ready flag; the CTAs take
// stamps from the work queue in order and wait for the flag of a stamp's
// chunk before staging it into shared memory (float stamps are widened
// there: no intermediate planes in HBM, no expansion kernels).  Finished
// records are counted per chunk of 1024; the CTA that completes a chunk
// raises a flag in mapped pinned memory and the host, polling, starts that
// chunk's device-to-host copy while the fits go on.  Against one launch per
// chunk this removes the launch tails (a fit takes ~0.25 ms, so every
// chunk launch ended with SMs draining) and the per-chunk expansion passes.
struct StreamBuf {
    void *in_a = nullptr, *in_b = nullptr;      // image / weight planes
    ngx_jacobian *jac = nullptr;
    double *psf = nullptr, *guess = nullptr, *ws = nullptr;
    unsigned char *mask = nullptr;
    ngx_lm_result *res = nullptr;
    int *flags = nullptr;                       // ready | done_count
    size_t cap_a = 0, cap_b = 0, cap_obj = 0, cap_psf = 0, cap_mask = 0,
           cap_flags = 0;
    cudaEvent_t small_up = nullptr, drained = nullptr;
};

static int grow_dev(void **p, size_t *cap, size_t need) {
    if (need <= *cap) return NGX_OK;
    NGX_CUDA(cudaDeviceSynchronize());
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    NGX_CUDA(cudaMalloc(p, need));
    *cap = need;
    return NGX_OK;
}

static bool is_pinned(const void *p) {
    cudaPointerAttributes attr;
    bool ok = cudaPointerGetAttributes(&attr, p) == cudaSuccess
        && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    return ok;
}

static bool lm_stream_eligible(int model, int nrow, int ncol, int npsf,
                               const ngx_lm_pars *lm_pars) {
    if (const char *e = getenv("NGX_NO_STREAM"))
        if (atoi(e)) return false;
    int np = model_npars(model), ng0 = model_ngauss(model);
    if (np <= 0 || ng0 <= 0 || model == NGX_CM || !lm_pars) return false;
    const bool analytic = lm_pars->analytic_jacobian
        && (model == NGX_GAUSS || model == NGX_EXP || model == NGX_DEV);
    const int ngauss = ng0 * (npsf > 0 ? npsf : 1);
    const int threads = nrow * ncol <= 1024 ? 64 : 128;
    const int nq = 32 * ((ncol + 7) / 8) * ((nrow + 3) / 4);
    size_t smem = 0;
    return !lm_weights_global(ngauss, analytic ? 0 : np + 1, nq, nrow, ncol,
                              threads / 32, &smem);
}

static int lm_fit_host_stream(int n, int model, const void *images_v,
                              const void *weights_v, int f32,
                              const unsigned char *maskbits,
                              const ngx_jacobian *jac, int nrow, int ncol,
                              const double *psf_pars, int npsf,
                              const double *guess, const ngx_prior *prior,
                              const ngx_lm_pars *lm_pars,
                              int ignore_zero_weight, ngx_lm_result *res) {
    const int np = model_npars(model);
    const size_t ntot = (size_t)nrow * ncol;
    const int nmb = (int)((ntot + 7) / 8);
    const size_t esz = f32 ? sizeof(float) : sizeof(double);
    const bool compact = f32 == 2;
    int batch = 32768;
    if (const char *e = getenv("NGX_STREAM_BATCH")) {
        int w = atoi(e);
        if (w >= 1) batch = w;
    }
    if (batch > n) batch = n;
    // input chunks (wait_ready): S, S, 2S, 4S and then 8S stamps each
    const int S = sm_count(), rchunk = 8 * S;
    const int dchunk = 1024;
    const int max_rc = 4 + (batch + rchunk - 1) / rchunk;
    const int max_dc = (batch + dchunk - 1) / dchunk;

    static thread_local StreamBuf sb[2];
    static thread_local cudaStream_t cs = nullptr, ks[2] = {nullptr, nullptr},
                                     ds = nullptr;
    static thread_local int *h_flags = nullptr;   // mapped: done flags, err
    static thread_local int *h_one = nullptr;     // pinned constant 1
    static thread_local size_t cap_hflags = 0;
    static thread_local ngx_lm_result *h_stage = nullptr;
    static thread_local size_t cap_stage = 0;
    if (!cs) {
        NGX_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        NGX_CUDA(cudaStreamCreateWithFlags(&ks[0], cudaStreamNonBlocking));
        NGX_CUDA(cudaStreamCreateWithFlags(&ks[1], cudaStreamNonBlocking));
        NGX_CUDA(cudaStreamCreateWithFlags(&ds, cudaStreamNonBlocking));
        NGX_CUDA(cudaHostAlloc((void **)&h_one, sizeof(int),
                               cudaHostAllocDefault));
        *h_one = 1;
        for (int b = 0; b < 2; b++) {
            NGX_CUDA(cudaEventCreateWithFlags(&sb[b].small_up,
                                              cudaEventDisableTiming));
            NGX_CUDA(cudaEventCreateWithFlags(&sb[b].drained,
                                              cudaEventDisableTiming));
        }
    }
    const size_t need_hflags = 2 * (size_t)(max_dc + 1);
    if (need_hflags > cap_hflags) {
        NGX_CUDA(cudaDeviceSynchronize());
        if (h_flags) cudaFreeHost(h_flags);
        h_flags = nullptr;
        cap_hflags = 0;
        NGX_CUDA(cudaHostAlloc((void **)&h_flags, need_hflags * sizeof(int),
                               cudaHostAllocMapped));
        cap_hflags = need_hflags;
    }
    // pageable stamps go through a ring of pinned slots, filled by a few host
    // threads (ngx_copy_pool.h) while the copy engine drains the other slots
    const bool in_pinned = is_pinned(images_v)
        && (compact || is_pinned(weights_v));
    const bool res_pinned = is_pinned(res);
    if (!res_pinned && (size_t)n > cap_stage) {
        if (h_stage) cudaFreeHost(h_stage);
        h_stage = nullptr;
        cap_stage = 0;
        NGX_CUDA(cudaMallocHost((void **)&h_stage,
                                (size_t)n * sizeof(ngx_lm_result)));
        cap_stage = n;
    }
    ngx_lm_result *res_dst = res_pinned ? res : h_stage;
    const int nbatch = (n + batch - 1) / batch;
    // results of the previous done-chunk wait in the staging buffer until
    // their copy has finished (pageable result arrays)
    cudaEvent_t d2h_ev[2] = {nullptr, nullptr};
    if (!res_pinned)
        for (int i = 0; i < 2; i++)
            NGX_CUDA(cudaEventCreateWithFlags(&d2h_ev[i],
                                              cudaEventDisableTiming));
    int rc = NGX_OK;
    int owed_lo = -1, owed_n = 0, owed_ev = 0, nd2h = 0;

    // test hook: stop queueing the feed at input chunk NGX_STREAM_FAIL_AT
    // (the call must fail promptly and leave the device usable)
    int fail_at = -1;
    if (const char *e = getenv("NGX_STREAM_FAIL_AT")) fail_at = atoi(e);
    int next_c[2] = {0, 0};     // next done-chunk to copy out, per buffer
    std::function<int(int, bool)> drain;
    int cur_batch = -1;
    // a large host-to-device copy on the copy stream
    auto h2d = [&](void *dst, const void *src, size_t bytes) -> int {
        if (in_pinned) {
            NGX_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice,
                                     cs));
            return NGX_OK;
        }
        // records that are ready leave while the host is busy copying
        const std::function<int()> between = [&]() -> int {
            for (int kk = cur_batch - 1; kk <= cur_batch; kk++)
                if (kk >= 0) {
                    int r = drain(kk, false);
                    if (r) return r;
                }
            return NGX_OK;
        };
        int brc = NGX_OK;
        NGX_CUDA(StageRing::get().h2d(dst, src, bytes, cs, &between, &brc));
        return brc;
    };

    // queue the input copies of batch k and launch its kernel
    auto enqueue = [&](int k) -> int {
        StreamBuf &B = sb[k & 1];
        const int lo = k * batch;
        const int cn = n - lo < batch ? n - lo : batch;
        int *hf = h_flags + (size_t)(k & 1) * (max_dc + 1);
        int r;
        if ((r = grow_dev(&B.in_a, &B.cap_a, (size_t)cn * ntot * esz))) return r;
        if (!compact
            && (r = grow_dev(&B.in_b, &B.cap_b, (size_t)cn * ntot * esz)))
            return r;
        if (compact && maskbits
            && (r = grow_dev((void **)&B.mask, &B.cap_mask, (size_t)cn * nmb)))
            return r;
        if ((size_t)cn > B.cap_obj) {
            NGX_CUDA(cudaDeviceSynchronize());
            cudaFree(B.jac); cudaFree(B.guess); cudaFree(B.ws); cudaFree(B.res);
            B.jac = nullptr; B.guess = nullptr; B.ws = nullptr; B.res = nullptr;
            B.cap_obj = 0;
            NGX_CUDA(cudaMalloc((void **)&B.jac, cn * sizeof(ngx_jacobian)));
            NGX_CUDA(cudaMalloc((void **)&B.guess,
                                (size_t)cn * 8 * sizeof(double)));
            NGX_CUDA(cudaMalloc((void **)&B.ws, cn * sizeof(double)));
            NGX_CUDA(cudaMalloc((void **)&B.res, cn * sizeof(ngx_lm_result)));
            B.cap_obj = cn;
        }
        if (npsf > 0
            && (r = grow_dev((void **)&B.psf, &B.cap_psf,
                             (size_t)cn * 6 * npsf * sizeof(double))))
            return r;
        if ((r = grow_dev((void **)&B.flags, &B.cap_flags,
                          (size_t)(max_rc + max_dc + 1) * sizeof(int))))
            return r;
        int *d_ready = B.flags, *d_done = B.flags + max_rc;
        next_c[k & 1] = 0;
        // the buffers are free once the records of batch k-2 are on the host
        NGX_CUDA(cudaStreamWaitEvent(cs, B.drained, 0));
        NGX_CUDA(cudaMemsetAsync(B.flags, 0,
                                 (size_t)(max_rc + max_dc + 1) * sizeof(int),
                                 cs));
        for (int i = 0; i <= max_dc; i++) hf[i] = 0;
        NGX_CUDA(cudaEventRecord(B.small_up, cs));
        NGX_CUDA(cudaStreamWaitEvent(ks[k & 1], B.small_up, 0));
        LMFeed feed;
        memset(&feed, 0, sizeof(feed));
        feed.ready = d_ready; feed.ready_chunk = rchunk;
        feed.src = f32; feed.nmb = nmb; feed.done_chunk = dchunk;
        feed.im32 = (const float *)B.in_a;
        feed.wt32 = (const float *)B.in_b;
        feed.wscalar = B.ws;
        feed.maskbits = (compact && maskbits) ? B.mask : nullptr;
        feed.done_count = d_done;
        NGX_CUDA(cudaHostGetDevicePointer((void **)&feed.host_done, hf, 0));
        feed.err = feed.host_done + max_dc;
        feed.err_dev = B.flags + max_rc + max_dc;
        auto launch = [&]() -> int {
            g_lm_feed = &feed;
            int lr = ngx_lm_fit(cn, model, (const double *)B.in_a,
                                (const double *)(compact ? B.in_a : B.in_b),
                                B.jac, nrow, ncol,
                                npsf > 0 ? B.psf : nullptr, npsf, B.guess,
                                prior, lm_pars, ignore_zero_weight, B.res,
                                ks[k & 1]);
            g_lm_feed = nullptr;
            return lr;
        };
        // Pinned stamps: every copy of the batch is queued (asynchronously,
        // ~70 us of host time - about when the first chunk lands) BEFORE the
        // launch, so the call also works under tools that make launches
        // synchronous (Nsight Compute).  Pageable stamps are staged by host
        // threads as the fits go on, so there the kernel has to start first.
        if (!in_pinned && (r = launch())) return r;
        cur_batch = in_pinned ? k - 1 : k;
        for (int c = 0, slo = 0; slo < cn; c++) {
            const int want = c < 2 ? S : c == 2 ? 2 * S : c == 3 ? 4 * S
                                                         : rchunk;
            const int sn = cn - slo < want ? cn - slo : want;
            const size_t g = (size_t)(lo + slo);
            if (fail_at >= 0 && c == fail_at)
                return fail(NGX_ERR_CUDA,
                            "ngx_lm_fit_host: feed interrupted (test hook "
                            "NGX_STREAM_FAIL_AT)%s%s");
            if ((r = h2d((char *)B.in_a + (size_t)slo * ntot * esz,
                         (const char *)images_v + g * ntot * esz,
                         (size_t)sn * ntot * esz)))
                return r;
            if (compact) {
                NGX_CUDA(cudaMemcpyAsync(
                    B.ws + slo, (const double *)weights_v + g,
                    sn * sizeof(double), cudaMemcpyHostToDevice, cs));
                if (maskbits)
                    NGX_CUDA(cudaMemcpyAsync(
                        B.mask + (size_t)slo * nmb, maskbits + g * nmb,
                        (size_t)sn * nmb, cudaMemcpyHostToDevice, cs));
            } else if ((r = h2d((char *)B.in_b + (size_t)slo * ntot * esz,
                                (const char *)weights_v + g * ntot * esz,
                                (size_t)sn * ntot * esz))) {
                return r;
            }
            if (c == 0 || (c >= 4 && (c - 4) % 8 == 0)) {
                // the small per-stamp arrays: of the first 8S stamps, then
                // of eight chunks at a time
                const int gw = c == 0 ? rchunk : 8 * rchunk;
                const int gn = cn - slo < gw ? cn - slo : gw;
                NGX_CUDA(cudaMemcpyAsync(B.jac + slo, jac + g,
                                         gn * sizeof(ngx_jacobian),
                                         cudaMemcpyHostToDevice, cs));
                if (npsf > 0)
                    NGX_CUDA(cudaMemcpyAsync(
                        B.psf + (size_t)slo * 6 * npsf,
                        psf_pars + g * 6 * npsf,
                        (size_t)gn * 6 * npsf * sizeof(double),
                        cudaMemcpyHostToDevice, cs));
                NGX_CUDA(cudaMemcpyAsync(B.guess + (size_t)slo * np,
                                         guess + g * np,
                                         (size_t)gn * np * sizeof(double),
                                         cudaMemcpyHostToDevice, cs));
            }
            NGX_CUDA(cudaMemcpyAsync(d_ready + c, h_one, sizeof(int),
                                     cudaMemcpyHostToDevice, cs));
            slo += sn;
        }
        if (in_pinned && (r = launch())) return r;
        cur_batch = k;
        return NGX_OK;
    };

    // staged records whose copy has finished go to the caller's array
    auto flush_owed = [&]() -> int {
        if (owed_lo < 0) return NGX_OK;
        NGX_CUDA(cudaEventSynchronize(d2h_ev[owed_ev]));
        memcpy(res + owed_lo, h_stage + owed_lo,
               (size_t)owed_n * sizeof(*res));
        owed_lo = -1;
        return NGX_OK;
    };

    const bool trace_d = getenv("NGX_STREAM_TRACE") != nullptr;
    auto now_us2 = []() {
        return std::chrono::duration<double, std::micro>(
            std::chrono::steady_clock::now().time_since_epoch()).count();
    };
    const double t_begin2 = now_us2();
    // copy the records of batch k out as its chunks complete (block = false:
    // only those that are complete now)
    drain = [&](int k, bool block) -> int {
        StreamBuf &B = sb[k & 1];
        const int lo = k * batch;
        const int cn = n - lo < batch ? n - lo : batch;
        const int ndc = (cn + dchunk - 1) / dchunk;
        volatile int *hf = h_flags + (size_t)(k & 1) * (max_dc + 1);
        bool finished = false;
        for (int &c = next_c[k & 1]; c < ndc; c++) {
            while (!finished && !hf[c]) {
                cudaError_t q = cudaStreamQuery(ks[k & 1]);
                if (q == cudaSuccess) finished = true;
                else if (q != cudaErrorNotReady) NGX_CUDA(q);
                else if (!block) return NGX_OK;
            }
            const int slo = c * dchunk;
            const int sn = cn - slo < dchunk ? cn - slo : dchunk;
            if (trace_d)
                fprintf(stderr, "[ngx stream] records %d.. done at %.0f us%s\n",
                        lo + slo, now_us2() - t_begin2,
                        finished ? " (kernel finished)" : "");
            NGX_CUDA(cudaMemcpyAsync(res_dst + lo + slo, B.res + slo,
                                     sn * sizeof(ngx_lm_result),
                                     cudaMemcpyDeviceToHost, ds));
            if (!res_pinned) {
                const int e = nd2h++ & 1;
                NGX_CUDA(cudaEventRecord(d2h_ev[e], ds));
                int r = flush_owed();
                if (r) return r;
                owed_lo = lo + slo; owed_n = sn; owed_ev = e;
            }
        }
        if (!block) return NGX_OK;
        NGX_CUDA(cudaStreamSynchronize(ks[k & 1]));
        NGX_CUDA(cudaEventRecord(B.drained, ds));
        if (hf[max_dc])
            return fail(NGX_ERR_CUDA,
                        "ngx_lm_fit_host: the stamp feed stalled%s%s");
        return NGX_OK;
    };

    const bool trace = getenv("NGX_STREAM_TRACE") != nullptr;
    auto now_us = []() {
        return std::chrono::duration<double, std::micro>(
            std::chrono::steady_clock::now().time_since_epoch()).count();
    };
    const double t_begin = trace ? now_us() : 0.0;
    for (int k = 0; k < nbatch && !rc; k++) {
        rc = enqueue(k);
        if (trace)
            fprintf(stderr, "[ngx stream] batch %d queued at %.0f us\n", k,
                    now_us() - t_begin);
        if (!rc && k > 0) rc = drain(k - 1, true);
    }
    if (!rc) rc = drain(nbatch - 1, true);
    if (trace)
        fprintf(stderr, "[ngx stream] kernels done at %.0f us\n",
                now_us() - t_begin);
    if (!rc) rc = flush_owed();
    if (rc) {
        // release kernels that still wait for stamps that will not come: the
        // device-side flag the waiters poll, raised by a copy on the record
        // stream (the copy stream may be the one that failed)
        for (int b = 0; b < 2; b++)
            if (sb[b].flags && sb[b].cap_flags
                    >= (size_t)(max_rc + max_dc + 1) * sizeof(int))
                cudaMemcpyAsync(sb[b].flags + max_rc + max_dc, h_one,
                                sizeof(int), cudaMemcpyHostToDevice, ds);
    }
    cudaStreamSynchronize(cs);
    cudaStreamSynchronize(ks[0]);
    cudaStreamSynchronize(ks[1]);
    cudaStreamSynchronize(ds);
    if (trace)
        fprintf(stderr, "[ngx stream] call done at %.0f us\n",
                now_us() - t_begin);
    if (!res_pinned)
        for (int i = 0; i < 2; i++) cudaEventDestroy(d2h_ev[i]);
    return rc;
}
#endif

#ifndef NGX_LM_WIDE_TU
// the host-buffer pipeline behind ngx_lm_fit_host (stamps as doubles) and
// ngx_lm_fit_host_f32 (stamps as floats: half the bytes over PCIe, which is
// what bounds the call; widened to double on the device, exactly)
static int lm_fit_host_impl(int n, int model, const void *images_v,
                            const void *weights_v, int f32,
                            const unsigned char *maskbits,
                            const ngx_jacobian *jac, int nrow, int ncol,
                            const double *psf_pars, int npsf,
                            const double *guess, const ngx_prior *prior,
                            const ngx_lm_pars *lm_pars,
                            int ignore_zero_weight, ngx_lm_result *res) {
    NGX_REQUIRE(n >= 0 && images_v && weights_v && jac && guess && res);
    if (n == 0) return NGX_OK;
    NGX_REQUIRE(nrow >= 1 && ncol >= 1 && npsf >= 0 && npsf <= 16);
    NGX_REQUIRE(lm_pars && lm_pars->maxfev > 0);
    // streaming pays once there is more than one chunk to overlap: two waves
    // of pinned stamps; pageable ones from ~32 MB (below that the driver's
    // own staging copy is as fast as waking the copy pool)
    const size_t plane_bytes = (size_t)n * nrow * ncol
        * (f32 ? sizeof(float) : sizeof(double)) * (f32 == 2 ? 1 : 2);
    const bool pinned_in = is_pinned(images_v)
        && (f32 == 2 || is_pinned(weights_v));
    // Pageable stamps are staged while the resident kernel already runs; a
    // tool or setting that makes launches synchronous (Nsight Compute, which
    // announces itself to the target through these variables;
    // CUDA_LAUNCH_BLOCKING) would stall that feed,
    // so there they keep the chunked pipeline.  Pinned stamps are queued
    // before the launch and stream everywhere.
    const char *clb = getenv("CUDA_LAUNCH_BLOCKING");
    const bool sync_launches = getenv("NV_NSIGHT_INJECTION_PORT_BASE")
        || getenv("NV_TPS_LAUNCH_TOKEN") || (clb && atoi(clb) != 0);
    const bool worth = pinned_in ? n >= 2 * sm_count()
                                 : plane_bytes >= ((size_t)32 << 20);
    if ((worth || getenv("NGX_STREAM_ALWAYS"))
        && (pinned_in || !sync_launches)
        && lm_stream_eligible(model, nrow, ncol, npsf, lm_pars))
        return lm_fit_host_stream(n, model, images_v, weights_v, f32,
                                  maskbits, jac, nrow, ncol, psf_pars, npsf,
                                  guess, prior, lm_pars, ignore_zero_weight,
                                  res);
    const double *images = (const double *)images_v;
    const double *weights = (const double *)weights_v;
    const float *images32 = (const float *)images_v;
    const float *weights32 = (const float *)weights_v;
    // f32 == 2: compact format - float32 pixels, one double weight per stamp
    // (weights_v), optional packed mask bits
    const bool compact = f32 == 2;
    const double *wscalar = (const double *)weights_v;
    int np = model_npars(model);
    NGX_REQUIRE(np > 0);
    // chunked, double-buffered pipeline: H2D of chunk i+1 overlaps the fit
    // of chunk i when the host arrays are pinned
    const size_t ntot = (size_t)nrow * ncol;
    // chunk schedule: the pipeline is bound by the H2D copy (PCIe), so what
    // is exposed is the first chunk's copy and the last chunk's fit: a short
    // first chunk (one wave of CTAs, ~20 MB) and then chunks of two waves,
    // whose fit (~0.6 ms) hides behind the next chunk's copy (~0.8 ms)
    const int wave = 4 * sm_count();
    int waves = 2;
    if (const char *e = getenv("NGX_CHUNK_WAVES")) {
        int w = atoi(e);
        if (w >= 1 && w <= 64) waves = w;
    }
    int chunk = waves * wave;
    if (chunk > n) chunk = n;
    int first_chunk = wave < chunk ? wave : chunk;
    static thread_local double *d_im[2] = {nullptr, nullptr};
    static thread_local double *d_wt[2] = {nullptr, nullptr};
    static thread_local ngx_jacobian *d_jac[2] = {nullptr, nullptr};
    static thread_local double *d_psf[2] = {nullptr, nullptr};
    static thread_local double *d_guess[2] = {nullptr, nullptr};
    static thread_local ngx_lm_result *d_res[2] = {nullptr, nullptr};
    static thread_local size_t cap_pix = 0, cap_obj = 0, cap_psf = 0;
    static thread_local float *d_im32[2] = {nullptr, nullptr};
    static thread_local float *d_wt32[2] = {nullptr, nullptr};
    static thread_local size_t cap_pix32 = 0;
    static thread_local double *d_ws[2] = {nullptr, nullptr};
    static thread_local unsigned char *d_mask[2] = {nullptr, nullptr};
    static thread_local size_t cap_ws = 0, cap_mask = 0;
    const int nmb = (int)((ntot + 7) / 8);
    static thread_local ngx_lm_result *h_stage = nullptr;   // pinned
    static thread_local size_t cap_stage = 0;
    // streams[1] copies in; the fits of even / odd chunks run on streams[0] /
    // streams[2] with their own work-queue counters, so the next chunk's CTAs
    // fill the tail of the previous chunk's launch
    static thread_local cudaStream_t streams[3] = {nullptr, nullptr, nullptr};
    if (!streams[0]) {
        NGX_CUDA(cudaStreamCreate(&streams[0]));
        NGX_CUDA(cudaStreamCreate(&streams[1]));
        NGX_CUDA(cudaStreamCreate(&streams[2]));
    }
    size_t need_pix = (size_t)chunk * ntot, need_obj = chunk;
    size_t need_psf = (size_t)chunk * 6 * (npsf > 0 ? npsf : 1);
    if (need_pix > cap_pix || need_obj > cap_obj || need_psf > cap_psf) {
        NGX_CUDA(cudaDeviceSynchronize());
        for (int b = 0; b < 2; b++) {
            cudaFree(d_im[b]); cudaFree(d_wt[b]); cudaFree(d_jac[b]);
            cudaFree(d_psf[b]); cudaFree(d_guess[b]); cudaFree(d_res[b]);
            NGX_CUDA(cudaMalloc((void **)&d_im[b], need_pix * sizeof(double)));
            NGX_CUDA(cudaMalloc((void **)&d_wt[b], need_pix * sizeof(double)));
            NGX_CUDA(cudaMalloc((void **)&d_jac[b],
                                need_obj * sizeof(ngx_jacobian)));
            NGX_CUDA(cudaMalloc((void **)&d_psf[b],
                                need_psf * sizeof(double)));
            NGX_CUDA(cudaMalloc((void **)&d_guess[b],
                                need_obj * 8 * sizeof(double)));
            NGX_CUDA(cudaMalloc((void **)&d_res[b],
                                need_obj * sizeof(ngx_lm_result)));
        }
        cap_pix = need_pix; cap_obj = need_obj; cap_psf = need_psf;
    }
    if (f32 && need_pix > cap_pix32) {
        NGX_CUDA(cudaDeviceSynchronize());
        for (int b = 0; b < 2; b++) {
            cudaFree(d_im32[b]); cudaFree(d_wt32[b]);
            NGX_CUDA(cudaMalloc((void **)&d_im32[b], need_pix * sizeof(float)));
            NGX_CUDA(cudaMalloc((void **)&d_wt32[b], need_pix * sizeof(float)));
        }
        cap_pix32 = need_pix;
    }
    if (compact && ((size_t)chunk > cap_ws
                    || (size_t)chunk * nmb > cap_mask)) {
        NGX_CUDA(cudaDeviceSynchronize());
        for (int b = 0; b < 2; b++) {
            cudaFree(d_ws[b]); cudaFree(d_mask[b]);
            NGX_CUDA(cudaMalloc((void **)&d_ws[b], chunk * sizeof(double)));
            NGX_CUDA(cudaMalloc((void **)&d_mask[b], (size_t)chunk * nmb));
        }
        cap_ws = chunk; cap_mask = (size_t)chunk * nmb;
    }
    // results: device -> pinned memory asynchronously.  If the caller's
    // array is itself pinned it is the target, else a pinned staging buffer
    // is copied out by the host while the next chunk runs.
    cudaPointerAttributes attr;
    bool res_pinned = cudaPointerGetAttributes(&attr, res) == cudaSuccess
        && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (!res_pinned && (size_t)n > cap_stage) {
        if (h_stage) cudaFreeHost(h_stage);
        NGX_CUDA(cudaMallocHost((void **)&h_stage,
                                (size_t)n * sizeof(ngx_lm_result)));
        cap_stage = n;
    }
    ngx_lm_result *res_dst = res_pinned ? res : h_stage;
    cudaEvent_t up[2], done[2];
    for (int b = 0; b < 2; b++) {
        NGX_CUDA(cudaEventCreateWithFlags(&up[b], cudaEventDisableTiming));
        NGX_CUDA(cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming));
    }
    int nchunks = 1 + (n - first_chunk + chunk - 1) / chunk;
    for (int c = 0; c < nchunks; c++) {
        int b = c & 1;
        int lo = c == 0 ? 0 : first_chunk + (c - 1) * chunk;
        int cn = c == 0 ? first_chunk : (n - lo < chunk ? n - lo : chunk);
        // buffer b is free once the kernel of chunk c-2 has finished and its
        // results were copied out (both on this buffer's compute stream)
        cudaStream_t ks = streams[b ? 2 : 0];
        NGX_CUDA(cudaStreamWaitEvent(streams[1], done[b], 0));
        if (compact) {
            NGX_CUDA(cudaMemcpyAsync(d_im32[b], images32 + lo * ntot,
                                     cn * ntot * sizeof(float),
                                     cudaMemcpyHostToDevice, streams[1]));
            NGX_CUDA(cudaMemcpyAsync(d_ws[b], wscalar + lo,
                                     cn * sizeof(double),
                                     cudaMemcpyHostToDevice, streams[1]));
            if (maskbits)
                NGX_CUDA(cudaMemcpyAsync(d_mask[b],
                                         maskbits + (size_t)lo * nmb,
                                         (size_t)cn * nmb,
                                         cudaMemcpyHostToDevice, streams[1]));
        } else if (f32) {
            NGX_CUDA(cudaMemcpyAsync(d_im32[b], images32 + lo * ntot,
                                     cn * ntot * sizeof(float),
                                     cudaMemcpyHostToDevice, streams[1]));
            NGX_CUDA(cudaMemcpyAsync(d_wt32[b], weights32 + lo * ntot,
                                     cn * ntot * sizeof(float),
                                     cudaMemcpyHostToDevice, streams[1]));
        } else {
            NGX_CUDA(cudaMemcpyAsync(d_im[b], images + lo * ntot,
                                     cn * ntot * sizeof(double),
                                     cudaMemcpyHostToDevice, streams[1]));
            NGX_CUDA(cudaMemcpyAsync(d_wt[b], weights + lo * ntot,
                                     cn * ntot * sizeof(double),
                                     cudaMemcpyHostToDevice, streams[1]));
        }
        NGX_CUDA(cudaMemcpyAsync(d_jac[b], jac + lo,
                                 cn * sizeof(ngx_jacobian),
                                 cudaMemcpyHostToDevice, streams[1]));
        if (npsf > 0)
            NGX_CUDA(cudaMemcpyAsync(d_psf[b],
                                     psf_pars + (size_t)lo * 6 * npsf,
                                     (size_t)cn * 6 * npsf * sizeof(double),
                                     cudaMemcpyHostToDevice, streams[1]));
        NGX_CUDA(cudaMemcpyAsync(d_guess[b], guess + (size_t)lo * np,
                                 (size_t)cn * np * sizeof(double),
                                 cudaMemcpyHostToDevice, streams[1]));
        NGX_CUDA(cudaEventRecord(up[b], streams[1]));
        NGX_CUDA(cudaStreamWaitEvent(ks, up[b], 0));
        if (compact) {
            const size_t cnt = (size_t)cn * ntot;
            int blocks = (int)((cnt + 255) / 256);
            if (blocks > 8 * sm_count()) blocks = 8 * sm_count();
            k_expand_compact<<<blocks, 256, 0, ks>>>(
                d_im32[b], d_ws[b], maskbits ? d_mask[b] : nullptr, d_im[b],
                d_wt[b], cn, (int)ntot, nmb);
            NGX_LAUNCH_CHECK();
        } else if (f32) {
            const size_t cnt = (size_t)cn * ntot;
            const int blocks = (int)((cnt / 4 + 255) / 256) + 1;
            k_widen<<<blocks, 256, 0, ks>>>(d_im32[b], d_im[b], cnt);
            k_widen<<<blocks, 256, 0, ks>>>(d_wt32[b], d_wt[b], cnt);
            NGX_LAUNCH_CHECK();
        }
        int rc = ngx_lm_fit(cn, model, d_im[b], d_wt[b], d_jac[b], nrow, ncol,
                            npsf > 0 ? d_psf[b] : nullptr, npsf, d_guess[b],
                            prior, lm_pars, ignore_zero_weight, d_res[b], ks);
        if (rc) return rc;
        NGX_CUDA(cudaMemcpyAsync(res_dst + lo, d_res[b],
                                 cn * sizeof(ngx_lm_result),
                                 cudaMemcpyDeviceToHost, ks));
        NGX_CUDA(cudaEventRecord(done[b], ks));
        if (!res_pinned && c > 0) {
            // chunk c-1 is complete once its event fires; chunk c is queued
            int pb = (c - 1) & 1;
            int plo = c == 1 ? 0 : first_chunk + (c - 2) * chunk;
            NGX_CUDA(cudaEventSynchronize(done[pb]));
            memcpy(res + plo, h_stage + plo, (size_t)(lo - plo) * sizeof(*res));
        }
    }
    NGX_CUDA(cudaStreamSynchronize(streams[0]));
    NGX_CUDA(cudaStreamSynchronize(streams[2]));
    NGX_CUDA(cudaStreamSynchronize(streams[1]));
    if (!res_pinned) {
        int plo = nchunks == 1 ? 0 : first_chunk + (nchunks - 2) * chunk;
        memcpy(res + plo, h_stage + plo, (size_t)(n - plo) * sizeof(*res));
    }
    for (int b = 0; b < 2; b++) {
        cudaEventDestroy(up[b]);
        cudaEventDestroy(done[b]);
    }
    return NGX_OK;
}
#endif

#ifndef NGX_LM_WIDE_TU
int ngx_lm_fit_host(int n, int model, const double *images,
                    const double *weights, const ngx_jacobian *jac, int nrow,
                    int ncol, const double *psf_pars, int npsf,
                    const double *guess, const ngx_prior *prior,
                    const ngx_lm_pars *lm_pars, int ignore_zero_weight,
                    ngx_lm_result *res) {
    return lm_fit_host_impl(n, model, images, weights, 0, nullptr, jac, nrow,
                            ncol, psf_pars, npsf, guess, prior, lm_pars,
                            ignore_zero_weight, res);
}
#endif

#ifndef NGX_LM_WIDE_TU
int ngx_lm_fit_host_f32(int n, int model, const float *images,
                        const float *weights, const ngx_jacobian *jac,
                        int nrow, int ncol, const double *psf_pars, int npsf,
                        const double *guess, const ngx_prior *prior,
                        const ngx_lm_pars *lm_pars, int ignore_zero_weight,
                        ngx_lm_result *res) {
    return lm_fit_host_impl(n, model, images, weights, 1, nullptr, jac, nrow,
                            ncol, psf_pars, npsf, guess, prior, lm_pars,
                            ignore_zero_weight, res);
}
#endif

#ifndef NGX_LM_WIDE_TU
int ngx_expand_compact(int n, int nrow, int ncol, const float *images32,
                       const double *weight_scalar,
                       const unsigned char *maskbits, double *images,
                       double *weights, void *stream) {
    NGX_REQUIRE(n >= 0 && nrow >= 1 && ncol >= 1);
    NGX_REQUIRE(images32 && weight_scalar && images && weights);
    if (n == 0) return NGX_OK;
    const size_t ntot = (size_t)nrow * ncol;
    const size_t cnt = (size_t)n * ntot;
    int blocks = (int)((cnt + 255) / 256 < (size_t)(8 * sm_count())
                           ? (cnt + 255) / 256 : (size_t)(8 * sm_count()));
    k_expand_compact<<<blocks, 256, 0, as_stream(stream)>>>(
        images32, weight_scalar, maskbits, images, weights, n, (int)ntot,
        (int)((ntot + 7) / 8));
    NGX_LAUNCH_CHECK();
    return NGX_OK;
}

int ngx_lm_fit_host_compact(int n, int model, const float *images,
                            const double *weight_scalar,
                            const unsigned char *maskbits,
                            const ngx_jacobian *jac, int nrow, int ncol,
                            const double *psf_pars, int npsf,
                            const double *guess, const ngx_prior *prior,
                            const ngx_lm_pars *lm_pars,
                            int ignore_zero_weight, ngx_lm_result *res) {
    return lm_fit_host_impl(n, model, images, weight_scalar, 2, maskbits, jac,
                            nrow, ncol, psf_pars, npsf, guess, prior, lm_pars,
                            ignore_zero_weight, res);
}
#endif

}  // extern "C"
