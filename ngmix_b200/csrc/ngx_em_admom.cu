// ngx_em: EMFitter*.go (ngmix/em/em.py:228-317) = prep_image (:95-120) +
//         em_run / _fixcen / _fixcov / _fluxonly (em/em_nb.py:15-1257)
// ngx_admom: admom (ngmix/admom/admom_nb.py:13-243)
//
// One CTA per stamp, the whole iteration loop inside the kernel, the stamp
// resident in shared memory.
//
// EM is REFERENCE-ORDER by construction.  On noisy PSF stamps the EM
// iteration count is chaotic: letting a compiler contract the reference's
// own loops into FMAs changes numiter on 14 of 64 golden stamps by up to 240
// iterations (measured with the oracle).  Bit-exact iteration counts
// therefore need the reference's exact arithmetic: every per-pixel quantity
// is computed unfused in the reference's operation order, and the pixel sums
// are accumulated in pixel order - by 6*ngauss+2 threads, one per sum, each
// walking a 64-pixel tile of per-pixel contributions that all 64 threads
// produced in parallel.  Only the convergence scalar (log of p and det) can
// differ from the reference, by an ulp of libm's log.
//
// admom converges in ~8 well-conditioned iterations, so it uses the fast
// path (fused evaluation, shuffle-tree reductions).
#include "ngx_core.cuh"
#include "ngx_host.cuh"

namespace ngx {

#define NGX_EM_MAX_CONV 48
#define NGX_EM_TILE 64          // pixels per tile == threads per CTA
#define NGX_EM_MAX_SLOTS 64     // 6*ngauss + 2 <= 62
// contributions are stored slot-major with a padded, 16-byte aligned stride:
// phase B (thread s walks slot s with 16-byte loads) spreads over the banks
#define NGX_EM_STRIDE (NGX_EM_TILE + 2)

struct EMGauss {
    double p, row, col, irr, irc, icc;
};

struct EMShared {
    EMGauss gm[10];              // pre-psf mixture being fitted
    EMGauss psf[16];
    // normalised convolved mixture and its log terms: ngauss*npsf entries in
    // dynamic shared memory (a fixed 48-entry array cost two resident CTAs)
    ngx_gauss *conv;
    double *logtau, *logdet;
    double tab[16];
    double psf_irr, psf_irc, psf_icc, psf_row, psf_col, psf_ipsum;
    double sky, elogL_last, frac_diff, p_last, shift;
    double red[NGX_EM_MAX_SLOTS];      // the accumulated sums
    double mn[2], mx[2];
    ngx_jacobian jac;
    int npix, numiter, status, done, fill_zero;
    int cnt[2], nzero[2];
};

// gmix_convolve_fill + gmix_set_norms (em_nb.py:343-354) and
// set_logtau_logdet (:658-675); thread k < ngauss*npsf
__device__ __forceinline__ void em_build_conv(EMShared *sh, int ngauss,
                                              int npsf, bool want_log) {
    int k = threadIdx.x;
    if (k >= ngauss * npsf) return;
    int io = k / npsf, ip = k - io * npsf;
    const EMGauss &o = sh->gm[io];
    const EMGauss &q = sh->psf[ip];
    ngx_gauss g;
    gauss_set(g, o.p * q.p * sh->psf_ipsum, o.row + (q.row - sh->psf_row),
              o.col + (q.col - sh->psf_col), o.irr + q.irr, o.irc + q.irc,
              o.icc + q.icc);
    if (gauss_set_norm(g) != NGX_ST_OK) {
        atomicOr(&sh->status, NGX_ST_RANGE_ERROR);
        return;
    }
    sh->conv[k] = g;
    if (want_log) {
        sh->logtau[k] = log(g.p);
        sh->logdet[k] = log(g.det);
    }
}

// slot layout per object gaussian ii: 6*ii + {0 pnew, 1 vsum, 2 usum,
// 3 v2sum, 4 uvsum, 5 u2sum}; then 6*ngauss: elogL, 6*ngauss+1: skysum
// NG > 0 fixes ngauss at compile time: the per-gaussian sums of a pixel then
// stay in registers until the pixel's total is known and are written to
// shared memory once, already scaled (NG = 0: runtime ngauss, scaled in
// place).  Same operations in the same order either way.
// GM: the stamp-sized arrays (shifted image, pixel coordinates, pixel codes)
// live in a global scratch buffer instead of shared memory, for stamps too
// large to be resident; the arithmetic is the same.
template <int VARIANT, int NG, bool GM = false>
__global__ void __launch_bounds__(NGX_EM_TILE, GM ? 1 : 10)
k_em(int n, const double *__restrict__ images,
     const double *__restrict__ weights, const ngx_jacobian *__restrict__ jac,
     int nrow, int ncol, const double *__restrict__ guess, int ngauss_in,
     const double *__restrict__ psf_pars, int npsf_in, ngx_em_conf conf,
     int ignore_zero_weight, double *__restrict__ pars_out,
     ngx_em_result *__restrict__ res, double *gscratch = nullptr,
     const int32_t *__restrict__ todo = nullptr) {
    extern __shared__ double smem_d[];
    __shared__ EMShared sh;
    constexpr bool DO_CEN = (VARIANT == NGX_EM_FULL || VARIANT == NGX_EM_FIXCOV);
    constexpr bool DO_COV = (VARIANT == NGX_EM_FULL || VARIANT == NGX_EM_FIXCEN);
    constexpr bool DO_LOGL = (VARIANT != NGX_EM_FLUXONLY);
    const int T = NGX_EM_TILE;
    const int ntot = nrow * ncol;
    const int ngauss = NG > 0 ? NG : ngauss_in;
    const int npsf = npsf_in > 0 ? npsf_in : 1;
    const int nslots = 6 * ngauss + 2;
    const int nconv = ngauss * npsf;
    // nslots * NGX_EM_STRIDE first (16-byte aligned for the phase-B loads)
    double *contrib = smem_d;
    const int stamp = blockIdx.x;
    if (stamp >= n) return;
    // retry rounds of the fitting chain: only the stamps still to do
    if (todo && todo[stamp] == 0) return;
    double *stamp_arrays = contrib + (size_t)nslots * NGX_EM_STRIDE;
    if (GM)
        stamp_arrays = gscratch
            + (size_t)stamp * (3 * (size_t)ntot + ((size_t)ntot + 7) / 8);
    double *val = stamp_arrays;         // ntot
    // pixel coordinates: cached per stamp only in the global-scratch variant.
    // With the stamp in shared memory they are recomputed from (row, col)
    // with the reference's formula (jacobian_nb.py:4-16, ~10 instructions per
    // pixel and iteration): 16 bytes per pixel less shared memory let 10
    // instead of 7 CTAs share an SM, and the kernel is bound by the latency
    // of its dependent chains, not by instruction count.
    double *pv = val + ntot;            // ntot (GM only)
    double *pu = pv + ntot;             // ntot (GM only)
    // per pixel: 255 excluded, 0 zero weight (filled each iteration), 1 normal
    unsigned char *code_g = (unsigned char *)(pu + ntot);
    ngx_gauss *conv_d = (ngx_gauss *)(
        GM ? contrib + (size_t)nslots * NGX_EM_STRIDE : val + ntot);  // nconv
    double *logtau_d = (double *)(conv_d + nconv);            // nconv
    double *logdet_d = logtau_d + nconv;                      // nconv
    unsigned char *code = GM ? code_g
                             : (unsigned char *)(logdet_d + nconv);

    const double *im = images + (size_t)stamp * ntot;
    const double *wt = weights + (size_t)stamp * ntot;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // ---- stage the stamp; min/max for prep_image
    double mn = INFINITY, mx = -INFINITY;
    int cnt = 0, nzero = 0;
    for (int p = threadIdx.x; p < ntot; p += T) {
        double w = wt[p], v = im[p];
        bool valid = !(ignore_zero_weight && w <= 0.0);
        val[p] = v;
        code[p] = valid ? (w > 0.0 ? 1 : 0) : 255;
        if (valid) cnt++;
        if (valid && !(w > 0.0)) nzero++;
        mn = fmin(mn, v);       // prep_image uses the whole image
        mx = fmax(mx, v);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        nzero += __shfl_xor_sync(0xffffffffu, nzero, o);
    }
    if (lane == 0) {
        sh.mn[warp] = mn; sh.mx[warp] = mx;
        sh.cnt[warp] = cnt; sh.nzero[warp] = nzero;
    }
    load_exp_table(sh.tab);
    if (threadIdx.x == 0) {
        sh.jac = jac[stamp];
        sh.status = 0;
        sh.done = 0;
        sh.numiter = 0;
        sh.frac_diff = nan("");
        sh.elogL_last = -9999.9e9;
        sh.conv = conv_d;
        sh.logtau = logtau_d;
        sh.logdet = logdet_d;
    }
    if (threadIdx.x < ngauss) {
        const double *g = guess + ((size_t)stamp * ngauss + threadIdx.x) * 6;
        EMGauss e = {g[0], g[1], g[2], g[3], g[4], g[5]};
        sh.gm[threadIdx.x] = e;
    }
    if (threadIdx.x < npsf) {
        EMGauss e = {1.0, 0.0, 0.0, 0.0, 0.0, 0.0};    // delta psf, em.py:256
        if (npsf_in > 0) {
            const double *g = psf_pars
                + ((size_t)stamp * npsf + threadIdx.x) * 6;
            e.p = g[0]; e.row = g[1]; e.col = g[2];
            e.irr = g[3]; e.irc = g[4]; e.icc = g[5];
        }
        sh.psf[threadIdx.x] = e;
    }
    __syncthreads();
    // pixel coordinates, jacobian_nb.py:4-16 (they do not change during the
    // iterations)
    if (GM) {
        for (int p = threadIdx.x; p < ntot; p += T) {
            int row = p / ncol, col = p - row * ncol;
            double rowdiff = row - sh.jac.row0, coldiff = col - sh.jac.col0;
            pv[p] = sh.jac.dvdrow * rowdiff + sh.jac.dvdcol * coldiff;
            pu[p] = sh.jac.dudrow * rowdiff + sh.jac.dudcol * coldiff;
        }
    }
    // (row, col) of this thread's pixel in the first tile
    const int row_first = (int)threadIdx.x / ncol;
    const int col_first = (int)threadIdx.x - row_first * ncol;
    const ngx_jacobian jc = sh.jac;
    if (threadIdx.x == 0) {
        double a = fmin(sh.mn[0], sh.mn[1]), b = fmax(sh.mx[0], sh.mx[1]);
        sh.npix = sh.cnt[0] + sh.cnt[1];
        sh.fill_zero = (sh.nzero[0] + sh.nzero[1]) > 0;
        if (conf.have_sky) {
            sh.sky = conf.sky;
            sh.shift = 0.0;
        } else {
            // prep_image, em.py:109-120
            double desired_minval = 0.001 * (b - a);
            sh.sky = desired_minval - a;
            sh.shift = sh.sky;
        }
        // psf.set_flux(1.0) (gmix.py:299-311) when a psf was given
        if (npsf_in > 0) {
            double psum0 = 0.0;
            for (int k = 0; k < npsf; k++) psum0 += sh.psf[k].p;
            double rat = 1.0 / psum0;
            for (int k = 0; k < npsf; k++) sh.psf[k].p *= rat;
        }
        // gmix_get_cen / gmix_get_moms of the psf, em_nb.py:1260-1294
        double row = 0.0, col = 0.0, psum = 0.0;
        for (int k = 0; k < npsf; k++) {
            double p = sh.psf[k].p;
            row += p * sh.psf[k].row;
            col += p * sh.psf[k].col;
            psum += p;
        }
        row /= psum;
        col /= psum;
        double irr = 0.0, irc = 0.0, icc = 0.0;
        for (int k = 0; k < npsf; k++) {
            double rd = sh.psf[k].row - row, cd = sh.psf[k].col - col;
            double p = sh.psf[k].p;
            irr += p * (sh.psf[k].irr + rd * rd);
            irc += p * (sh.psf[k].irc + rd * cd);
            icc += p * (sh.psf[k].icc + cd * cd);
        }
        sh.psf_row = row; sh.psf_col = col; sh.psf_ipsum = 1.0 / psum;
        sh.psf_irr = irr / psum; sh.psf_irc = irc / psum;
        sh.psf_icc = icc / psum;
        double pl = 0.0;
        for (int k = 0; k < ngauss; k++) pl += sh.gm[k].p;
        sh.p_last = pl;
        if (sh.npix == 0) sh.status = NGX_ST_FATAL;
    }
    __syncthreads();
    if (!conf.have_sky) {
        double s = sh.shift;
        for (int p = threadIdx.x; p < ntot; p += T) val[p] += s;
    }
    const double area = sh.jac.scale * sh.jac.scale;
    em_build_conv(&sh, ngauss, npsf, DO_LOGL);
    __syncthreads();

    // ---- iterations
    for (int it = 0; it < conf.maxiter && !sh.status && !sh.done; it++) {
        const double sky = sh.sky;
        if (sh.fill_zero) {
            // fill_zero_weight_pixels, em_nb.py:1297-1315 (apodized eval)
            for (int p = threadIdx.x; p < ntot; p += T) {
                if (code[p] == 0) {
                    double v, u;
                    if (GM) {
                        v = pv[p]; u = pu[p];
                    } else {
                        int row = p / ncol, col = p - row * ncol;
                        double rowdiff = row - jc.row0, coldiff = col - jc.col0;
                        v = jc.dvdrow * rowdiff + jc.dvdcol * coldiff;
                        u = jc.dudrow * rowdiff + jc.dudcol * coldiff;
                    }
                    double s = 0.0;
                    for (int k = 0; k < ngauss * npsf; k++) {
                        const ngx_gauss &g = sh.conv[k];
                        double vdiff = v - g.row, udiff = u - g.col;
                        double chi2 = g.dcc * vdiff * vdiff
                            + g.drr * udiff * udiff
                            - 2.0 * g.drc * vdiff * udiff;
                        double mv = 0.0;
                        if (chi2 < 25.0 && chi2 >= 0.0) {
                            double x = -0.5 * chi2;
                            mv = g.pnorm * fexp_ref_tab(x, sh.tab) * area;
                            if (chi2 > 20.0) mv *= apod_window_x(x);
                        }
                        s += mv;
                    }
                    val[p] = sky + s;
                }
            }
        }
        if (threadIdx.x < nslots) sh.red[threadIdx.x] = 0.0;
        double acc = 0.0;          // thread s accumulates slot s
        __syncthreads();

        int prow = row_first, pcol = col_first;
        for (int base = 0; base < ntot; base += T) {
            // ---- phase A: one pixel per thread, reference operation order
            int p = base + threadIdx.x;
            bool valid = (p < ntot) && (code[p] != 255);
            int bad = 0;
            double *cme = contrib + threadIdx.x;
            double v, u;
            if (GM) {
                v = p < ntot ? pv[p] : 0.0;
                u = p < ntot ? pu[p] : 0.0;
            } else {
                double rowdiff = prow - jc.row0, coldiff = pcol - jc.col0;
                v = jc.dvdrow * rowdiff + jc.dvdcol * coldiff;
                u = jc.dudrow * rowdiff + jc.dudcol * coldiff;
                pcol += T;
                while (pcol >= ncol) { pcol -= ncol; prow++; }
            }
            if (valid) {
                double gsum = 0.0, logL = 0.0;
                double cg[NG > 0 ? NG : 1][6];
#pragma unroll
                for (int ii = 0; ii < (NG > 0 ? NG : ngauss); ii++) {
                    double gi = 0.0, tv = 0.0, tu = 0.0, tv2 = 0.0, tuv = 0.0,
                           tu2 = 0.0;
                    for (int k = ii * npsf; k < (ii + 1) * npsf; k++) {
                        const ngx_gauss &g = sh.conv[k];
                        double vdiff = v - g.row;
                        double udiff = u - g.col;
                        double u2 = udiff * udiff;
                        double v2 = vdiff * vdiff;
                        double uv = udiff * vdiff;
                        double chi2 = g.dcc * v2 + g.drr * u2
                            - 2.0 * g.drc * uv;
                        double vl;
                        // hard cut, NO apodization: em_nb.py:222-227
                        if (chi2 < 25.0 && chi2 >= 0.0)
                            vl = g.pnorm * fexp_ref_tab(-0.5 * chi2, sh.tab) * area;
                        else
                            vl = 0.0;
                        gi += vl;
                        gsum += vl;
                        if (DO_CEN) {
                            tv += v * vl;
                            tu += u * vl;
                        }
                        if (DO_COV) {
                            tv2 += v2 * vl;
                            tuv += uv * vl;
                            tu2 += u2 * vl;
                        }
                        if (DO_LOGL)
                            logL += vl * (sh.logtau[k] - 0.5 * sh.logdet[k]
                                          - 0.5 * chi2);
                    }
                    if (NG > 0) {
                        cg[ii][0] = gi; cg[ii][1] = tv; cg[ii][2] = tu;
                        cg[ii][3] = tv2; cg[ii][4] = tuv; cg[ii][5] = tu2;
                    } else {
                        double *c = cme + (size_t)(6 * ii) * NGX_EM_STRIDE;
                        c[0] = gi; c[NGX_EM_STRIDE] = tv;
                        c[2 * NGX_EM_STRIDE] = tu; c[3 * NGX_EM_STRIDE] = tv2;
                        c[4 * NGX_EM_STRIDE] = tuv; c[5 * NGX_EM_STRIDE] = tu2;
                    }
                }
                if (DO_LOGL) {
                    if (gsum == 0.0) logL = 0.0;
                    else logL *= 1.0 / gsum;
                }
                double gtot = gsum + sky;
                if (gtot == 0.0) {
                    bad = 1;
                } else {
                    double pval = val[p];
                    cme[(size_t)(6 * ngauss) * NGX_EM_STRIDE] = logL;
                    cme[(size_t)(6 * ngauss + 1) * NGX_EM_STRIDE]
                        = sky * pval / gtot;
                    double factor = pval / gtot;
                    if (NG > 0) {
#pragma unroll
                        for (int ii = 0; ii < (NG > 0 ? NG : 1); ii++)
#pragma unroll
                            for (int m = 0; m < 6; m++)
                                cme[(size_t)(6 * ii + m) * NGX_EM_STRIDE]
                                    = cg[ii][m] * factor;
                    } else {
                        for (int s = 0; s < 6 * ngauss; s++)
                            cme[(size_t)s * NGX_EM_STRIDE] *= factor;
                    }
                }
            } else {
                // a pixel that does not take part adds exact zeros (x + 0 = x)
                for (int s = 0; s < nslots; s++)
                    cme[(size_t)s * NGX_EM_STRIDE] = 0.0;
            }
            if (bad) atomicOr(&sh.status, NGX_ST_RANGE_ERROR);   // gtot == 0
            __syncthreads();
            if (sh.status) break;
            // ---- phase B: thread s adds its slot in pixel order
            if (threadIdx.x < nslots) {
                const double2 *c = (const double2 *)(
                    contrib + (size_t)threadIdx.x * NGX_EM_STRIDE);
#pragma unroll 8
                for (int t = 0; t < NGX_EM_TILE / 2; t++) {
                    double2 x = c[t];
                    acc += x.x;
                    acc += x.y;
                }
            }
            __syncthreads();
        }
        if (sh.status) break;       // uniform
        if (threadIdx.x < nslots) sh.red[threadIdx.x] = acc;
        __syncthreads();

        // ---- M step: gmix_set_from_sums*, thread i < ngauss
        if (threadIdx.x < ngauss) {
            int i = threadIdx.x;
            EMGauss g = sh.gm[i];
            double p = sh.red[6 * i];
            if (VARIANT != NGX_EM_FLUXONLY && p == 0.0) {
                // numba raises ZeroDivisionError on pinv = 1.0/p
                atomicOr(&sh.status, NGX_ST_ZERO_DIV);
            } else {
                double pinv = (VARIANT != NGX_EM_FLUXONLY) ? 1.0 / p : 0.0;
                if (DO_CEN) {
                    g.row = sh.red[6 * i + 1] * pinv;
                    g.col = sh.red[6 * i + 2] * pinv;
                }
                if (DO_COV) {
                    double irr = sh.red[6 * i + 3] * pinv;
                    double irc = sh.red[6 * i + 4] * pinv;
                    double icc = sh.red[6 * i + 5] * pinv;
                    irr = irr - sh.psf_irr;
                    irc = irc - sh.psf_irc;
                    icc = icc - sh.psf_icc;
                    if (irr < 0.0 || icc < 0.0) {
                        irr = 1.0e-4; irc = 0.0; icc = 1.0e-4;
                    }
                    double det = irr * icc - irc * irc;
                    if (det < NGX_LOW_DETVAL) {
                        double Tt = irr + icc;
                        irr = icc = Tt / 2;
                        irc = 0.0;
                    }
                    g.irr = irr; g.irc = irc; g.icc = icc;
                }
                g.p = p;
                sh.gm[i] = g;
            }
        }
        __syncthreads();
        if (!sh.status) em_build_conv(&sh, ngauss, npsf, DO_LOGL);
        if (threadIdx.x == 0 && !sh.status) {
            if (conf.vary_sky) sh.sky = sh.red[6 * ngauss + 1] / sh.npix;
            int numiter = it + 1;
            sh.numiter = numiter;
            if (VARIANT == NGX_EM_FLUXONLY) {
                double psum = 0.0;
                for (int k = 0; k < ngauss; k++) psum += sh.gm[k].p;
                if (numiter >= conf.miniter) {
                    sh.frac_diff = fabs(psum / sh.p_last - 1);
                    if (sh.frac_diff < conf.tol) sh.done = 1;
                }
                sh.p_last = psum;
            } else {
                double elogL = sh.red[6 * ngauss];
                if (numiter >= conf.miniter) {
                    if (elogL == 0.0) {
                        atomicOr(&sh.status, NGX_ST_RANGE_ERROR);
                    } else {
                        sh.frac_diff = fabs((elogL - sh.elogL_last) / elogL);
                        if (sh.frac_diff < conf.tol) sh.done = 1;
                    }
                }
                sh.elogL_last = elogL;
            }
        }
        __syncthreads();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        ngx_em_result r;
        r.flags = 0;
        r.numiter = sh.numiter;
        r.fdiff = sh.frac_diff;
        r.sky = sh.sky;
        if (sh.status) r.flags = NGX_EM_RANGE_ERROR;
        else if (sh.numiter >= conf.maxiter) r.flags = NGX_EM_MAXITER;
        res[stamp] = r;
    }
    if (threadIdx.x < ngauss) {
        double *o = pars_out + ((size_t)stamp * ngauss + threadIdx.x) * 6;
        EMGauss g = sh.gm[threadIdx.x];
        o[0] = g.p; o[1] = g.row; o[2] = g.col;
        o[3] = g.irr; o[4] = g.irc; o[5] = g.icc;
    }
}

// =====================================================================
// adaptive moments
struct AMShared {
    ngx_gauss wt;
    GaussV wv;
    double tab[16];
    double red[40];
    ngx_jacobian jac;
    double roworig, colorig, e1old, e2old, Told;
    int flags, done, numiter, npix_valid;
};

// GM: pixel values and variances in a global scratch buffer (stamps too large
// for shared memory)
template <bool GM>
__global__ void __launch_bounds__(128, 4)
k_admom(int n, const double *__restrict__ images,
        const double *__restrict__ weights,
        const ngx_jacobian *__restrict__ jac, int nrow, int ncol,
        const double *__restrict__ guess, ngx_admom_conf conf,
        ngx_admom_result *__restrict__ res, double *gscratch) {
    extern __shared__ double smem_d[];
    __shared__ AMShared sh;
    __shared__ ngx_admom_result out;
    const int T = blockDim.x;
    const int ntot = nrow * ncol;
    const int stamp = blockIdx.x;
    if (stamp >= n) return;
    double *val = GM ? gscratch + (size_t)stamp * 2 * ntot : smem_d;  // ntot
    double *var = val + ntot;        // ntot: 1/ierr^2, < 0 for dropped pixels
    double *rowd = GM ? smem_d : var + ntot;
    double *cold = rowd + nrow;
    double *scratch = cold + ncol;   // 37 * nwarps
    const double *im = images + (size_t)stamp * ntot;
    const double *wt = weights + (size_t)stamp * ntot;
    int cnt = 0;
    for (int p = threadIdx.x; p < ntot; p += T) {
        double w = wt[p];
        val[p] = im[p];
        if (w > 0.0) {
            // var = 1/(ierr*ierr), ierr = sqrt(w): admom_nb.py:147
            double ie = sqrt(w);
            var[p] = 1.0 / (ie * ie);
            cnt++;
        } else {
            var[p] = -1.0;           // dropped (Observation default)
        }
    }
    (void)cnt;
    load_exp_table(sh.tab);
    if (threadIdx.x == 0) {
        sh.jac = jac[stamp];
        const double *g = guess + (size_t)stamp * 6;
        gauss_set(sh.wt, g[0], g[1], g[2], g[3], g[4], g[5]);
        sh.roworig = g[1];
        sh.colorig = g[2];
        sh.e1old = sh.e2old = sh.Told = nan("");
        sh.flags = 0;
        sh.done = 0;
        sh.numiter = 0;
        memset(&out, 0, sizeof(out));
    }
    __syncthreads();
    for (int p = threadIdx.x; p < nrow; p += T) rowd[p] = p - sh.jac.row0;
    for (int p = threadIdx.x; p < ncol; p += T) cold[p] = p - sh.jac.col0;
    const double area = sh.jac.scale * sh.jac.scale;
    const double dvdrow = sh.jac.dvdrow, dvdcol = sh.jac.dvdcol;
    const double dudrow = sh.jac.dudrow, dudcol = sh.jac.dudcol;
    __syncthreads();

    int i = 0;
    for (i = 0; i < conf.maxiter; i++) {
        if (threadIdx.x == 0) {
            if (sh.wt.det < NGX_LOW_DETVAL) {
                sh.flags = NGX_LOW_DET;
                sh.done = 1;
            } else {
                gauss_set_norm(sh.wt);
                sh.wv = make_gaussv(sh.wt, area);
            }
        }
        __syncthreads();
        if (sh.done) break;
        // ---- admom_censums :111-128
        double cs[4] = {0.0, 0.0, 0.0, 0.0};
        {
            GaussV g = sh.wv;
            for (int p = threadIdx.x; p < ntot; p += T) {
                if (var[p] < 0.0) continue;
                int row = p / ncol, col = p - row * ncol;
                double v = dvdrow * rowd[row] + dvdcol * cold[col];
                double u = dudrow * rowd[row] + dudcol * cold[col];
                double weight = eval_fast(g, v, u, sh.tab);
                double wdata = weight * val[p];
                cs[0] = fma(wdata, v, cs[0]);
                cs[1] = fma(wdata, u, cs[1]);
                cs[2] += wdata;
                cs[3] += 1.0;
            }
        }
        block_sum_to<4>(cs, scratch, sh.red);
        if (threadIdx.x == 0) {
            cs[0] = sh.red[0]; cs[1] = sh.red[1]; cs[2] = sh.red[2];
            cs[3] = sh.red[3];
            // clear_result + censums bookkeeping
            out.npix = (int)cs[3];
            out.wsum = 0.0;
            for (int k = 0; k < 7; k++) out.sums[k] = 0.0;
            for (int k = 0; k < 49; k++) out.sums_cov[k] = 0.0;
            for (int k = 0; k < 6; k++) out.pars[k] = nan("");
            out.rho4 = nan("");
            out.sums[0] = cs[0]; out.sums[1] = cs[1]; out.sums[5] = cs[2];
            if (cs[2] <= 0.0) {
                sh.flags = NGX_NONPOS_FLUX;
                sh.done = 1;
            } else {
                sh.wt.row = cs[0] / cs[2];
                sh.wt.col = cs[1] / cs[2];
                if (fabs(sh.wt.row - sh.roworig) > conf.shiftmax
                    || fabs(sh.wt.col - sh.colorig) > conf.shiftmax) {
                    sh.flags = NGX_CEN_SHIFT;
                    sh.done = 1;
                } else {
                    sh.wv.row = sh.wt.row;
                    sh.wv.col = sh.wt.col;
                }
            }
        }
        __syncthreads();
        if (sh.done) break;
        // ---- admom_momsums :131-175: 7 sums, 28 unique cov sums, wsum, n
        double ms[37];
#pragma unroll
        for (int k = 0; k < 37; k++) ms[k] = 0.0;
        {
            GaussV g = sh.wv;
            double vcen = g.row, ucen = g.col;
            double dcc = -2.0 * g.a, drr = -2.0 * g.b, drc = g.c;
            for (int p = threadIdx.x; p < ntot; p += T) {
                double vr = var[p];
                if (vr < 0.0) continue;
                int row = p / ncol, col = p - row * ncol;
                double v = dvdrow * rowd[row] + dvdcol * cold[col];
                double u = dudrow * rowd[row] + dudcol * cold[col];
                double weight = eval_fast(g, v, u, sh.tab);
                double vmod = v - vcen, umod = u - ucen;
                double wdata = weight * val[p];
                double w2 = weight * weight;
                double chi2 = dcc * vmod * vmod + drr * umod * umod
                    - 2.0 * drc * vmod * umod;
                double F[7];
                F[0] = v;
                F[1] = u;
                F[2] = umod * umod - vmod * vmod;
                F[3] = 2 * vmod * umod;
                F[4] = umod * umod + vmod * vmod;
                F[5] = 1.0;
                F[6] = chi2 * chi2;
                double w2v = w2 * vr;
                int q = 7;
#pragma unroll
                for (int a = 0; a < 7; a++) {
                    ms[a] = fma(wdata, F[a], ms[a]);
                    double t = w2v * F[a];
#pragma unroll
                    for (int b = 0; b <= a; b++) {
                        ms[q] = fma(t, F[b], ms[q]);
                        q++;
                    }
                }
                ms[35] += weight;
                ms[36] += 1.0;
            }
        }
        // halving-butterfly reductions, one value per lane (32 + 5 sums): the
        // unrolled 37 x 5 shuffle trees were most of the kernel's code
        {
            double hi[5] = {ms[32], ms[33], ms[34], ms[35], ms[36]};
            block_sum_to<32>(*(double (*)[32])ms, scratch, sh.red);
            block_sum_to<5>(hi, scratch, sh.red + 32);
        }
        if (threadIdx.x == 0) {
            const double *R = sh.red;     // the 37 block totals
            out.npix = (int)R[36];
            out.wsum = R[35];
            for (int k = 0; k < 7; k++) out.sums[k] = R[k];
            int q = 7;
            for (int a = 0; a < 7; a++)
                for (int b = 0; b <= a; b++) {
                    out.sums_cov[a * 7 + b] = R[q];
                    out.sums_cov[b * 7 + a] = R[q];
                    q++;
                }
            for (int k = 0; k < 6; k++) out.pars[k] = nan("");
            out.rho4 = nan("");
            if (R[5] <= 0.0) {
                sh.flags = NGX_NONPOS_FLUX;
                sh.done = 1;
            } else {
                double finv = 1.0 / R[5];
                double M1 = R[2] * finv, M2 = R[3] * finv;
                double Tm = R[4] * finv;
                double Irr = 0.5 * (Tm - M1), Icc = 0.5 * (Tm + M1);
                double Irc = 0.5 * M2;
                if (Tm <= 0.0) {
                    sh.flags = NGX_NONPOS_SIZE;
                    sh.done = 1;
                } else {
                    double e1 = (Icc - Irr) / Tm, e2 = 2 * Irc / Tm;
                    if (fabs(e1 - sh.e1old) < conf.etol
                        && fabs(e2 - sh.e2old) < conf.etol
                        && fabs(Tm / sh.Told - 1.) < conf.Ttol) {
                        out.pars[0] = sh.wt.row;
                        out.pars[1] = sh.wt.col;
                        out.pars[2] = sh.wt.icc - sh.wt.irr;
                        out.pars[3] = 2.0 * sh.wt.irc;
                        out.pars[4] = sh.wt.icc + sh.wt.irr;
                        out.pars[5] = 1.0;
                        out.rho4 = R[6] / R[5];
                        sh.done = 1;
                    } else {
                        if (!conf.cenonly) {
                            // deweight_moments :178-226
                            double detm = Irr * Icc - Irc * Irc;
                            double Wrr = sh.wt.irr, Wrc = sh.wt.irc,
                                   Wcc = sh.wt.icc;
                            double detw = Wrr * Wcc - Wrc * Wrc;
                            if (detm <= NGX_LOW_DETVAL
                                || detw <= NGX_LOW_DETVAL) {
                                sh.flags = NGX_LOW_DET;
                                sh.done = 1;
                            } else {
                                double idetw = 1.0 / detw, idetm = 1.0 / detm;
                                double Nrr = Icc * idetm - Wcc * idetw;
                                double Ncc = Irr * idetm - Wrr * idetw;
                                double Nrc = -Irc * idetm + Wrc * idetw;
                                double detn = Nrr * Ncc - Nrc * Nrc;
                                if (detn <= NGX_LOW_DETVAL) {
                                    sh.flags = NGX_LOW_DET;
                                    sh.done = 1;
                                } else {
                                    double idetn = 1. / detn;
                                    sh.wt.irr = Ncc * idetn;
                                    sh.wt.icc = Nrr * idetn;
                                    sh.wt.irc = -Nrc * idetn;
                                    sh.wt.det = sh.wt.irr * sh.wt.icc
                                        - sh.wt.irc * sh.wt.irc;
                                }
                            }
                        }
                        sh.e1old = e1; sh.e2old = e2; sh.Told = Tm;
                    }
                }
            }
        }
        __syncthreads();
        if (sh.done) break;
    }
    if (threadIdx.x == 0) {
        // python leaves i at maxiter-1 when the loop is exhausted
        if (i == conf.maxiter) i = conf.maxiter - 1;
        out.numiter = i + 1;
        out.flags = sh.flags;
        if (out.numiter == conf.maxiter) out.flags = NGX_MAXITER;
        out.wt_norm = sh.wt.norm;
        res[stamp] = out;
    }
}

static int pick_threads(int ntot) {
    if (ntot <= 400) return 32;
    if (ntot <= 1600) return 64;
    return 128;
}

// stamps to fit in the current call (ngx_em_todo); NULL = all
static thread_local const int32_t *g_em_todo = nullptr;

template <int VARIANT>
static int launch_em(int n, const double *images, const double *weights,
                     const ngx_jacobian *jac, int nrow, int ncol,
                     const double *guess, int ngauss, const double *psf_pars,
                     int npsf, const ngx_em_conf &conf, int ignore_zero_weight,
                     double *pars_out, ngx_em_result *res,
                     cudaStream_t stream) {
    int ntot = nrow * ncol;
    int nslots = 6 * ngauss + 2;
    int nconv = ngauss * (npsf > 0 ? npsf : 1);
    size_t smem = sizeof(double)
        * ((size_t)ntot + (size_t)nslots * NGX_EM_STRIDE + 2 * nconv)
        + sizeof(ngx_gauss) * (size_t)nconv + (size_t)ntot + 16;
    if (smem > (size_t)smem_optin_bytes()) {
        // stamp-sized arrays in a stream-ordered global scratch buffer,
        // a bounded number of stamps per launch
        size_t smem_g = sizeof(double)
            * ((size_t)nslots * NGX_EM_STRIDE + 2 * nconv)
            + sizeof(ngx_gauss) * (size_t)nconv + 16;
        if (smem_g > (size_t)smem_optin_bytes())
            return fail(NGX_ERR_ARG, "ngx_em: mixture too large%s%s");
        const size_t per = 3 * (size_t)ntot + ((size_t)ntot + 7) / 8;
        int chunk = (int)((size_t)(1u << 28) / per);       // 2 GiB of doubles
        if (chunk < 1) chunk = 1;
        if (chunk > n) chunk = n;
        double *scr = nullptr;
        NGX_CUDA(cudaMallocAsync((void **)&scr, sizeof(double) * per * chunk,
                                 stream));
        NGX_CUDA(cudaFuncSetAttribute(
            k_em<VARIANT, 0, true>,
            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
        const int np1 = npsf > 0 ? npsf : 0;
        for (int i0 = 0; i0 < n; i0 += chunk) {
            const int m = n - i0 < chunk ? n - i0 : chunk;
            k_em<VARIANT, 0, true><<<m, NGX_EM_TILE, smem_g, stream>>>(
                m, images + (size_t)i0 * ntot, weights + (size_t)i0 * ntot,
                jac + i0, nrow, ncol, guess + (size_t)i0 * ngauss * 6, ngauss,
                psf_pars ? psf_pars + (size_t)i0 * np1 * 6 : nullptr, npsf,
                conf, ignore_zero_weight, pars_out + (size_t)i0 * ngauss * 6,
                res + i0, scr, g_em_todo ? g_em_todo + i0 : nullptr);
        }
        cudaError_t e = cudaGetLastError();
        cudaFreeAsync(scr, stream);
        NGX_CUDA(e);
        return NGX_OK;
    }
#define NGX_EM_LAUNCH(NG_)                                                   \
    do {                                                                     \
        NGX_CUDA(cudaFuncSetAttribute(                                       \
            k_em<VARIANT, NG_>,                                              \
            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
        k_em<VARIANT, NG_><<<n, NGX_EM_TILE, smem, stream>>>(                \
            n, images, weights, jac, nrow, ncol, guess, ngauss, psf_pars,    \
            npsf, conf, ignore_zero_weight, pars_out, res, nullptr,          \
            g_em_todo);                                                      \
    } while (0)
    // the common mixture sizes get the register-resident variant
    if (ngauss == 1) NGX_EM_LAUNCH(1);
    else if (ngauss == 2) NGX_EM_LAUNCH(2);
    else if (ngauss == 3) NGX_EM_LAUNCH(3);
    else NGX_EM_LAUNCH(0);
#undef NGX_EM_LAUNCH
    NGX_LAUNCH_CHECK();
    return NGX_OK;
}

}  // namespace ngx

using namespace ngx;

extern "C" {

int ngx_em(int n, int variant, const double *images, const double *weights,
           const ngx_jacobian *jac, int nrow, int ncol, const double *guess,
           int ngauss, const double *psf_pars, int npsf,
           const ngx_em_conf *conf, int ignore_zero_weight, double *pars_out,
           ngx_em_result *res, void *stream) {
    NGX_REQUIRE(n >= 0 && images && weights && jac && guess && conf);
    NGX_REQUIRE(pars_out && res && nrow >= 1 && ncol >= 1);
    NGX_REQUIRE(ngauss >= 1 && ngauss <= 10 && npsf >= 0 && npsf <= 16);
    NGX_REQUIRE(npsf == 0 || psf_pars);
    NGX_REQUIRE(ngauss * (npsf > 0 ? npsf : 1) <= NGX_EM_MAX_CONV);
    if (n == 0) return NGX_OK;
    cudaStream_t s = as_stream(stream);
    switch (variant) {
    case NGX_EM_FULL:
        return launch_em<NGX_EM_FULL>(n, images, weights, jac, nrow, ncol,
                                      guess, ngauss, psf_pars, npsf, *conf,
                                      ignore_zero_weight, pars_out, res, s);
    case NGX_EM_FIXCEN:
        return launch_em<NGX_EM_FIXCEN>(n, images, weights, jac, nrow, ncol,
                                        guess, ngauss, psf_pars, npsf, *conf,
                                        ignore_zero_weight, pars_out, res, s);
    case NGX_EM_FIXCOV:
        return launch_em<NGX_EM_FIXCOV>(n, images, weights, jac, nrow, ncol,
                                        guess, ngauss, psf_pars, npsf, *conf,
                                        ignore_zero_weight, pars_out, res, s);
    case NGX_EM_FLUXONLY:
        return launch_em<NGX_EM_FLUXONLY>(n, images, weights, jac, nrow, ncol,
                                          guess, ngauss, psf_pars, npsf,
                                          *conf, ignore_zero_weight, pars_out,
                                          res, s);
    default:
        return fail(NGX_ERR_ARG, "ngx_em: unknown variant%s%s");
    }
}

int ngx_em_todo(int n, int variant, const double *images,
                const double *weights, const ngx_jacobian *jac, int nrow,
                int ncol, const double *guess, int ngauss,
                const double *psf_pars, int npsf, const ngx_em_conf *conf,
                int ignore_zero_weight, const int32_t *todo, double *pars_out,
                ngx_em_result *res, void *stream) {
    g_em_todo = todo;
    int rc = ngx_em(n, variant, images, weights, jac, nrow, ncol, guess,
                    ngauss, psf_pars, npsf, conf, ignore_zero_weight,
                    pars_out, res, stream);
    g_em_todo = nullptr;
    return rc;
}

int ngx_admom(int n, const double *images, const double *weights,
              const ngx_jacobian *jac, int nrow, int ncol, const double *guess,
              const ngx_admom_conf *conf, ngx_admom_result *res,
              void *stream) {
    NGX_REQUIRE(n >= 0 && images && weights && jac && guess && conf && res);
    NGX_REQUIRE(nrow >= 1 && ncol >= 1 && conf->maxiter >= 1);
    if (n == 0) return NGX_OK;
    int ntot = nrow * ncol;
    int threads = pick_threads(ntot);
    int nwarps = threads / 32;
    size_t smem = sizeof(double)
        * (2 * (size_t)ntot + nrow + ncol + 37 * (size_t)nwarps);
    cudaStream_t st = as_stream(stream);
    if (smem > (size_t)smem_optin_bytes()) {
        // pixel values and variances in a stream-ordered scratch buffer
        size_t smem_g = sizeof(double) * (nrow + ncol + 37 * (size_t)nwarps);
        const size_t per = 2 * (size_t)ntot;
        int chunk = (int)((size_t)(1u << 28) / per);       // 2 GiB of doubles
        if (chunk < 1) chunk = 1;
        if (chunk > n) chunk = n;
        double *scr = nullptr;
        NGX_CUDA(cudaMallocAsync((void **)&scr, sizeof(double) * per * chunk,
                                 st));
        for (int i0 = 0; i0 < n; i0 += chunk) {
            const int m = n - i0 < chunk ? n - i0 : chunk;
            k_admom<true><<<m, threads, smem_g, st>>>(
                m, images + (size_t)i0 * ntot, weights + (size_t)i0 * ntot,
                jac + i0, nrow, ncol, guess + (size_t)i0 * 6, *conf, res + i0,
                scr);
        }
        cudaError_t e = cudaGetLastError();
        cudaFreeAsync(scr, st);
        NGX_CUDA(e);
        return NGX_OK;
    }
    NGX_CUDA(cudaFuncSetAttribute(
        k_admom<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
        (int)smem));
    k_admom<false><<<n, threads, smem, st>>>(
        n, images, weights, jac, nrow, ncol, guess, *conf, res, nullptr);
    NGX_LAUNCH_CHECK();
    return NGX_OK;
}

int ngx_em_host(int n, int variant, const double *images,
                const double *weights, const ngx_jacobian *jac, int nrow,
                int ncol, const double *guess, int ngauss,
                const double *psf_pars, int npsf, const ngx_em_conf *conf,
                int ignore_zero_weight, double *pars_out,
                ngx_em_result *res) {
    NGX_REQUIRE(n >= 0 && images && weights && jac && guess && conf);
    if (n == 0) return NGX_OK;
    DevBuf<ngx_jacobian> dj(n, jac);
    DevBuf<double> dg((size_t)n * ngauss * 6, guess);
    DevBuf<double> dp(npsf > 0 ? (size_t)n * npsf * 6 : 0, psf_pars);
    DevBuf<double> dout((size_t)n * ngauss * 6);
    DevBuf<ngx_em_result> dr(n);
    NGX_CUDA(dj.err); NGX_CUDA(dg.err);
    NGX_CUDA(dp.err); NGX_CUDA(dout.err); NGX_CUDA(dr.err);
    NGX_CUDA(dout.zero());
    int rc = planes_pipeline(
        n, (size_t)nrow * ncol, images, weights,
        [&](int lo, int cn, const double *di, const double *dw,
            cudaStream_t q) {
            return ngx_em(cn, variant, di, dw, dj.d + lo, nrow, ncol,
                          dg.d + (size_t)lo * ngauss * 6, ngauss,
                          npsf > 0 ? dp.d + (size_t)lo * npsf * 6 : nullptr,
                          npsf, conf, ignore_zero_weight,
                          dout.d + (size_t)lo * ngauss * 6, dr.d + lo, q);
        });
    if (rc) return rc;
    NGX_CUDA(dout.to_host(pars_out));
    NGX_CUDA(dr.to_host(res));
    return NGX_OK;
}

int ngx_admom_host(int n, const double *images, const double *weights,
                   const ngx_jacobian *jac, int nrow, int ncol,
                   const double *guess, const ngx_admom_conf *conf,
                   ngx_admom_result *res) {
    NGX_REQUIRE(n >= 0 && images && weights && jac && guess && conf && res);
    if (n == 0) return NGX_OK;
    DevBuf<ngx_jacobian> dj(n, jac);
    DevBuf<double> dg((size_t)n * 6, guess);
    DevBuf<ngx_admom_result> dr(n);
    NGX_CUDA(dj.err); NGX_CUDA(dg.err); NGX_CUDA(dr.err);
    int rc = planes_pipeline(
        n, (size_t)nrow * ncol, images, weights,
        [&](int lo, int cn, const double *di, const double *dw,
            cudaStream_t q) {
            return ngx_admom(cn, di, dw, dj.d + lo, nrow, ncol,
                             dg.d + (size_t)lo * 6, conf, dr.d + lo, q);
        });
    if (rc) return rc;
    NGX_CUDA(dr.to_host(res));
    return NGX_OK;
}

}  // extern "C"
