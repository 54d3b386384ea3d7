// Prior rows of the LM residual vector and their forward-difference
// Jacobian, evaluated by one thread per object:
//   PriorSimpleSep.fill_fdiff      ngmix/joint_prior.py:86-120
//   PriorBDFSep.fill_fdiff         ngmix/joint_prior.py:565-599
//   PriorBDSep.fill_fdiff          ngmix/joint_prior.py:364-399
//   get_lnprob_scalar              ngmix/joint_prior.py:136-150, 543-563
//   FitModel._fill_prior_jacobian  ngmix/fitting/results.py:572-625
// over CenPrior (priors/multivariate.py:48-71), GPriorBA (priors/shape.py:
// 520-564), FlatPrior / TwoSidedErf / Normal (priors/priors.py:85-157,
// 219-337, 424-482).
#pragma once
#include "ngx_core.cuh"

namespace ngx {

#define NGX_STEP_PRIOR 1.0e-8       // results.py:935
#define NGX_MAX_PRIOR_ROWS LM_MAXP

static __device__ __noinline__ int prior1d_lnprob(const ngx_prior1d &p, double val,
                                     double *lnp) {
    switch (p.kind) {
    case NGX_P_FLAT:
        if (val < p.a || val > p.b) return NGX_ST_RANGE_ERROR;
        *lnp = 0.0;
        return NGX_ST_OK;
    case NGX_P_ERF: {
        double p1 = 0.5 * erf((p.c - val) / p.d);
        double p2 = 0.5 * erf((val - p.a) / p.b);
        double pr = p1 + p2;
        *lnp = (pr <= 0.0) ? -INFINITY : log(pr);
        return NGX_ST_OK;
    }
    case NGX_P_NORMAL: {
        double s2inv = 1.0 / (p.b * p.b);
        double diff = p.a - val;
        *lnp = -0.5 * diff * diff * s2inv;
        return NGX_ST_OK;
    }
    case NGX_P_LOGNORMAL: {
        double v = val - p.d;
        if (v <= 0.0) return NGX_ST_RANGE_ERROR;
        double logval = log(v);
        double dl = logval - p.a;
        double chi2 = p.b * (dl * dl);
        *lnp = -0.5 * chi2 - logval - p.c;
        return NGX_ST_OK;
    }
    case NGX_P_TRUNCGAUSS: {
        if (val < p.c || val > p.d) return NGX_ST_RANGE_ERROR;
        double ivar = 1.0 / (p.b * p.b);
        double diff = val - p.a;
        *lnp = -0.5 * diff * diff * ivar;
        return NGX_ST_OK;
    }
    default:
        *lnp = 0.0;
        return NGX_ST_OK;
    }
}

__device__ inline int prior1d_fdiff(const ngx_prior1d &p, double val,
                                    double *fd) {
    if (p.kind == NGX_P_NORMAL) {
        double sinv = 1.0 / p.b;
        *fd = (val - p.a) * sinv;
        return NGX_ST_OK;
    }
    if (p.kind == NGX_P_TRUNCGAUSS) {
        if (val < p.c || val > p.d) return NGX_ST_RANGE_ERROR;
        *fd = (val - p.a) * (1.0 / p.b);
        return NGX_ST_OK;
    }
    double lnp;
    int st = prior1d_lnprob(p, val, &lnp);
    if (st) return st;
    if (p.kind == NGX_P_FLAT) { *fd = 0.0; return NGX_ST_OK; }
    double c = -2 * lnp;
    if (c < 0.0) c = 0.0;
    *fd = sqrt(c);
    return NGX_ST_OK;
}

static __device__ __noinline__ int gprior_lnprob(const ngx_prior &pr, double g1, double g2,
                                    double *lnp) {
    double sig2 = pr.g_sigma * pr.g_sigma;
    double sig2inv = 1.0 / sig2;
    double gsq = g1 * g1 + g2 * g2;
    double omgsq = 1.0 - gsq;
    if (omgsq <= 0.0) return NGX_ST_RANGE_ERROR;
    *lnp = 2 * log(omgsq) - 0.5 * gsq * sig2inv;
    return NGX_ST_OK;
}

__device__ inline double clip_sqrt(double lnp) {
    double c = -2 * lnp;
    if (c < 0.0) c = 0.0;
    return sqrt(c);
}

__host__ __device__ inline int prior_nband(const ngx_prior &pr) {
    return pr.nband > 1 ? pr.nband : 1;
}
// flux prior of a band
__device__ inline const ngx_prior1d &prior_F(const ngx_prior &pr, int band) {
    return band == 0 ? pr.F
        : band <= 2 ? pr.F_more[band - 1]
        : band <= 6 ? pr.F_extra[band - 3] : pr.F_xwide[band - 7];
}
// number of residual slots reserved for the prior (results.py:1050-1076)
__host__ __device__ inline int prior_nslots(int model, const ngx_prior &pr) {
    if (!pr.present) return 0;
    if (model == NGX_COELLIP) return 3 + 2 * pr.ncoellip;
    return (model == NGX_BD ? 6 : 5) + prior_nband(pr);
}
// number of rows the prior actually fills
__host__ __device__ inline int prior_nrows(int model, const ngx_prior &pr) {
    if (!pr.present) return 0;
    if (model == NGX_BDF) return 5 + prior_nband(pr);
    if (model == NGX_BD) return 6 + prior_nband(pr);
    if (model == NGX_COELLIP) return 3 + 2 * pr.ncoellip;
    return 4 + prior_nband(pr);
}

// one row of prior.fill_fdiff; row index in the reference's order.
// simple: 0 cen1, 1 cen2, 2 g, 3 T, 4.. F per band   (sqrt(clip(-2 lnp)))
// bdf   : 0 cen1, 1 cen2, 2 g, 3 T, 4 fracdev, 5.. F per band (signed
//         get_fdiff)
// bd    : 0 cen1, 1 cen2, 2 g, 3 T, 4 logTratio, 5 fracdev, 6.. F per band
static __device__ __noinline__ int prior_row(int model, const ngx_prior &pr,
                                const double *p, int row, double *out) {
    double lnp;
    int st;
    if (model == NGX_BD) {
        switch (row) {
        case 0: *out = (p[0] - pr.cen1) * (1.0 / pr.sigma1); return 0;
        case 1: *out = (p[1] - pr.cen2) * (1.0 / pr.sigma2); return 0;
        case 2:
            st = gprior_lnprob(pr, p[2], p[3], &lnp);
            if (st) return st;
            *out = clip_sqrt(lnp);
            return 0;
        case 3: return prior1d_fdiff(pr.T, p[4], out);
        case 4: return prior1d_fdiff(pr.logTratio, p[5], out);
        case 5: return prior1d_fdiff(pr.fracdev, p[6], out);
        default:
            return prior1d_fdiff(prior_F(pr, row - 6), p[7 + (row - 6)], out);
        }
    }
    if (model == NGX_BDF) {
        switch (row) {
        case 0: *out = (p[0] - pr.cen1) * (1.0 / pr.sigma1); return 0;
        case 1: *out = (p[1] - pr.cen2) * (1.0 / pr.sigma2); return 0;
        case 2:
            st = gprior_lnprob(pr, p[2], p[3], &lnp);
            if (st) return st;
            *out = clip_sqrt(lnp);
            return 0;
        case 3: return prior1d_fdiff(pr.T, p[4], out);
        case 4: return prior1d_fdiff(pr.fracdev, p[5], out);
        default:
            return prior1d_fdiff(prior_F(pr, row - 5), p[6 + (row - 5)], out);
        }
    }
    switch (row) {
    case 0: {
        double d1 = pr.cen1 - p[0];
        lnp = -0.5 * d1 * d1 * (1.0 / (pr.sigma1 * pr.sigma1));
        break;
    }
    case 1: {
        double d2 = pr.cen2 - p[1];
        lnp = -0.5 * d2 * d2 * (1.0 / (pr.sigma2 * pr.sigma2));
        break;
    }
    case 2:
        st = gprior_lnprob(pr, p[2], p[3], &lnp);
        if (st) return st;
        break;
    default:
        if (model == NGX_COELLIP) {
            // PriorCoellipSame.fill_fdiff (joint_prior.py:963-990): the T
            // prior on every T_i, then the flux prior on every F_i
            const int ng = pr.ncoellip, k = row - 3;
            st = k < ng ? prior1d_lnprob(pr.T, p[4 + k], &lnp)
                        : prior1d_lnprob(pr.F, p[4 + k], &lnp);
        } else if (row == 3) {
            st = prior1d_lnprob(pr.T, p[4], &lnp);
        } else {
            st = prior1d_lnprob(prior_F(pr, row - 4), p[5 + (row - 4)], &lnp);
        }
        if (st) return st;
    }
    *out = clip_sqrt(lnp);
    return 0;
}

// which row depends on parameter ipar
__device__ inline int prior_row_of_par(int model, int ipar) {
    if (ipar < 2) return ipar;
    if (ipar < 4) return 2;
    // bdf: 4->3 (T), 5->4, 6->5; bd: 4->3, 5->4, 6->5, 7->6; simple: 4->3 (T),
    // 5.. -> 4.. (F per band)
    return ipar - 1;
}

// all rows; returns NGX_ST_*
__device__ inline int prior_fill_fdiff(int model, const ngx_prior &pr,
                                       const double *p, double *f) {
    int nr = prior_nrows(model, pr);
    for (int r = 0; r < nr; r++) {
        int st = prior_row(model, pr, p, r, &f[r]);
        if (st) return st;
    }
    return NGX_ST_OK;
}

// prior.get_lnprob_scalar
static __device__ __noinline__ int prior_lnprob(int model, const ngx_prior &pr,
                                   const double *p, double *out) {
    if (!pr.present) { *out = 0.0; return NGX_ST_OK; }
    double s2inv1 = 1.0 / (pr.sigma1 * pr.sigma1);
    double s2inv2 = 1.0 / (pr.sigma2 * pr.sigma2);
    double d1 = pr.cen1 - p[0];
    double d2 = pr.cen2 - p[1];
    double lnp = -0.5 * d1 * d1 * s2inv1 - 0.5 * d2 * d2 * s2inv2;
    double t;
    if (gprior_lnprob(pr, p[2], p[3], &t)) return NGX_ST_RANGE_ERROR;
    lnp += t;
    if (model == NGX_COELLIP) {
        // PriorCoellipSame.get_lnprob_scalar, joint_prior.py:941-961
        const int ng = pr.ncoellip;
        for (int k = 0; k < 2 * ng; k++) {
            if (prior1d_lnprob(k < ng ? pr.T : pr.F, p[4 + k], &t))
                return NGX_ST_RANGE_ERROR;
            lnp += t;
        }
        *out = lnp;
        return NGX_ST_OK;
    }
    if (prior1d_lnprob(pr.T, p[4], &t)) return NGX_ST_RANGE_ERROR;
    lnp += t;
    int fi = 5;
    if (model == NGX_BD) {
        if (prior1d_lnprob(pr.logTratio, p[5], &t)) return NGX_ST_RANGE_ERROR;
        lnp += t;
        if (prior1d_lnprob(pr.fracdev, p[6], &t)) return NGX_ST_RANGE_ERROR;
        lnp += t;
        fi = 7;
    }
    if (model == NGX_BDF) {
        if (prior1d_lnprob(pr.fracdev, p[5], &t)) return NGX_ST_RANGE_ERROR;
        lnp += t;
        fi = 6;
    }
    for (int b = 0; b < prior_nband(pr); b++) {
        if (prior1d_lnprob(prior_F(pr, b), p[fi + b], &t))
            return NGX_ST_RANGE_ERROR;
        lnp += t;
    }
    *out = lnp;
    return NGX_ST_OK;
}

// Forward-difference Jacobian of the prior rows, one-sided with backward
// fallback.  Row r depends only on the parameters prior_row_of_par maps to
// it, so the other entries of the reference's dense difference are exactly
// (f0 - f0)/step = 0 and are not evaluated.  jp is [nrows][np].
static __device__ __noinline__ int prior_jacobian(int model, const ngx_prior &pr,
                                     const double *pars, int np,
                                     const double *f0, double *jp) {
    int nr = prior_nrows(model, pr);
    for (int i = 0; i < nr * np; i++) jp[i] = 0.0;
    double p[LM_MAXP];
    for (int ipar = 0; ipar < np; ipar++) {
        int r = prior_row_of_par(model, ipar);
        double step = NGX_STEP_PRIOR * fmax(1.0, fabs(pars[ipar]));
        for (int k = 0; k < np; k++) p[k] = pars[k];
        p[ipar] = pars[ipar] + step;
        double fs;
        if (prior_row(model, pr, p, r, &fs)) {
            step = -step;
            p[ipar] = pars[ipar] + step;
            if (prior_row(model, pr, p, r, &fs)) return NGX_ST_RANGE_ERROR;
        }
        double d = 0.0;
        if (isfinite(f0[r]) && isfinite(fs)) d = (fs - f0[r]) / step;
        jp[r * np + ipar] = d;
    }
    return NGX_ST_OK;
}

}  // namespace ngx
