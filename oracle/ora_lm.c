/*
 * TEST INFRASTRUCTURE (see oracle.h).  CPU restatement of the reference's
 * Levenberg-Marquardt fit:
 *   Fitter.go                  ngmix/fitting/fitters.py:64-112
 *   FitModel.calc_fdiff        ngmix/fitting/results.py:439-466
 *   FitModel.calc_jacobian     ngmix/fitting/results.py:487-570
 *   _fill_prior_jacobian       ngmix/fitting/results.py:572-625
 *   get_model_deriv_data       ngmix/fitting/results.py:955-1010
 *   PriorSimpleSep.fill_fdiff  ngmix/joint_prior.py:86-120
 *   PriorBDFSep.fill_fdiff     ngmix/joint_prior.py:565-599
 *   run_leastsq / _test_cov    ngmix/fitting/leastsqbound.py:33-180
 *   FitModel.set_fit_result    ngmix/fitting/results.py:45-72,142-210
 * ora_lm_fit: single band, single epoch; ora_lm_fit_multi: ObsList /
 * MultiBandObsList (simple models).  No bounds.
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

#define PDEF (-9.999e9)
#define CDEF (9.999e9)
#define BIGVAL (9999.0e47)
#define STEP_PRIOR 1.0e-8

double ora_enorm(int n, const double *x);

static int is_simple(int model) {
    return model == ORA_GAUSS || model == ORA_EXP || model == ORA_DEV
        || model == ORA_TURB;
}

/* ---------------- 1-d priors: lnprob and get_fdiff ---------------- */
/* priors/priors.py: FlatPrior :85-99, TwoSidedErf :219-247, Normal :424-437 */
static int prior1d_lnprob(const ora_prior1d *p, double val, double *lnp) {
    switch (p->kind) {
    case ORA_P_FLAT:
        if (val < p->a || val > p->b) return ORA_RANGE_ERROR;
        *lnp = 0.0;
        return ORA_OK;
    case ORA_P_ERF: {
        /* a=minval b=width_at_min c=maxval d=width_at_max */
        double p1 = 0.5 * erf((p->c - val) / p->d);
        double p2 = 0.5 * erf((val - p->a) / p->b);
        double pr = p1 + p2;
        *lnp = (pr <= 0.0) ? -INFINITY : log(pr);
        return ORA_OK;
    }
    case ORA_P_NORMAL: {
        /* a=mean b=sigma */
        double s2inv = 1.0 / (p->b * p->b);
        double diff = p->a - val;
        *lnp = -0.5 * diff * diff * s2inv;
        return ORA_OK;
    }
    case ORA_P_LOGNORMAL: {
        /* LogNormal.get_lnprob_scalar, priors/priors.py:741-765:
         * a=logmean b=logivar c=lnprob_max d=shift */
        double v = val - p->d;
        if (v <= 0.0) return ORA_RANGE_ERROR;
        double logval = log(v);
        double chi2 = p->b * ((logval - p->a) * (logval - p->a));
        *lnp = -0.5 * chi2 - logval - p->c;
        return ORA_OK;
    }
    case ORA_P_TRUNCGAUSS: {
        /* TruncatedGaussian.get_lnprob_scalar, priors/priors.py:1080-1094:
         * a=mean b=sigma c=minval d=maxval */
        if (val < p->c || val > p->d) return ORA_RANGE_ERROR;
        double ivar = 1.0 / (p->b * p->b);
        double diff = val - p->a;
        *lnp = -0.5 * diff * diff * ivar;
        return ORA_OK;
    }
    default:
        *lnp = 0.0;
        return ORA_OK;
    }
}

/* get_fdiff of the 1-d priors: Flat :143-157, TwoSidedErf :297-311,
 * Normal :469-478 */
static int prior1d_fdiff(const ora_prior1d *p, double val, double *fd) {
    if (p->kind == ORA_P_NORMAL) {
        double sinv = 1.0 / p->b;
        *fd = (val - p->a) * sinv;
        return ORA_OK;
    }
    if (p->kind == ORA_P_TRUNCGAUSS) {
        /* TruncatedGaussian.get_fdiff, priors/priors.py:1115-1126 */
        if (val < p->c || val > p->d) return ORA_RANGE_ERROR;
        *fd = (val - p->a) * (1.0 / p->b);
        return ORA_OK;
    }
    double lnp;
    int st = prior1d_lnprob(p, val, &lnp);
    if (st) return st;
    if (p->kind == ORA_P_FLAT) { *fd = 0.0; return ORA_OK; }
    double c = -2 * lnp;
    if (c < 0.0) c = 0.0;
    *fd = sqrt(c);
    return ORA_OK;
}

/* GPriorBA.get_lnprob_scalar2d, priors/shape.py:555-564 */
static int gprior_lnprob(const ora_prior *pr, double g1, double g2,
                         double *lnp) {
    double sig2 = pr->g_sigma * pr->g_sigma;
    double sig2inv = 1.0 / sig2;
    double gsq = g1 * g1 + g2 * g2;
    double omgsq = 1.0 - gsq;
    if (omgsq <= 0.0) return ORA_RANGE_ERROR;
    *lnp = 2 * log(omgsq) - 0.5 * gsq * sig2inv;
    return ORA_OK;
}

/* number of prior slots, results.py:1050-1076 (nband = 1) */
static int n_prior_slots(int model, const ora_prior *pr) {
    if (!pr || !pr->present) return 0;
    if (model == ORA_BD) return 7;
    /* CoellipFitModel._set_n_prior_pars, results.py:652-660 */
    if (model == ORA_COELLIP) return 3 + 2 * pr->ncoellip;
    return 6;
}

/* prior.fill_fdiff: returns number of rows filled, or -1 on range error */
static int prior_fill_fdiff(int model, const ora_prior *pr, const double *pars,
                            double *fdiff) {
    if (!pr || !pr->present) return 0;
    int index = 0;
    if (model == ORA_BD) {
        /* PriorBDSep.fill_fdiff, joint_prior.py:364-399, signed rows */
        double sinv1 = 1.0 / pr->sigma1, sinv2 = 1.0 / pr->sigma2;
        fdiff[index++] = (pars[0] - pr->cen1) * sinv1;
        fdiff[index++] = (pars[1] - pr->cen2) * sinv2;
        double lnp;
        if (gprior_lnprob(pr, pars[2], pars[3], &lnp)) return -1;
        double c = -2 * lnp;
        if (c < 0.0) c = 0.0;
        fdiff[index++] = sqrt(c);
        double fd;
        if (prior1d_fdiff(&pr->T, pars[4], &fd)) return -1;
        fdiff[index++] = fd;
        if (prior1d_fdiff(&pr->logTratio, pars[5], &fd)) return -1;
        fdiff[index++] = fd;
        if (prior1d_fdiff(&pr->fracdev, pars[6], &fd)) return -1;
        fdiff[index++] = fd;
        if (prior1d_fdiff(&pr->F, pars[7], &fd)) return -1;
        fdiff[index++] = fd;
        return index;
    }
    if (model == ORA_BDF) {
        /* joint_prior.py:565-599, signed rows */
        double sinv1 = 1.0 / pr->sigma1, sinv2 = 1.0 / pr->sigma2;
        fdiff[index++] = (pars[0] - pr->cen1) * sinv1;
        fdiff[index++] = (pars[1] - pr->cen2) * sinv2;
        double lnp;
        if (gprior_lnprob(pr, pars[2], pars[3], &lnp)) return -1;
        double c = -2 * lnp;
        if (c < 0.0) c = 0.0;
        fdiff[index++] = sqrt(c);
        double fd;
        if (prior1d_fdiff(&pr->T, pars[4], &fd)) return -1;
        fdiff[index++] = fd;
        if (prior1d_fdiff(&pr->fracdev, pars[5], &fd)) return -1;
        fdiff[index++] = fd;
        if (prior1d_fdiff(&pr->F, pars[6], &fd)) return -1;
        fdiff[index++] = fd;
        return index;
    }
    /* joint_prior.py:86-120: sqrt(clip(-2 lnp)) rows */
    double s2inv1 = 1.0 / (pr->sigma1 * pr->sigma1);
    double s2inv2 = 1.0 / (pr->sigma2 * pr->sigma2);
    double d1 = pr->cen1 - pars[0];
    double d2 = pr->cen2 - pars[1];
    fdiff[index++] = -0.5 * d1 * d1 * s2inv1;
    fdiff[index++] = -0.5 * d2 * d2 * s2inv2;
    double lnp;
    if (gprior_lnprob(pr, pars[2], pars[3], &lnp)) return -1;
    fdiff[index++] = lnp;
    if (model == ORA_COELLIP) {
        /* PriorCoellipSame.fill_fdiff, joint_prior.py:963-990 */
        int ng = pr->ncoellip;
        for (int i = 0; i < ng; i++) {
            if (prior1d_lnprob(&pr->T, pars[4 + i], &lnp)) return -1;
            fdiff[index++] = lnp;
        }
        for (int i = 0; i < ng; i++) {
            if (prior1d_lnprob(&pr->F, pars[4 + ng + i], &lnp)) return -1;
            fdiff[index++] = lnp;
        }
    } else {
        if (prior1d_lnprob(&pr->T, pars[4], &lnp)) return -1;
        fdiff[index++] = lnp;
        if (prior1d_lnprob(&pr->F, pars[5], &lnp)) return -1;
        fdiff[index++] = lnp;
    }
    for (int i = 0; i < index; i++) {
        double c = -2 * fdiff[i];
        if (c < 0.0) c = 0.0;       /* clip(min=0); NaN passes through */
        fdiff[i] = sqrt(c);
    }
    return index;
}

/* prior.get_lnprob_scalar, joint_prior.py:136-150 / :543-563 */
static int prior_lnprob(int model, const ora_prior *pr, const double *pars,
                        double *out) {
    if (!pr || !pr->present) { *out = 0.0; return ORA_OK; }
    double s2inv1 = 1.0 / (pr->sigma1 * pr->sigma1);
    double s2inv2 = 1.0 / (pr->sigma2 * pr->sigma2);
    double d1 = pr->cen1 - pars[0];
    double d2 = pr->cen2 - pars[1];
    double lnp = -0.5 * d1 * d1 * s2inv1 - 0.5 * d2 * d2 * s2inv2;
    double t;
    if (gprior_lnprob(pr, pars[2], pars[3], &t)) return ORA_RANGE_ERROR;
    lnp += t;
    if (model == ORA_COELLIP) {
        /* PriorCoellipSame.get_lnprob_scalar, joint_prior.py:941-961 */
        int ng = pr->ncoellip;
        for (int i = 0; i < 2 * ng; i++) {
            if (prior1d_lnprob(i < ng ? &pr->T : &pr->F, pars[4 + i], &t))
                return ORA_RANGE_ERROR;
            lnp += t;
        }
        *out = lnp;
        return ORA_OK;
    }
    if (prior1d_lnprob(&pr->T, pars[4], &t)) return ORA_RANGE_ERROR;
    lnp += t;
    int fi = 5;
    if (model == ORA_BD) {
        /* PriorBDSep.get_lnprob_scalar, joint_prior.py:343-362 */
        if (prior1d_lnprob(&pr->logTratio, pars[5], &t)) return ORA_RANGE_ERROR;
        lnp += t;
        if (prior1d_lnprob(&pr->fracdev, pars[6], &t)) return ORA_RANGE_ERROR;
        lnp += t;
        fi = 7;
    }
    if (model == ORA_BDF) {
        if (prior1d_lnprob(&pr->fracdev, pars[5], &t)) return ORA_RANGE_ERROR;
        lnp += t;
        fi = 6;
    }
    if (prior1d_lnprob(&pr->F, pars[fi], &t)) return ORA_RANGE_ERROR;
    lnp += t;
    *out = lnp;
    return ORA_OK;
}

/* ---------------- FitModel context ---------------- */
typedef struct {
    int model, npars, ngauss0, npsf, ngauss;
    const ora_pixel *pixels;
    int npix;
    ora_gauss psf[16];
    ora_gauss gm0[16];
    ora_gauss *gm;          /* ngauss0 * max(npsf,1) */
    const ora_prior *prior;
    int n_prior_pars, fdiff_size;
    /* scratch for the jacobian */
    double *dout;           /* 6 * npix */
    double *vv, *uu, *area;
    /* bounded fits (leastsqbound.py:183-263): MINPACK works on internal
     * (unconstrained) parameters; NULL when no prior member has bounds */
    const ora_bounds *bounds;
} fitctx;

/* _internal2external_lambda, leastsqbound.py:216-229 */
static double b_i2e(const ora_bounds *b, int i, double x) {
    switch (b->type[i]) {
    case 1: return b->lo[i] - 1.0 + sqrt(x * x + 1.0);
    case 2: return b->hi[i] + 1.0 - sqrt(x * x + 1.0);
    case 3: return b->lo[i] + ((b->hi[i] - b->lo[i]) / 2.0) * (sin(x) + 1.0);
    default: return x;
    }
}
/* _external2internal_lambda, leastsqbound.py:248-261 */
static double b_e2i(const ora_bounds *b, int i, double x) {
    double t;
    switch (b->type[i]) {
    case 1: t = x - b->lo[i] + 1.0; return sqrt(t * t - 1.0);
    case 2: t = b->hi[i] - x + 1.0; return sqrt(t * t - 1.0);
    case 3: return asin((2.0 * (x - b->lo[i]) / (b->hi[i] - b->lo[i])) - 1.0);
    default: return x;
    }
}
/* _internal2external_grad, leastsqbound.py:183-199 */
static double b_grad(const ora_bounds *b, int i, double v) {
    switch (b->type[i]) {
    case 1: return v / sqrt(v * v + 1.0);
    case 2: return -v / sqrt(v * v + 1.0);
    case 3: return (b->hi[i] - b->lo[i]) * cos(v) / 2.0;
    default: return 1.0;
    }
}

/* _fill_gmix_all, results.py:306-347 */
static int ctx_fill_gmix(fitctx *c, const double *pars) {
    if (c->npsf > 0) {
        if (ora_gmix_fill(c->model, pars, c->npars, c->gm0, c->ngauss0))
            return ORA_RANGE_ERROR;
        ora_gmix_convolve_fill(c->gm, c->gm0, c->ngauss0, c->psf, c->npsf);
    } else {
        if (ora_gmix_fill(c->model, pars, c->npars, c->gm, c->ngauss0))
            return ORA_RANGE_ERROR;
    }
    return ORA_OK;
}

/* calc_fdiff, results.py:439-466 */
static int ctx_calc_fdiff(fitctx *c, const double *pars, double *fdiff) {
    int m = c->fdiff_size;
    for (int i = 0; i < m; i++) fdiff[i] = 0.0;
    int bad = 0;
    if (ctx_fill_gmix(c, pars)) bad = 1;
    int start = 0;
    if (!bad) {
        start = prior_fill_fdiff(c->model, c->prior, pars, fdiff);
        if (start < 0) bad = 1;
    }
    if (!bad) {
        if (ora_fill_fdiff(c->gm, c->ngauss, c->pixels, c->npix, fdiff,
                           start))
            bad = 1;
    }
    if (bad) {
        for (int i = 0; i < m; i++) fdiff[i] = -INFINITY;   /* LOWVAL */
        return ORA_RANGE_ERROR;
    }
    return ORA_OK;
}

/* _fill_prior_jacobian, results.py:572-625; jac is column major m x npars */
static int ctx_prior_jacobian(fitctx *c, const double *pars, double *jac,
                              int *nrows) {
    *nrows = 0;
    if (!c->prior || !c->prior->present) return ORA_OK;
    int m = c->fdiff_size, np = c->npars;
    double f0[ORA_MAXP] = {0}, fs[ORA_MAXP] = {0}, p[ORA_MAXP];
    int n = prior_fill_fdiff(c->model, c->prior, pars, f0);
    if (n < 0) return ORA_RANGE_ERROR;
    for (int ipar = 0; ipar < np; ipar++) {
        double step = STEP_PRIOR * fmax(1.0, fabs(pars[ipar]));
        for (int k = 0; k < np; k++) p[k] = pars[k];
        p[ipar] = pars[ipar] + step;
        if (prior_fill_fdiff(c->model, c->prior, p, fs) < 0) {
            step = -step;
            p[ipar] = pars[ipar] + step;
            if (prior_fill_fdiff(c->model, c->prior, p, fs) < 0)
                return ORA_RANGE_ERROR;
        }
        for (int k = 0; k < n; k++) {
            double d = 0.0;
            if (isfinite(f0[k]) && isfinite(fs[k])) d = (fs[k] - f0[k]) / step;
            jac[k + (size_t)ipar * m] = d;
        }
    }
    *nrows = n;
    return ORA_OK;
}

/* calc_jacobian, results.py:487-570; simple models only */
static int ctx_calc_jacobian(fitctx *c, const double *pars, double *jac) {
    int m = c->fdiff_size, np = c->npars;
    memset(jac, 0, sizeof(double) * (size_t)m * np);
    int start = 0;
    int bad = 0;
    if (ctx_prior_jacobian(c, pars, jac, &start)) bad = 1;
    if (!bad && ctx_fill_gmix(c, pars)) bad = 1;
    double g1 = pars[2], g2 = pars[3], T = pars[4], flux = pars[5];
    if (!bad && (T == 0.0 || flux == 0.0)) bad = 1;
    if (!bad) {
        /* get_model_deriv_data, results.py:955-1010 */
        int ng = c->ngauss;
        int npsf = c->npsf > 0 ? c->npsf : 1;
        double gpars[6 * 160];
        double dcov[9 * 160];
        const ora_gauss *gm0 = c->npsf > 0 ? c->gm0 : c->gm;
        double gsq = g1 * g1 + g2 * g2;
        double f = 2.0 / (1.0 + gsq);
        double dfac = -f / (1.0 + gsq);
        double de1dg1 = f + 2.0 * g1 * g1 * dfac;
        double de1dg2 = 2.0 * g1 * g2 * dfac;
        double de2dg1 = de1dg2;
        double de2dg2 = f + 2.0 * g2 * g2 * dfac;
        for (int ig = 0; ig < ng; ig++) {
            const ora_gauss *g = &c->gm[ig];
            const ora_gauss *g0 = &gm0[ig / npsf];
            double *gp = gpars + 6 * ig;
            gp[0] = g->p; gp[1] = g->row; gp[2] = g->col;
            gp[3] = g->irr; gp[4] = g->irc; gp[5] = g->icc;
            double Tk = g0->irr + g0->icc;
            double *dc = dcov + 9 * ig;
            dc[0] = -0.5 * Tk * de1dg1; dc[1] = 0.5 * Tk * de2dg1;
            dc[2] = 0.5 * Tk * de1dg1;
            dc[3] = -0.5 * Tk * de1dg2; dc[4] = 0.5 * Tk * de2dg2;
            dc[5] = 0.5 * Tk * de1dg2;
            dc[6] = g0->irr / T; dc[7] = g0->irc / T; dc[8] = g0->icc / T;
        }
        memset(c->dout, 0, sizeof(double) * 6 * (size_t)c->npix);
        ora_deriv_images(gpars, dcov, ng, c->vv, c->uu, c->area, c->npix,
                         c->dout);
        for (int i = 0; i < c->npix; i++) {
            double ierr = c->pixels[i].ierr;
            for (int k = 0; k < 5; k++)
                jac[start + i + (size_t)k * m] =
                    c->dout[(size_t)(1 + k) * c->npix + i] * ierr;
            jac[start + i + (size_t)5 * m] = c->dout[i] * (ierr / flux);
        }
    }
    if (bad) {
        memset(jac, 0, sizeof(double) * (size_t)m * np);
        return ORA_RANGE_ERROR;
    }
    return ORA_OK;
}

/* wfunc / wDfun of leastsqbound (leastsqbound.py:469-489): the residuals at
 * i2e(x); the jacobian's columns scaled by d(external)/d(internal) */
static int cb_fdiff(void *ctx, int m, int n, const double *x, double *fvec) {
    (void)m;
    fitctx *c = (fitctx *)ctx;
    if (c->bounds) {
        double xe[ORA_MAXP];
        for (int i = 0; i < n; i++) xe[i] = b_i2e(c->bounds, i, x[i]);
        ctx_calc_fdiff(c, xe, fvec);
        return 0;
    }
    ctx_calc_fdiff(c, x, fvec);
    return 0;
}

static int cb_jac(void *ctx, int m, int n, const double *x, double *fjac) {
    fitctx *c = (fitctx *)ctx;
    if (c->bounds) {
        double xe[ORA_MAXP] = {0};
        for (int i = 0; i < n; i++) xe[i] = b_i2e(c->bounds, i, x[i]);
        ctx_calc_jacobian(c, xe, fjac);
        for (int j = 0; j < n; j++) {
            double g = b_grad(c->bounds, j, x[j]);
            for (int i = 0; i < m; i++) fjac[i + (size_t)j * m] *= g;
        }
        return 0;
    }
    ctx_calc_jacobian(c, x, fjac);
    return 0;
}

/* symmetric eigenvalues by cyclic Jacobi; returns 0 ok, 1 not finite */
static int sym_eigvals(int n, const double *a_in, double *ev) {
    double a[ORA_MAXP * ORA_MAXP];
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            double v = 0.5 * (a_in[i * n + j] + a_in[j * n + i]);
            if (!isfinite(v)) return 1;
            a[i * n + j] = v;
        }
    for (int sweep = 0; sweep < 100; sweep++) {
        double off = 0.0, dg = 0.0;
        for (int i = 0; i < n; i++) {
            dg += a[i * n + i] * a[i * n + i];
            for (int j = i + 1; j < n; j++) off += a[i * n + j] * a[i * n + j];
        }
        if (off <= 1e-40 * dg || off == 0.0) break;
        for (int p = 0; p < n; p++)
            for (int q = p + 1; q < n; q++) {
                double apq = a[p * n + q];
                if (apq == 0.0) continue;
                double theta = (a[q * n + q] - a[p * n + p]) / (2.0 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0)
                    / (fabs(theta) + sqrt(theta * theta + 1.0));
                double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
                for (int k = 0; k < n; k++) {
                    double akp = a[k * n + p], akq = a[k * n + q];
                    a[k * n + p] = cs * akp - sn * akq;
                    a[k * n + q] = sn * akp + cs * akq;
                }
                for (int k = 0; k < n; k++) {
                    double apk = a[p * n + k], aqk = a[q * n + k];
                    a[p * n + k] = cs * apk - sn * aqk;
                    a[q * n + k] = sn * apk + cs * aqk;
                }
            }
    }
    for (int i = 0; i < n; i++) ev[i] = a[i * n + i];
    return 0;
}

static void set_defaults(ora_lm_result *res, int np) {
    for (int i = 0; i < np; i++) {
        res->pars[i] = PDEF;
        res->pars_err[i] = CDEF;
    }
    for (int i = 0; i < np * np; i++) res->pars_cov[i] = CDEF;
}

static int ctx_init(fitctx *c, int model, const ora_pixel *pixels, int npix,
                    const double *psf_pars, int npsf, const ora_prior *prior) {
    memset(c, 0, sizeof(*c));
    c->model = model;
    c->npars = ora_model_npars(model);
    c->ngauss0 = ora_model_ngauss(model);
    if (model == ORA_COELLIP) {
        /* CoellipFitModel._set_npars, results.py:662-666 */
        c->ngauss0 = prior ? prior->ncoellip : 0;
        c->npars = 4 + 2 * c->ngauss0;
    }
    c->npsf = npsf;
    c->ngauss = c->ngauss0 * (npsf > 0 ? npsf : 1);
    c->pixels = pixels;
    c->npix = npix;
    c->prior = prior;
    c->n_prior_pars = n_prior_slots(model, prior);
    c->fdiff_size = npix + c->n_prior_pars;
    if (npsf > 0) ora_gmix_fill(ORA_FULL, psf_pars, 6 * npsf, c->psf, npsf);
    c->gm = (ora_gauss *)calloc((size_t)c->ngauss, sizeof(ora_gauss));
    c->dout = (double *)malloc(sizeof(double) * 9 * (size_t)npix);
    c->vv = c->dout + 6 * (size_t)npix;
    c->uu = c->vv + npix;
    c->area = c->uu + npix;
    for (int i = 0; i < npix; i++) {
        c->vv[i] = pixels[i].v;
        c->uu[i] = pixels[i].u;
        c->area[i] = pixels[i].area;
    }
    return 0;
}

static void ctx_free(fitctx *c) {
    free(c->gm);
    free(c->dout);
}

int ora_lm_calc_fdiff(int model, const ora_pixel *pixels, int npix,
                      const double *psf_pars, int npsf, const ora_prior *prior,
                      const double *pars, double *fdiff, int fdiff_size) {
    fitctx c;
    ctx_init(&c, model, pixels, npix, psf_pars, npsf, prior);
    if (fdiff_size != c.fdiff_size) { ctx_free(&c); return -1; }
    int st = ctx_calc_fdiff(&c, pars, fdiff);
    ctx_free(&c);
    return st;
}

int ora_lm_calc_jacobian(int model, const ora_pixel *pixels, int npix,
                         const double *psf_pars, int npsf,
                         const ora_prior *prior, const double *pars,
                         double *jac, int fdiff_size) {
    fitctx c;
    ctx_init(&c, model, pixels, npix, psf_pars, npsf, prior);
    if (fdiff_size != c.fdiff_size) { ctx_free(&c); return -1; }
    int st = ctx_calc_jacobian(&c, pars, jac);
    ctx_free(&c);
    return st;
}

/* Fitter.go on one stamp; bounds (nullable) = prior.bounds of the reference:
 * leastsqbound's transformed fit, leastsqbound.py:445-552 */
int ora_lm_fit_bounded(int model, const double *image, const double *weight,
                       int nrow, int ncol, const ora_jac *jac,
                       const double *psf_pars, int npsf, const double *guess,
                       const ora_prior *prior, const ora_bounds *bounds,
                       int maxfev, double ftol, double xtol,
                       int analytic_jacobian, int ignore_zero_weight,
                       ora_lm_result *res) {
    memset(res, 0, sizeof(*res));
    if (bounds && !bounds->present) bounds = NULL;
    int ntot = nrow * ncol;
    ora_pixel *pixels = (ora_pixel *)malloc(sizeof(ora_pixel) * (size_t)ntot);
    int npixv = ntot;
    if (ignore_zero_weight) {
        npixv = 0;
        for (int i = 0; i < ntot; i++) if (weight[i] > 0.0) npixv++;
        if (npixv == 0) { free(pixels); return -2; }   /* GMixFatalError */
    }
    ora_fill_pixels(pixels, npixv, image, weight, nrow, ncol, jac,
                    ignore_zero_weight);
    fitctx c;
    ctx_init(&c, model, pixels, npixv, psf_pars, npsf, prior);
    c.bounds = bounds;
    int np = c.npars, m = c.fdiff_size;

    double x[ORA_MAXP];
    for (int i = 0; i < np; i++)
        x[i] = bounds ? b_e2i(bounds, i, guess[i]) : guess[i];
    double *fvec = (double *)malloc(sizeof(double) * (size_t)m * (np + 1));
    double *fjac = fvec + m;
    int ipvt[ORA_MAXP];
    double qtf[ORA_MAXP];
    int nfev = 0, njev = 0, ier;
    int use_dfun = analytic_jacobian
        && (model == ORA_GAUSS || model == ORA_EXP || model == ORA_DEV);
    (void)is_simple;
    if (use_dfun)
        ier = ora_lmder(cb_fdiff, cb_jac, &c, m, np, x, fvec, fjac, ftol,
                        xtol, 0.0, maxfev, 100.0, &nfev, &njev, ipvt, qtf);
    else
        ier = ora_lmdif(cb_fdiff, &c, m, np, x, fvec, fjac, ftol, xtol, 0.0,
                        maxfev, 2.220446049250313e-16, 100.0, &nfev, ipvt,
                        qtf);

    /* run_leastsq, leastsqbound.py:70-126 */
    int flags = 0;
    res->nfev = nfev;
    res->njev = njev;
    res->ier = ier;
    if (bounds) {
        /* x = i2e(x_internal); fjac from internal to external parameters:
         * column j of R (parameter ipvt[j]) over that parameter's gradient
         * (leastsqbound.py:531-538, with the 1-based ipvt its code assumes) */
        for (int j = 0; j < np; j++) {
            double g = b_grad(bounds, ipvt[j], x[ipvt[j]]);
            for (int i = 0; i <= j; i++) fjac[i + (size_t)j * m] /= g;
        }
        for (int i = 0; i < np; i++) x[i] = b_i2e(bounds, i, x[i]);
    }
    for (int i = 0; i < np; i++) res->pars[i] = x[i];
    for (int i = 0; i < np * np; i++) res->pars_cov0[i] = NAN;

    /* cov_x = P R^-1 R^-T P^T (scipy leastsq), None if R is singular */
    int have_cov0 = 0;
    if (ier >= 1 && ier <= 4) {
        double rinv[ORA_MAXP * ORA_MAXP];
        int sing = 0;
        for (int j = 0; j < np; j++)
            if (fjac[j + (size_t)j * m] == 0.0) sing = 1;
        if (!sing) {
            /* invert upper triangular R (np x np in fjac) */
            memset(rinv, 0, sizeof(rinv));
            for (int j = 0; j < np; j++) {
                rinv[j * np + j] = 1.0 / fjac[j + (size_t)j * m];
                for (int i = j - 1; i >= 0; i--) {
                    double s = 0.0;
                    for (int k = i + 1; k <= j; k++)
                        s += fjac[i + (size_t)k * m] * rinv[k * np + j];
                    rinv[i * np + j] = -s / fjac[i + (size_t)i * m];
                }
            }
            double prinv[ORA_MAXP * ORA_MAXP];
            for (int j = 0; j < np; j++)
                for (int k = 0; k < np; k++)
                    prinv[ipvt[j] * np + k] = rinv[j * np + k];
            for (int a = 0; a < np; a++)
                for (int b = 0; b < np; b++) {
                    double s = 0.0;
                    for (int k = 0; k < np; k++)
                        s += prinv[a * np + k] * prinv[b * np + k];
                    res->pars_cov0[a * np + b] = s;
                }
            have_cov0 = 1;
            for (int i = 0; i < np * np; i++)
                if (!isfinite(res->pars_cov0[i])) have_cov0 = have_cov0;
        }
    }

    if (ier > 4) {
        flags |= 1 << (ier - 5);
        set_defaults(res, np);
    } else if (!have_cov0) {
        flags |= ORA_LM_SINGULAR_MATRIX;
        for (int i = 0; i < np; i++) res->pars_err[i] = CDEF;
        for (int i = 0; i < np * np; i++) res->pars_cov[i] = CDEF;
    } else {
        c.bounds = NULL;            /* func(pars) at the external pars */
        ctx_calc_fdiff(&c, res->pars, fvec);
        int dof = m - c.n_prior_pars - np;
        if (dof == 0) {
            for (int i = 0; i < np; i++) res->pars_err[i] = CDEF;
            for (int i = 0; i < np * np; i++) res->pars_cov[i] = CDEF;
            flags |= ORA_ZERO_DOF;
        } else {
            double ssum = 0.0;
            for (int i = c.n_prior_pars; i < m; i++) ssum += fvec[i] * fvec[i];
            double s_sq = ssum / dof;
            for (int i = 0; i < np * np; i++)
                res->pars_cov[i] = res->pars_cov0[i] * s_sq;
            int cflags = 0;
            double ev[ORA_MAXP];
            if (sym_eigvals(np, res->pars_cov, ev)) {
                cflags |= ORA_EIG_NOTFINITE;
            } else {
                for (int i = 0; i < np; i++)
                    if (ev[i] < 0) cflags |= ORA_LM_NEG_COV_EIG;
                for (int i = 0; i < np; i++)
                    if (res->pars_cov[i * np + i] < 0)
                        cflags |= ORA_LM_NEG_COV_DIAG;
            }
            if (cflags) {
                flags |= cflags;
                for (int i = 0; i < np; i++) res->pars_err[i] = CDEF;
            } else {
                for (int i = 0; i < np; i++)
                    res->pars_err[i] = sqrt(res->pars_cov[i * np + i]);
            }
        }
    }
    res->flags = flags;

    /* set_fit_result, results.py:45-72 */
    res->lnprob = NAN; res->s2n_numer = NAN; res->s2n_denom = NAN;
    res->chi2per = NAN; res->s2n = NAN; res->T = NAN; res->T_err = NAN;
    res->flux = NAN; res->flux_err = NAN;
    if (flags == 0) {
        double lnp_prior = 0.0, out4[4];
        int bad = 0;
        if (prior_lnprob(model, prior, res->pars, &lnp_prior)) bad = 1;
        if (!bad && ctx_fill_gmix(&c, res->pars)) bad = 1;
        if (!bad && ora_get_loglike(c.gm, c.ngauss, pixels, npixv, out4))
            bad = 1;
        if (bad) {
            res->lnprob = -INFINITY; res->s2n_numer = 0.0;
            res->s2n_denom = BIGVAL; res->npix = 0;
        } else {
            res->lnprob = out4[0] + lnp_prior;
            res->s2n_numer = out4[1]; res->s2n_denom = out4[2];
            res->npix = (int)out4[3];
        }
        double s2n = 0.0;
        if (res->s2n_denom > 0) s2n = res->s2n_numer / sqrt(res->s2n_denom);
        double chi2 = res->lnprob / (-0.5);
        int dof = res->npix - np;
        res->chi2per = chi2 / dof;
        res->dof = dof;
        res->s2n = s2n;
        res->T = res->pars[4];
        res->T_err = sqrt(res->pars_cov[4 * np + 4]);
        int fs = np - 1;
        res->flux = res->pars[fs];
        res->flux_err = sqrt(res->pars_cov[fs * np + fs]);
    }
    free(fvec);
    ctx_free(&c);
    free(pixels);
    return 0;
}

int ora_lm_fit(int model, const double *image, const double *weight,
               int nrow, int ncol, const ora_jac *jac, const double *psf_pars,
               int npsf, const double *guess, const ora_prior *prior,
               int maxfev, double ftol, double xtol, int analytic_jacobian,
               int ignore_zero_weight, ora_lm_result *res) {
    return ora_lm_fit_bounded(model, image, weight, nrow, ncol, jac, psf_pars,
                              npsf, guess, prior, NULL, maxfev, ftol, xtol,
                              analytic_jacobian, ignore_zero_weight, res);
}

/* ------------------------------------------------------------------ */
/* Multi-epoch / multi-band fit: Fitter.go on an ObsList or MultiBandObsList
 * (results.py:212-347 _set_obs / _fill_gmix_all, :439-466 calc_fdiff,
 * :487-570 calc_jacobian, :142-200 calc_lnprob; get_band_pars :1025-1047;
 * PriorSimpleSep with a list of flux priors, joint_prior.py:20-150).
 * Simple models with the analytic jacobian.  Stamps are ordered band by band,
 * as the reference walks them. */
typedef struct {
    int model, npars, nband, nstamp, n_prior_pars, fdiff_size, totpix;
    fitctx *c;                      /* one per stamp */
    const int32_t *band;            /* band of each stamp */
    const ora_prior *prior;         /* cen, g, T (F unused) */
    const ora_prior1d *Fp;          /* flux prior of each band */
    const ora_bounds *bounds;       /* leastsqbound's transforms, or NULL */
} mfitctx;

/* get_band_pars (results.py:1013-1047): everything up to the fluxes, then
 * the band's flux; ns = 5 (simple), 6 (bdf), 7 (bd) shared parameters */
static void m_band_pars_n(int ns, const double *pars, int band, double *bp) {
    for (int i = 0; i < ns; i++) bp[i] = pars[i];
    bp[ns] = pars[ns + band];
}
static void m_band_pars(const double *pars, int band, double *bp) {
    m_band_pars_n(5, pars, band, bp);
}
static int m_nshared(int model) {
    return model == ORA_BD ? 7 : model == ORA_BDF ? 6 : 5;
}

static int m_fill_gmix_all(mfitctx *m, const double *pars) {
    for (int i = 0; i < m->nstamp; i++) {
        double bp[ORA_MAXP];
        m_band_pars_n(m_nshared(m->model), pars, m->band[i], bp);
        if (ctx_fill_gmix(&m->c[i], bp)) return ORA_RANGE_ERROR;
    }
    return ORA_OK;
}

/* PriorSimpleSep.fill_fdiff with nband flux priors; rows filled or -1 */
static int m_prior_fill_fdiff(const mfitctx *m, const double *pars,
                              double *fdiff) {
    const ora_prior *pr = m->prior;
    if (!pr || !pr->present) return 0;
    int index = 0;
    if (m->model == ORA_BDF || m->model == ORA_BD) {
        /* PriorBDFSep / PriorBDSep.fill_fdiff with a flux prior per band
         * (joint_prior.py:565-599, 364-399): signed rows */
        double sinv1 = 1.0 / pr->sigma1, sinv2 = 1.0 / pr->sigma2;
        fdiff[index++] = (pars[0] - pr->cen1) * sinv1;
        fdiff[index++] = (pars[1] - pr->cen2) * sinv2;
        double lnpg;
        if (gprior_lnprob(pr, pars[2], pars[3], &lnpg)) return -1;
        double c = -2 * lnpg;
        if (c < 0.0) c = 0.0;
        fdiff[index++] = sqrt(c);
        double fd;
        if (prior1d_fdiff(&pr->T, pars[4], &fd)) return -1;
        fdiff[index++] = fd;
        int k = 5;
        if (m->model == ORA_BD) {
            if (prior1d_fdiff(&pr->logTratio, pars[k++], &fd)) return -1;
            fdiff[index++] = fd;
        }
        if (prior1d_fdiff(&pr->fracdev, pars[k++], &fd)) return -1;
        fdiff[index++] = fd;
        for (int b = 0; b < m->nband; b++) {
            if (prior1d_fdiff(&m->Fp[b], pars[k + b], &fd)) return -1;
            fdiff[index++] = fd;
        }
        return index;
    }
    double s2inv1 = 1.0 / (pr->sigma1 * pr->sigma1);
    double s2inv2 = 1.0 / (pr->sigma2 * pr->sigma2);
    double d1 = pr->cen1 - pars[0];
    double d2 = pr->cen2 - pars[1];
    fdiff[index++] = -0.5 * d1 * d1 * s2inv1;
    fdiff[index++] = -0.5 * d2 * d2 * s2inv2;
    double lnp;
    if (gprior_lnprob(pr, pars[2], pars[3], &lnp)) return -1;
    fdiff[index++] = lnp;
    if (prior1d_lnprob(&pr->T, pars[4], &lnp)) return -1;
    fdiff[index++] = lnp;
    for (int b = 0; b < m->nband; b++) {
        if (prior1d_lnprob(&m->Fp[b], pars[5 + b], &lnp)) return -1;
        fdiff[index++] = lnp;
    }
    for (int i = 0; i < index; i++) {
        double c = -2 * fdiff[i];
        if (c < 0.0) c = 0.0;
        fdiff[i] = sqrt(c);
    }
    return index;
}

static int m_prior_lnprob(const mfitctx *m, const double *pars, double *out) {
    const ora_prior *pr = m->prior;
    if (!pr || !pr->present) { *out = 0.0; return ORA_OK; }
    double s2inv1 = 1.0 / (pr->sigma1 * pr->sigma1);
    double s2inv2 = 1.0 / (pr->sigma2 * pr->sigma2);
    double d1 = pr->cen1 - pars[0];
    double d2 = pr->cen2 - pars[1];
    double lnp = -0.5 * d1 * d1 * s2inv1 - 0.5 * d2 * d2 * s2inv2;
    double t;
    if (gprior_lnprob(pr, pars[2], pars[3], &t)) return ORA_RANGE_ERROR;
    lnp += t;
    if (prior1d_lnprob(&pr->T, pars[4], &t)) return ORA_RANGE_ERROR;
    lnp += t;
    int k = 5;
    if (m->model == ORA_BD) {
        if (prior1d_lnprob(&pr->logTratio, pars[k++], &t))
            return ORA_RANGE_ERROR;
        lnp += t;
    }
    if (m->model == ORA_BD || m->model == ORA_BDF) {
        if (prior1d_lnprob(&pr->fracdev, pars[k++], &t))
            return ORA_RANGE_ERROR;
        lnp += t;
    }
    for (int b = 0; b < m->nband; b++) {
        if (prior1d_lnprob(&m->Fp[b], pars[k + b], &t)) return ORA_RANGE_ERROR;
        lnp += t;
    }
    *out = lnp;
    return ORA_OK;
}

static int m_calc_fdiff(mfitctx *m, const double *pars, double *fdiff) {
    int n = m->fdiff_size;
    for (int i = 0; i < n; i++) fdiff[i] = 0.0;
    int bad = 0;
    if (m_fill_gmix_all(m, pars)) bad = 1;
    int start = 0;
    if (!bad) {
        start = m_prior_fill_fdiff(m, pars, fdiff);
        if (start < 0) bad = 1;
    }
    for (int i = 0; i < m->nstamp && !bad; i++) {
        fitctx *c = &m->c[i];
        if (ora_fill_fdiff(c->gm, c->ngauss, c->pixels, c->npix, fdiff, start))
            bad = 1;
        start += c->npix;
    }
    if (bad) {
        for (int i = 0; i < n; i++) fdiff[i] = -INFINITY;
        return ORA_RANGE_ERROR;
    }
    return ORA_OK;
}

static int m_prior_jacobian(mfitctx *m, const double *pars, double *jac,
                            int *nrows) {
    *nrows = 0;
    if (!m->prior || !m->prior->present) return ORA_OK;
    int mm = m->fdiff_size, np = m->npars;
    double f0[ORA_MAXP] = {0}, fs[ORA_MAXP] = {0}, p[ORA_MAXP];
    int n = m_prior_fill_fdiff(m, pars, f0);
    if (n < 0) return ORA_RANGE_ERROR;
    for (int ipar = 0; ipar < np; ipar++) {
        double step = STEP_PRIOR * fmax(1.0, fabs(pars[ipar]));
        for (int k = 0; k < np; k++) p[k] = pars[k];
        p[ipar] = pars[ipar] + step;
        if (m_prior_fill_fdiff(m, p, fs) < 0) {
            step = -step;
            p[ipar] = pars[ipar] + step;
            if (m_prior_fill_fdiff(m, p, fs) < 0) return ORA_RANGE_ERROR;
        }
        for (int k = 0; k < n; k++) {
            double d = 0.0;
            if (isfinite(f0[k]) && isfinite(fs[k])) d = (fs[k] - f0[k]) / step;
            jac[k + (size_t)ipar * mm] = d;
        }
    }
    *nrows = n;
    return ORA_OK;
}

static int m_calc_jacobian(mfitctx *m, const double *pars, double *jac) {
    int mm = m->fdiff_size, np = m->npars;
    memset(jac, 0, sizeof(double) * (size_t)mm * np);
    int start = 0, bad = 0;
    if (m_prior_jacobian(m, pars, jac, &start)) bad = 1;
    if (!bad && m_fill_gmix_all(m, pars)) bad = 1;
    /* the reference walks the bands and raises at the first zero T / flux */
    for (int b = 0; b < m->nband && !bad; b++)
        if (pars[4] == 0.0 || pars[5 + b] == 0.0) bad = 1;
    for (int is = 0; is < m->nstamp && !bad; is++) {
        fitctx *c = &m->c[is];
        double bp[6];
        m_band_pars(pars, m->band[is], bp);
        /* the single-stamp jacobian of this stamp in band parameters */
        int mloc = c->npix;
        double *jl = (double *)calloc((size_t)mloc * 6, sizeof(double));
        fitctx tmp = *c;
        tmp.prior = NULL;
        tmp.n_prior_pars = 0;
        tmp.fdiff_size = mloc;
        tmp.npars = 6;
        if (ctx_calc_jacobian(&tmp, bp, jl)) bad = 1;
        if (!bad) {
            int fluxcol = 5 + m->band[is];
            for (int i = 0; i < mloc; i++) {
                for (int k = 0; k < 5; k++)
                    jac[start + i + (size_t)k * mm] = jl[i + (size_t)k * mloc];
                jac[start + i + (size_t)fluxcol * mm]
                    = jl[i + (size_t)5 * mloc];
            }
        }
        free(jl);
        start += mloc;
    }
    if (bad) {
        memset(jac, 0, sizeof(double) * (size_t)mm * np);
        return ORA_RANGE_ERROR;
    }
    return ORA_OK;
}

/* with bounds MINPACK iterates on internal parameters (leastsqbound.py:
 * 445-552 wfunc / wDfun), exactly as for one stamp above */
static int mcb_fdiff(void *ctx, int mm, int n, const double *x, double *fvec) {
    (void)mm;
    mfitctx *m = (mfitctx *)ctx;
    if (m->bounds) {
        double xe[ORA_MAXP] = {0};
        for (int i = 0; i < n; i++) xe[i] = b_i2e(m->bounds, i, x[i]);
        m_calc_fdiff(m, xe, fvec);
        return 0;
    }
    m_calc_fdiff(m, x, fvec);
    return 0;
}
static int mcb_jac(void *ctx, int mm, int n, const double *x, double *fjac) {
    mfitctx *m = (mfitctx *)ctx;
    if (m->bounds) {
        double xe[ORA_MAXP] = {0};
        for (int i = 0; i < n; i++) xe[i] = b_i2e(m->bounds, i, x[i]);
        m_calc_jacobian(m, xe, fjac);
        for (int j = 0; j < n; j++) {
            double g = b_grad(m->bounds, j, x[j]);
            for (int i = 0; i < mm; i++) fjac[i + (size_t)j * mm] *= g;
        }
        return 0;
    }
    m_calc_jacobian(m, x, fjac);
    return 0;
}

/* Fitter.go on nstamp stamps of one object.  bands[nstamp] non-decreasing,
 * guess/pars have 5 + nband entries, Fp[nband] flux priors (NULL without a
 * prior).  Returns -2 when a stamp has no usable pixel (GMixFatalError). */
int ora_lm_fit_multi(int model, int nstamp, const double *images,
                     const double *weights, int nrow, int ncol,
                     const ora_jac *jacs, const double *psf_pars, int npsf,
                     const int32_t *bands, int nband, const double *guess,
                     const ora_prior *prior, const ora_prior1d *Fp, int maxfev,
                     double ftol, double xtol, int ignore_zero_weight,
                     ora_lm_result *res) {
    return ora_lm_fit_multi2(model, nstamp, images, weights, nrow, ncol, jacs,
                             psf_pars, npsf, bands, nband, guess, prior, Fp,
                             maxfev, ftol, xtol, 1, ignore_zero_weight, res);
}

/* the same for every LM model: gauss / exp / dev with the analytic jacobian
 * (lmder) or, with analytic_jacobian = 0 and always for bdf / bd, MINPACK's
 * forward differences (lmdif), as Fitter.go chooses (fitters.py:93-98) */
int ora_lm_fit_multi2(int model, int nstamp, const double *images,
                      const double *weights, int nrow, int ncol,
                      const ora_jac *jacs, const double *psf_pars, int npsf,
                      const int32_t *bands, int nband, const double *guess,
                      const ora_prior *prior, const ora_prior1d *Fp,
                      int maxfev, double ftol, double xtol,
                      int analytic_jacobian, int ignore_zero_weight,
                      ora_lm_result *res) {
    return ora_lm_fit_multi_bounded(model, nstamp, images, weights, nrow, ncol,
                                    jacs, psf_pars, npsf, bands, nband, guess,
                                    prior, Fp, NULL, maxfev, ftol, xtol,
                                    analytic_jacobian, ignore_zero_weight,
                                    res);
}

/* ... and with prior.bounds (run_leastsq(..., bounds=fit_model.bounds),
 * fitters.py:99-108): NULL or !present = unbounded */
int ora_lm_fit_multi_bounded(int model, int nstamp, const double *images,
                             const double *weights, int nrow, int ncol,
                             const ora_jac *jacs, const double *psf_pars,
                             int npsf, const int32_t *bands, int nband,
                             const double *guess, const ora_prior *prior,
                             const ora_prior1d *Fp, const ora_bounds *bounds,
                             int maxfev, double ftol, double xtol,
                             int analytic_jacobian, int ignore_zero_weight,
                             ora_lm_result *res) {
    memset(res, 0, sizeof(*res));
    if (bounds && !bounds->present) bounds = NULL;
    const int simple = model == ORA_GAUSS || model == ORA_EXP
        || model == ORA_DEV;
    if (!(simple || model == ORA_TURB || model == ORA_BDF || model == ORA_BD))
        return -1;
    const int nsh = m_nshared(model);
    if (nband < 1 || nsh + nband > ORA_MAXP || nstamp < 1) return -1;
    int ntot = nrow * ncol;
    mfitctx m;
    memset(&m, 0, sizeof(m));
    m.model = model; m.nband = nband; m.nstamp = nstamp;
    m.npars = nsh + nband;
    m.band = bands; m.prior = prior; m.Fp = Fp;
    /* results.py:1050-1076: bd 6 + nband, everything else 5 + nband */
    m.n_prior_pars = (prior && prior->present)
        ? (model == ORA_BD ? 6 : 5) + nband : 0;
    m.c = (fitctx *)calloc((size_t)nstamp, sizeof(fitctx));
    ora_pixel **pix = (ora_pixel **)calloc((size_t)nstamp, sizeof(*pix));
    int rc = 0;
    for (int is = 0; is < nstamp; is++) {
        const double *im = images + (size_t)is * ntot;
        const double *wt = weights + (size_t)is * ntot;
        int npixv = ntot;
        if (ignore_zero_weight) {
            npixv = 0;
            for (int i = 0; i < ntot; i++) if (wt[i] > 0.0) npixv++;
        }
        if (npixv == 0) rc = -2;
        pix[is] = (ora_pixel *)malloc(sizeof(ora_pixel) * (size_t)(npixv + 1));
        ora_fill_pixels(pix[is], npixv, im, wt, nrow, ncol, &jacs[is],
                        ignore_zero_weight);
        ctx_init(&m.c[is], model, pix[is], npixv,
                 psf_pars ? psf_pars + (size_t)is * 6 * npsf : NULL, npsf,
                 NULL);
        m.totpix += npixv;
    }
    m.fdiff_size = m.totpix + m.n_prior_pars;
    int np = m.npars, mm = m.fdiff_size;
    if (rc == 0) {
        double x[ORA_MAXP];
        m.bounds = bounds;
        for (int i = 0; i < np; i++)
            x[i] = bounds ? b_e2i(bounds, i, guess[i]) : guess[i];
        double *fvec = (double *)malloc(sizeof(double) * (size_t)mm * (np + 1));
        double *fjac = fvec + mm;
        int ipvt[ORA_MAXP];
        double qtf[ORA_MAXP];
        int nfev = 0, njev = 0;
        int ier;
        if (simple && analytic_jacobian)
            ier = ora_lmder(mcb_fdiff, mcb_jac, &m, mm, np, x, fvec, fjac,
                            ftol, xtol, 0.0, maxfev, 100.0, &nfev, &njev,
                            ipvt, qtf);
        else
            ier = ora_lmdif(mcb_fdiff, &m, mm, np, x, fvec, fjac, ftol, xtol,
                            0.0, maxfev, 2.220446049250313e-16, 100.0, &nfev,
                            ipvt, qtf);
        int flags = 0;
        res->nfev = nfev; res->njev = njev; res->ier = ier;
        if (bounds) {
            /* leastsqbound.py:531-538 with the 1-based ipvt its code assumes */
            for (int j = 0; j < np; j++) {
                double g = b_grad(bounds, ipvt[j], x[ipvt[j]]);
                for (int i = 0; i <= j; i++) fjac[i + (size_t)j * mm] /= g;
            }
            for (int i = 0; i < np; i++) x[i] = b_i2e(bounds, i, x[i]);
            m.bounds = NULL;        /* everything below is at external pars */
        }
        for (int i = 0; i < np; i++) res->pars[i] = x[i];
        for (int i = 0; i < np * np; i++) res->pars_cov0[i] = NAN;
        int have_cov0 = 0;
        if (ier >= 1 && ier <= 4) {
            double rinv[ORA_MAXP * ORA_MAXP];
            int sing = 0;
            for (int j = 0; j < np; j++)
                if (fjac[j + (size_t)j * mm] == 0.0) sing = 1;
            if (!sing) {
                memset(rinv, 0, sizeof(rinv));
                for (int j = 0; j < np; j++) {
                    rinv[j * np + j] = 1.0 / fjac[j + (size_t)j * mm];
                    for (int i = j - 1; i >= 0; i--) {
                        double s = 0.0;
                        for (int k = i + 1; k <= j; k++)
                            s += fjac[i + (size_t)k * mm] * rinv[k * np + j];
                        rinv[i * np + j] = -s / fjac[i + (size_t)i * mm];
                    }
                }
                double prinv[ORA_MAXP * ORA_MAXP];
                for (int j = 0; j < np; j++)
                    for (int k = 0; k < np; k++)
                        prinv[ipvt[j] * np + k] = rinv[j * np + k];
                for (int a = 0; a < np; a++)
                    for (int b = 0; b < np; b++) {
                        double s = 0.0;
                        for (int k = 0; k < np; k++)
                            s += prinv[a * np + k] * prinv[b * np + k];
                        res->pars_cov0[a * np + b] = s;
                    }
                have_cov0 = 1;
            }
        }
        if (ier > 4) {
            flags |= 1 << (ier - 5);
            set_defaults(res, np);
        } else if (!have_cov0) {
            flags |= ORA_LM_SINGULAR_MATRIX;
            for (int i = 0; i < np; i++) res->pars_err[i] = CDEF;
            for (int i = 0; i < np * np; i++) res->pars_cov[i] = CDEF;
        } else {
            m_calc_fdiff(&m, res->pars, fvec);
            int dof = mm - m.n_prior_pars - np;
            if (dof == 0) {
                for (int i = 0; i < np; i++) res->pars_err[i] = CDEF;
                for (int i = 0; i < np * np; i++) res->pars_cov[i] = CDEF;
                flags |= ORA_ZERO_DOF;
            } else {
                double ssum = 0.0;
                for (int i = m.n_prior_pars; i < mm; i++)
                    ssum += fvec[i] * fvec[i];
                double s_sq = ssum / dof;
                for (int i = 0; i < np * np; i++)
                    res->pars_cov[i] = res->pars_cov0[i] * s_sq;
                int cflags = 0;
                double ev[ORA_MAXP];
                if (sym_eigvals(np, res->pars_cov, ev)) {
                    cflags |= ORA_EIG_NOTFINITE;
                } else {
                    for (int i = 0; i < np; i++)
                        if (ev[i] < 0) cflags |= ORA_LM_NEG_COV_EIG;
                    for (int i = 0; i < np; i++)
                        if (res->pars_cov[i * np + i] < 0)
                            cflags |= ORA_LM_NEG_COV_DIAG;
                }
                if (cflags) {
                    flags |= cflags;
                    for (int i = 0; i < np; i++) res->pars_err[i] = CDEF;
                } else {
                    for (int i = 0; i < np; i++)
                        res->pars_err[i] = sqrt(res->pars_cov[i * np + i]);
                }
            }
        }
        res->flags = flags;
        res->lnprob = NAN; res->s2n_numer = NAN; res->s2n_denom = NAN;
        res->chi2per = NAN; res->s2n = NAN; res->T = NAN; res->T_err = NAN;
        res->flux = NAN; res->flux_err = NAN;
        if (flags == 0) {
            double lnp_prior = 0.0;
            int bad = 0;
            /* calc_lnprob: priors first, then the mixtures */
            if (m_prior_lnprob(&m, res->pars, &lnp_prior)) bad = 1;
            if (!bad && m_fill_gmix_all(&m, res->pars)) bad = 1;
            double lnprob = 0.0, num = 0.0, den = 0.0;
            int npix = 0;
            for (int is = 0; is < nstamp && !bad; is++) {
                double out4[4];
                fitctx *c = &m.c[is];
                if (ora_get_loglike(c->gm, c->ngauss, c->pixels, c->npix,
                                    out4)) {
                    bad = 1;
                    break;
                }
                lnprob += out4[0]; num += out4[1]; den += out4[2];
                npix += (int)out4[3];
            }
            if (bad) {
                res->lnprob = -INFINITY; res->s2n_numer = 0.0;
                res->s2n_denom = BIGVAL; res->npix = 0;
            } else {
                res->lnprob = lnprob + lnp_prior;
                res->s2n_numer = num; res->s2n_denom = den; res->npix = npix;
            }
            double s2n = 0.0;
            if (res->s2n_denom > 0)
                s2n = res->s2n_numer / sqrt(res->s2n_denom);
            int dof = res->npix - np;
            res->chi2per = (res->lnprob / (-0.5)) / dof;
            res->dof = dof;
            res->s2n = s2n;
            res->T = res->pars[4];
            res->T_err = sqrt(res->pars_cov[4 * np + 4]);
            /* band 0; the caller slices pars / pars_cov for more bands */
            res->flux = res->pars[nsh];
            res->flux_err = sqrt(res->pars_cov[nsh * np + nsh]);
        }
        free(fvec);
    }
    for (int is = 0; is < nstamp; is++) {
        ctx_free(&m.c[is]);
        free(pix[is]);
    }
    free(pix);
    free(m.c);
    return rc;
}

/* ---------------- threaded batch driver (cpu_baseline) ---------------- */
typedef struct {
    int lo, hi;
    int model, nrow, ncol, npsf, maxfev, analytic;
    const double *images, *weights, *jacs, *psf, *guess;
    const ora_prior *prior;
    double ftol, xtol;
    ora_lm_result *res;
} lm_job;

static void *lm_worker(void *arg) {
    lm_job *j = (lm_job *)arg;
    int np = ora_model_npars(j->model);
    size_t npx = (size_t)j->nrow * j->ncol;
    for (int i = j->lo; i < j->hi; i++) {
        ora_lm_fit(j->model, j->images + i * npx, j->weights + i * npx,
                   j->nrow, j->ncol, (const ora_jac *)(j->jacs + 8 * (size_t)i),
                   j->psf ? j->psf + (size_t)i * 6 * j->npsf : NULL, j->npsf,
                   j->guess + (size_t)i * np, j->prior, j->maxfev, j->ftol,
                   j->xtol, j->analytic, 1, &j->res[i]);
    }
    return NULL;
}

int ora_lm_fit_batch(int n, int nthreads, int model, const double *images,
                     const double *weights, int nrow, int ncol,
                     const double *jacs, const double *psf_pars, int npsf,
                     const double *guess, const ora_prior *prior, int maxfev,
                     double ftol, double xtol, int analytic_jacobian,
                     ora_lm_result *res) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > n) nthreads = n;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    lm_job *jobs = (lm_job *)malloc(sizeof(lm_job) * nthreads);
    for (int t = 0; t < nthreads; t++) {
        lm_job *j = &jobs[t];
        j->lo = (int)((long long)n * t / nthreads);
        j->hi = (int)((long long)n * (t + 1) / nthreads);
        j->model = model; j->nrow = nrow; j->ncol = ncol; j->npsf = npsf;
        j->maxfev = maxfev; j->analytic = analytic_jacobian;
        j->images = images; j->weights = weights; j->jacs = jacs;
        j->psf = psf_pars; j->guess = guess; j->prior = prior;
        j->ftol = ftol; j->xtol = xtol; j->res = res;
        pthread_create(&th[t], NULL, lm_worker, j);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
    return 0;
}
