/*
 * TEST INFRASTRUCTURE (see oracle.h).  CPU restatement of the per-pixel
 * Gaussian-mixture arithmetic of the reference:
 *   ngmix/fastexp_nb.py, ngmix/gmix/gmix_nb.py, ngmix/gmix/render_nb.py,
 *   ngmix/pixels/pixels_nb.py, ngmix/jacobian/jacobian_nb.py,
 *   ngmix/fitting/derivs_nb.py
 * Scalar fp64, evaluated in the reference's own operation order; built with
 * -ffp-contract=off because numba emits no fused multiply-adds.
 */
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include "oracle.h"

#define GMIX_LOW_DETVAL 1.0e-200     /* gmix_nb.py:11 */
#define FASTEXP_MAX_CHI2 25.0        /* fastexp_nb.py:80 */
#define FASTEXP_APOD_CHI2 20.0       /* fastexp_nb.py:85 */
#define APOD_IWIDTH (1.0 / (FASTEXP_MAX_CHI2 - FASTEXP_APOD_CHI2))

/* exp(i), i = -15..0: numpy.exp of the integers (fastexp_nb.py:5-16,90-94).
 * Filled at first use with libm exp, which is correctly rounded for these
 * arguments on glibc; the table is checked bit-for-bit against the
 * reference's _EXP_LOOKUP in tests/test_oracle_golden.py. */
static double EXP_LOOKUP[16];
static int exp_lookup_set = 0;
static void set_lookup(void) {
    if (!exp_lookup_set) {
        for (int i = 0; i < 16; i++) EXP_LOOKUP[i] = exp((double)(i - 15));
        exp_lookup_set = 1;
    }
}

/* exp5_smooth, fastexp_nb.py:223-262 */
double ora_fexp(double x) {
    set_lookup();
    int ival = (int)(x - 0.5);        /* truncation toward zero, as numba */
    double f = x - ival;
    int index = ival - (-15);
    double expval = EXP_LOOKUP[index];
    expval *= 1.0000011318561302 + f * (
        0.999993601071577 + f * (
            0.49992478810274166 + f * (
                0.16674612720799442 + f * (
                    0.042330947141114836 + f * 0.008197933236258961))));
    return expval;
}

/* fastexp_nb.py:97-117 */
double ora_apod_window(double chi2) {
    double u = (FASTEXP_MAX_CHI2 - chi2) * APOD_IWIDTH;
    return u * u * u * (10.0 + u * (-15.0 + 6.0 * u));
}

/* fastexp_nb.py:120-135 */
double ora_apod_window_deriv(double chi2) {
    double u = (FASTEXP_MAX_CHI2 - chi2) * APOD_IWIDTH;
    double umu = u * (1.0 - u);
    return -30.0 * umu * umu * APOD_IWIDTH;
}

void ora_fexp_arr(int n, const double *x, double *out) {
    for (int i = 0; i < n; i++) out[i] = ora_fexp(x[i]);
}

void ora_apod_arr(int n, const double *chi2, double *w, double *dw) {
    for (int i = 0; i < n; i++) {
        w[i] = ora_apod_window(chi2[i]);
        dw[i] = ora_apod_window_deriv(chi2[i]);
    }
}

/* ------------------------------------------------------------------ */
/* tables, gmix_nb.py:243-304 */
static const double PVALS_EXP[6] = {
    0.00061601229677880041, 0.0079461395724623237, 0.053280454055540001,
    0.21797364640726541, 0.45496740582554868, 0.26521634184240478};
static const double FVALS_EXP[6] = {
    0.002467115141477932, 0.018147435573256168, 0.07944063151366336,
    0.27137669897479122, 0.79782256866993773, 2.1623306025075739};
static const double PVALS_DEV[10] = {
    6.5288960012625658e-05, 0.00044199216814302695, 0.0020859587871659754,
    0.0075913681418996841, 0.02260266219257237, 0.056532254390212859,
    0.11939049233042602, 0.20969545753234975, 0.29254151133139222,
    0.28905301416582552};
static const double FVALS_DEV[10] = {
    2.9934935706271918e-07, 3.4651596338231207e-06, 2.4807910570562753e-05,
    1.4307404300535354e-04, 7.2753169298239500e-04, 3.4582464394427260e-03,
    1.6086645440719100e-02, 7.7006776775654429e-02, 4.1012562102501476e-01,
    2.9812509778548648e00};
static const double PVALS_TURB[3] = {
    0.596510042804182, 0.4034898268889178, 1.303069003078001e-07};
static const double FVALS_TURB[3] = {
    0.5793612389470884, 1.621860687127999, 7.019347162356363};
static const double PVALS_GAUSS[1] = {1.0};
static const double FVALS_GAUSS[1] = {1.0};

int ora_model_ngauss(int model) {
    switch (model) {
    case ORA_GAUSS: return 1;
    case ORA_TURB: return 3;
    case ORA_EXP: return 6;
    case ORA_DEV: return 10;
    case ORA_BDF: case ORA_BD: case ORA_CM: return 16;
    default: return -1;
    }
}

int ora_model_npars(int model) {
    switch (model) {
    case ORA_GAUSS: case ORA_TURB: case ORA_EXP: case ORA_DEV: case ORA_CM:
        return 6;
    case ORA_BDF: return 7;
    case ORA_BD: return 8;
    default: return -1;
    }
}

/* gmix_nb.py:221-240 */
static void gauss2d_set(ora_gauss *g, double p, double row, double col,
                        double irr, double irc, double icc) {
    g->norm_set = 0;
    g->drr = NAN; g->drc = NAN; g->dcc = NAN; g->norm = NAN; g->pnorm = NAN;
    g->p = p; g->row = row; g->col = col;
    g->irr = irr; g->irc = irc; g->icc = icc;
    g->det = irr * icc - irc * irc;
}

/* gmix_nb.py:652-678 */
int ora_g1g2_to_e1e2(double g1, double g2, double *e1, double *e2) {
    double g = sqrt(g1 * g1 + g2 * g2);
    if (g >= 1) return ORA_RANGE_ERROR;
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
    return ORA_OK;
}

/* gmix_nb.py:307-340 */
static int fill_simple(ora_gauss *gmix, int n, const double *pars,
                       const double *fvals, const double *pvals) {
    double row = pars[0], col = pars[1], g1 = pars[2], g2 = pars[3];
    double T = pars[4], flux = pars[5];
    double e1, e2;
    if (ora_g1g2_to_e1e2(g1, g2, &e1, &e2)) return ORA_RANGE_ERROR;
    for (int i = 0; i < n; i++) {
        double T_i_2 = 0.5 * T * fvals[i];
        double flux_i = flux * pvals[i];
        gauss2d_set(&gmix[i], flux_i, row, col, T_i_2 * (1 - e1), T_i_2 * e2,
                    T_i_2 * (1 + e1));
    }
    return ORA_OK;
}

/* gmix_nb.py:561-593 */
double ora_get_cm_Tfactor(double fracdev, double TdByTe) {
    double ifracdev = 1.0 - fracdev;
    double Tfactor = 0.0;
    for (int i = 0; i < 6; i++) {
        double p = PVALS_EXP[i] * ifracdev;
        double f = FVALS_EXP[i];
        Tfactor += p * f;
    }
    for (int i = 0; i < 10; i++) {
        double p = PVALS_DEV[i] * fracdev;
        double f = FVALS_DEV[i] * TdByTe;
        Tfactor += p * f;
    }
    return 1.0 / Tfactor;
}

/* shared body of gmix_fill_cm / bd / bdf, gmix_nb.py:430-558 */
static int fill_composite(ora_gauss *gmix, double row, double col, double g1,
                          double g2, double T, double flux, double fracdev,
                          double TdByTe) {
    double ifracdev = 1.0 - fracdev;
    double e1, e2;
    if (ora_g1g2_to_e1e2(g1, g2, &e1, &e2)) return ORA_RANGE_ERROR;
    for (int i = 0; i < 16; i++) {
        double p, f;
        if (i < 6) {
            p = PVALS_EXP[i] * ifracdev;
            f = FVALS_EXP[i];
        } else {
            p = PVALS_DEV[i - 6] * fracdev;
            f = FVALS_DEV[i - 6] * TdByTe;
        }
        double T_i_2 = 0.5 * T * f;
        double flux_i = flux * p;
        gauss2d_set(&gmix[i], flux_i, row, col, T_i_2 * (1 - e1), T_i_2 * e2,
                    T_i_2 * (1 + e1));
    }
    return ORA_OK;
}

int ora_gmix_fill_cm(double fracdev, double TdByTe, double Tfactor,
                     const double *pars, ora_gauss *gmix) {
    return fill_composite(gmix, pars[0], pars[1], pars[2], pars[3],
                          pars[4] * Tfactor, pars[5], fracdev, TdByTe);
}

int ora_gmix_fill(int model, const double *pars, int npars, ora_gauss *gmix,
                  int ngauss) {
    switch (model) {
    case ORA_GAUSS: return fill_simple(gmix, 1, pars, FVALS_GAUSS, PVALS_GAUSS);
    case ORA_TURB: return fill_simple(gmix, 3, pars, FVALS_TURB, PVALS_TURB);
    case ORA_EXP: return fill_simple(gmix, 6, pars, FVALS_EXP, PVALS_EXP);
    case ORA_DEV: return fill_simple(gmix, 10, pars, FVALS_DEV, PVALS_DEV);
    case ORA_BDF: {               /* gmix_nb.py:515-558 */
        double TdByTe = 1.0;
        double fracdev = pars[5];
        double Tfactor = ora_get_cm_Tfactor(fracdev, TdByTe);
        double T = pars[4] * Tfactor;
        return fill_composite(gmix, pars[0], pars[1], pars[2], pars[3], T,
                              pars[6], fracdev, TdByTe);
    }
    case ORA_BD: {                /* gmix_nb.py:469-512 */
        double TdByTe = pow(10.0, pars[5]);
        double fracdev = pars[6];
        double Tfactor = ora_get_cm_Tfactor(fracdev, TdByTe);
        double T = pars[4] * Tfactor;
        return fill_composite(gmix, pars[0], pars[1], pars[2], pars[3], T,
                              pars[7], fracdev, TdByTe);
    }
    case ORA_COELLIP: {           /* gmix_nb.py:375-405 */
        double e1, e2;
        if (ora_g1g2_to_e1e2(pars[2], pars[3], &e1, &e2))
            return ORA_RANGE_ERROR;
        for (int i = 0; i < ngauss; i++) {
            double T = pars[4 + i];
            double Thalf = 0.5 * T;
            double flux = pars[4 + ngauss + i];
            gauss2d_set(&gmix[i], flux, pars[0], pars[1], Thalf * (1 - e1),
                        Thalf * e2, Thalf * (1 + e1));
        }
        return ORA_OK;
    }
    case ORA_FULL:                /* gmix_nb.py:408-427 */
        for (int i = 0; i < ngauss; i++) {
            const double *q = pars + 6 * i;
            gauss2d_set(&gmix[i], q[0], q[1], q[2], q[3], q[4], q[5]);
        }
        return ORA_OK;
    default:
        (void)npars;
        return -1;
    }
}

/* gmix_nb.py:176-218 */
int ora_gmix_set_norms(ora_gauss *gmix, int n) {
    for (int i = 0; i < n; i++) {
        ora_gauss *g = &gmix[i];
        if (g->det < GMIX_LOW_DETVAL) return ORA_RANGE_ERROR;
        double T = g->irr + g->icc;
        if (T <= GMIX_LOW_DETVAL) return ORA_RANGE_ERROR;
        double idet = 1.0 / g->det;
        g->drr = g->irr * idet;
        g->drc = g->irc * idet;
        g->dcc = g->icc * idet;
        g->norm = 1.0 / (2 * M_PI * sqrt(g->det));
        g->pnorm = g->p * g->norm;
        g->norm_set = 1;
    }
    return ORA_OK;
}

/* gmix_nb.py:108-130 */
static void gmix_get_cen(const ora_gauss *gmix, int n, double *row,
                         double *col, double *psum) {
    double r = 0.0, c = 0.0, ps = 0.0;
    for (int i = 0; i < n; i++) {
        double p = gmix[i].p;
        r += p * gmix[i].row;
        c += p * gmix[i].col;
        ps += p;
    }
    *row = r / ps; *col = c / ps; *psum = ps;
}

/* gmix_nb.py:609-649 */
void ora_gmix_convolve_fill(ora_gauss *self, const ora_gauss *gmix, int n,
                            const ora_gauss *psf, int npsf) {
    double prow, pcol, ppsum;
    gmix_get_cen(psf, npsf, &prow, &pcol, &ppsum);
    double psf_ipsum = 1.0 / ppsum;
    int itot = 0;
    for (int iobj = 0; iobj < n; iobj++) {
        const ora_gauss *o = &gmix[iobj];
        for (int ipsf = 0; ipsf < npsf; ipsf++) {
            const ora_gauss *q = &psf[ipsf];
            double p = o->p * q->p * psf_ipsum;
            double row = o->row + (q->row - prow);
            double col = o->col + (q->col - pcol);
            double irr = o->irr + q->irr;
            double irc = o->irc + q->irc;
            double icc = o->icc + q->icc;
            gauss2d_set(&self[itot], p, row, col, irr, irc, icc);
            itot++;
        }
    }
}

/* ------------------------------------------------------------------ */
/* jacobian_nb.py:4-16 and pixels_nb.py:6-58 */
int ora_fill_pixels(ora_pixel *pixels, int npixels, const double *image,
                    const double *weight, int nrow, int ncol,
                    const ora_jac *jac, int ignore_zero_weight) {
    double pixel_area = jac->scale * jac->scale;
    int ipixel = 0;
    for (int row = 0; row < nrow; row++) {
        for (int col = 0; col < ncol; col++) {
            double ivar = weight[row * ncol + col];
            if (ignore_zero_weight && ivar <= 0.0) continue;
            if (ipixel >= npixels) return -1;
            ora_pixel *pix = &pixels[ipixel];
            double rowdiff = row - jac->row0;
            double coldiff = col - jac->col0;
            pix->v = jac->dvdrow * rowdiff + jac->dvdcol * coldiff;
            pix->u = jac->dudrow * rowdiff + jac->dudcol * coldiff;
            pix->area = pixel_area;
            pix->val = image[row * ncol + col];
            if (ivar < 0.0) ivar = 0.0;
            pix->ierr = sqrt(ivar);
            pix->fdiff = 0.0;
            ipixel++;
        }
    }
    return ipixel;
}

/* pixels_nb.py:61-94 */
void ora_fill_coords(ora_coord *coords, int nrow, int ncol,
                     const ora_jac *jac) {
    double pixel_area = jac->scale * jac->scale;
    int ic = 0;
    for (int row = 0; row < nrow; row++) {
        for (int col = 0; col < ncol; col++) {
            double rowdiff = row - jac->row0;
            double coldiff = col - jac->col0;
            coords[ic].v = jac->dvdrow * rowdiff + jac->dvdcol * coldiff;
            coords[ic].u = jac->dudrow * rowdiff + jac->dudcol * coldiff;
            coords[ic].area = pixel_area;
            ic++;
        }
    }
}

/* ------------------------------------------------------------------ */
/* gmix_nb.py:28-63 */
static inline double gauss_eval_fast(const ora_gauss *g, double v, double u,
                                     double area) {
    double model_val = 0.0;
    double vdiff = v - g->row;
    double udiff = u - g->col;
    double chi2 = g->dcc * vdiff * vdiff + g->drr * udiff * udiff
        - 2.0 * g->drc * vdiff * udiff;
    if (chi2 < FASTEXP_MAX_CHI2 && chi2 >= 0.0) {
        model_val = g->pnorm * ora_fexp(-0.5 * chi2) * area;
        if (chi2 > FASTEXP_APOD_CHI2) model_val *= ora_apod_window(chi2);
    }
    return model_val;
}

/* gmix_nb.py:66-92 */
static inline double gauss_eval_exact(const ora_gauss *g, double v, double u,
                                      double area) {
    double vdiff = v - g->row;
    double udiff = u - g->col;
    double chi2 = g->dcc * vdiff * vdiff + g->drr * udiff * udiff
        - 2.0 * g->drc * vdiff * udiff;
    return g->pnorm * exp(-0.5 * chi2) * area;
}

/* gmix_nb.py:14-25; exported for em/admom */
double ora_gmix_eval_fast(const ora_gauss *gmix, int n, double v, double u,
                          double area) {
    double s = 0.0;
    for (int i = 0; i < n; i++) s += gauss_eval_fast(&gmix[i], v, u, area);
    return s;
}

/* render_nb.py:9-36 */
int ora_render(ora_gauss *gmix, int n, const ora_coord *coords, int ncoords,
               double *image, int fast_exp) {
    if (gmix[0].norm_set == 0)
        if (ora_gmix_set_norms(gmix, n)) return ORA_RANGE_ERROR;
    if (fast_exp) {
        for (int i = 0; i < ncoords; i++)
            image[i] += ora_gmix_eval_fast(gmix, n, coords[i].v, coords[i].u,
                                           coords[i].area);
    } else {
        for (int i = 0; i < ncoords; i++) {
            double s = 0.0;
            for (int k = 0; k < n; k++)
                s += gauss_eval_exact(&gmix[k], coords[i].v, coords[i].u,
                                      coords[i].area);
            image[i] += s;
        }
    }
    return ORA_OK;
}

/* gmix_nb.py:877-900 */
int ora_fill_fdiff(ora_gauss *gmix, int n, const ora_pixel *pixels, int npix,
                   double *fdiff, int start) {
    if (gmix[0].norm_set == 0)
        if (ora_gmix_set_norms(gmix, n)) return ORA_RANGE_ERROR;
    for (int i = 0; i < npix; i++) {
        const ora_pixel *px = &pixels[i];
        double model_val = ora_gmix_eval_fast(gmix, n, px->v, px->u, px->area);
        fdiff[start + i] = (model_val - px->val) * px->ierr;
    }
    return ORA_OK;
}

/* gmix_nb.py:824-874 */
int ora_get_loglike(ora_gauss *gmix, int n, const ora_pixel *pixels, int npix,
                    double *out4) {
    if (gmix[0].norm_set == 0)
        if (ora_gmix_set_norms(gmix, n)) return ORA_RANGE_ERROR;
    double loglike = 0.0, s2n_numer = 0.0, s2n_denom = 0.0;
    int np = 0;
    for (int i = 0; i < npix; i++) {
        const ora_pixel *px = &pixels[i];
        double model_val = ora_gmix_eval_fast(gmix, n, px->v, px->u, px->area);
        double ivar = px->ierr * px->ierr;
        double val = px->val;
        double diff = model_val - val;
        loglike += diff * diff * ivar;
        s2n_numer += val * model_val * ivar;
        s2n_denom += model_val * model_val * ivar;
        np += 1;
    }
    loglike *= -0.5;
    out4[0] = loglike; out4[1] = s2n_numer; out4[2] = s2n_denom;
    out4[3] = (double)np;
    return ORA_OK;
}

/* gmix_nb.py:903-937 */
int ora_get_model_s2n_sum(ora_gauss *gmix, int n, const ora_pixel *pixels,
                          int npix, double *s2n_sum) {
    if (gmix[0].norm_set == 0)
        if (ora_gmix_set_norms(gmix, n)) return ORA_RANGE_ERROR;
    double s = 0.0;
    for (int i = 0; i < npix; i++) {
        const ora_pixel *px = &pixels[i];
        double model_val = ora_gmix_eval_fast(gmix, n, px->v, px->u, px->area);
        double ivar = px->ierr * px->ierr;
        s += model_val * model_val * ivar;
    }
    *s2n_sum = s;
    return ORA_OK;
}

/* derivs_nb.py:40-127; out is (6, npix) row major, zeroed by the caller */
void ora_deriv_images(const double *gpars, const double *dcov, int ngauss,
                      const double *vv, const double *uu, const double *area,
                      int npix, double *out) {
    const double TWO_PI = 2.0 * M_PI;
    double trs[3];
    for (int ig = 0; ig < ngauss; ig++) {
        const double *gp = gpars + 6 * ig;
        const double *dc = dcov + 9 * ig;
        double p = gp[0], vcen = gp[1], ucen = gp[2];
        double irr = gp[3], irc = gp[4], icc = gp[5];
        double det = irr * icc - irc * irc;
        if (det <= 0.0) continue;
        double norm = p / (TWO_PI * sqrt(det));
        double w11 = icc / det;
        double w12 = -irc / det;
        double w22 = irr / det;
        for (int a = 0; a < 3; a++)
            trs[a] = w11 * dc[3 * a + 0] + 2.0 * w12 * dc[3 * a + 1]
                + w22 * dc[3 * a + 2];
        for (int ipix = 0; ipix < npix; ipix++) {
            double dv = vv[ipix] - vcen;
            double du = uu[ipix] - ucen;
            double qv = w11 * dv + w12 * du;
            double qu = w12 * dv + w22 * du;
            double chi2 = dv * qv + du * qu;
            if (chi2 >= FASTEXP_MAX_CHI2 || chi2 < 0.0) continue;
            double val = norm * ora_fexp(-0.5 * chi2) * area[ipix];
            double valc;
            if (chi2 > FASTEXP_APOD_CHI2) {
                double w = ora_apod_window(chi2);
                valc = val * (w - 2.0 * ora_apod_window_deriv(chi2));
                val *= w;
            } else {
                valc = val;
            }
            out[0 * npix + ipix] += val;
            out[1 * npix + ipix] += valc * qv;
            out[2 * npix + ipix] += valc * qu;
            for (int a = 0; a < 3; a++) {
                double quad = qv * qv * dc[3 * a + 0]
                    + 2.0 * qv * qu * dc[3 * a + 1]
                    + qu * qu * dc[3 * a + 2];
                out[(3 + a) * npix + ipix] += 0.5 * (valc * quad - val * trs[a]);
            }
        }
    }
}

/* ------------------------------------------------------------------ */
/* weighted moments, gmix_nb.py:681-821.  The weight is the mixture evaluated
 * with the exact exponential (gmix_eval_pixel, gmix_nb.py:95-105). */
static double gmix_eval_exact(const ora_gauss *gmix, int n, double v,
                              double u, double area) {
    double s = 0.0;
    for (int k = 0; k < n; k++) s += gauss_eval_exact(&gmix[k], v, u, area);
    return s;
}

int ora_weighted_sums(ora_gauss *wt, int ngauss, const ora_pixel *pixels,
                      int npix, double maxrad, int nmom, ora_mom_result *res) {
    if (wt[0].norm_set == 0)
        if (ora_gmix_set_norms(wt, ngauss)) return ORA_RANGE_ERROR;
    double maxrad2 = maxrad * maxrad;
    double vcen = wt[0].row, ucen = wt[0].col;
    double F[ORA_NMOM_MAX];
    for (int ip = 0; ip < npix; ip++) {
        const ora_pixel *px = &pixels[ip];
        double v = px->v - vcen, u = px->u - ucen;
        double r2 = u * u + v * v;
        /* :708 requires ierr > 0 as well; the higher-order version :773 does
         * not */
        if (!(r2 < maxrad2)) continue;
        if (nmom == 6 && !(px->ierr > 0.0)) continue;
        double weight = gmix_eval_exact(wt, ngauss, px->v, px->u, px->area);
        double var = 1.0 / (px->ierr * px->ierr);
        double wdata = weight * px->val;
        double w2 = weight * weight;
        F[0] = px->v;
        F[1] = px->u;
        if (nmom == 6) {
            F[2] = u * u - v * v;
            F[3] = 2 * v * u;
            F[4] = r2;
            F[5] = 1.0;
        } else {
            double u2 = u * u, v2 = v * v, vu = v * u;
            double u4 = u2 * u2, v4 = v2 * v2;
            double r4 = r2 * r2, r6 = r4 * r2, r8 = r6 * r2;
            F[2] = u2 - v2;
            F[3] = 2 * vu;
            F[4] = r2;
            F[5] = 1.0;
            F[6] = u * r2;
            F[7] = v * r2;
            F[8] = u * (u2 - 3 * v2);
            F[9] = v * (3 * u2 - v2);
            F[10] = r4;
            F[11] = r2 * (u2 - v2);
            F[12] = r2 * 2 * u * v;
            F[13] = u4 - 6 * u2 * v2 + v4;
            F[14] = (u2 - v2) * 4 * u * v;
            F[15] = r6;
            F[16] = r8;
        }
        res->wsum += weight;
        res->npix += 1;
        for (int i = 0; i < nmom; i++) {
            res->sums[i] += wdata * F[i];
            for (int j = 0; j < nmom; j++)
                res->sums_cov[i * nmom + j] += w2 * var * F[i] * F[j];
        }
    }
    return ORA_OK;
}

int ora_weighted_sums_go(const double *image, const double *weight, int nrow,
                         int ncol, const ora_jac *jac, const double *wt_pars,
                         int ngauss, double maxrad, int nmom,
                         int ignore_zero_weight, ora_mom_result *res) {
    int ntot = nrow * ncol;
    ora_pixel *pixels = (ora_pixel *)malloc(sizeof(ora_pixel) * (size_t)ntot);
    ora_gauss *wt = (ora_gauss *)malloc(sizeof(ora_gauss) * (size_t)ngauss);
    int npix = ora_fill_pixels(pixels, ntot, image, weight, nrow, ncol, jac,
                               ignore_zero_weight);
    memset(res, 0, sizeof(*res));
    int st = ora_gmix_fill(ORA_FULL, wt_pars, 6 * ngauss, wt, ngauss);
    if (st == ORA_OK)
        st = ora_weighted_sums(wt, ngauss, pixels, npix, maxrad, nmom, res);
    free(pixels);
    free(wt);
    return st;
}

/* ------------------------------------------------------------------ */
/* PSFFluxFitModel.go (fitting/results.py:702-770): zero-lag
 * cross-correlation of the exact-exp render of the template mixture with the
 * image, over ALL pixels (zero weights drop out through wt), then chi^2 at
 * the fitted flux.  totpix counts obs.pixels (after the zero-weight cut),
 * the dof counts weight > 0 (get_effective_npix). */
int ora_template_flux(int nstamp, const double *images, const double *weights,
                      int nrow, int ncol, const ora_jac *jacs,
                      const double *gm_pars, int ngauss, int normalize,
                      int ignore_zero_weight, ora_flux_result *res) {
    int ntot = nrow * ncol;
    ora_gauss *gm = (ora_gauss *)malloc(sizeof(ora_gauss) * (size_t)ngauss);
    double *norms = (double *)malloc(sizeof(double) * (size_t)nstamp);
    double xcorr = 0.0, msq = 0.0, chi2 = 0.0;
    int totpix = 0, effpix = 0, flags = 0;
    double flux = -9.999e9, flux_err = 9.999e9;
    memset(res, 0, sizeof(*res));
    for (int ipass = 1; ipass <= 2; ipass++) {
        for (int is = 0; is < nstamp; is++) {
            const double *im = images + (size_t)is * ntot;
            const double *wt = weights + (size_t)is * ntot;
            const ora_jac *jac = &jacs[is];
            double pars[6 * 16];
            for (int i = 0; i < 6 * ngauss; i++)
                pars[i] = gm_pars[(size_t)is * 6 * ngauss + i];
            double psum = 0.0;
            for (int k = 0; k < ngauss; k++) psum += pars[6 * k];
            if (normalize) {
                /* GMix.set_flux(1.0), gmix.py:299-311 */
                double rat = 1.0 / psum;
                for (int k = 0; k < ngauss; k++) pars[6 * k] *= rat;
            }
            if (ipass == 1) {
                double nsum = 0.0;
                for (int k = 0; k < ngauss; k++) nsum += pars[6 * k];
                norms[is] = nsum;
            } else {
                double cur = 0.0;
                for (int k = 0; k < ngauss; k++) cur += pars[6 * k];
                double rat = (flux * norms[is]) / cur;
                for (int k = 0; k < ngauss; k++) pars[6 * k] *= rat;
            }
            if (ora_gmix_fill(ORA_FULL, pars, 6 * ngauss, gm, ngauss)
                || ora_gmix_set_norms(gm, ngauss)) {
                free(gm); free(norms);
                return ORA_RANGE_ERROR;
            }
            double area = jac->scale * jac->scale;
            double s1 = 0.0, s2 = 0.0, s3 = 0.0;
            for (int row = 0; row < nrow; row++)
                for (int col = 0; col < ncol; col++) {
                    double rowdiff = row - jac->row0, coldiff = col - jac->col0;
                    double v = jac->dvdrow * rowdiff + jac->dvdcol * coldiff;
                    double u = jac->dudrow * rowdiff + jac->dudcol * coldiff;
                    double model = 0.0;
                    for (int k = 0; k < ngauss; k++)
                        model += gauss_eval_exact(&gm[k], v, u, area);
                    double d = im[row * ncol + col], w = wt[row * ncol + col];
                    if (ipass == 1) {
                        s1 += model * d * w;
                        s2 += model * model * w;
                        if (!(ignore_zero_weight && w <= 0.0)) totpix++;
                        if (w > 0) effpix++;
                    } else {
                        double diff = model - d;
                        s3 += diff * diff * w;
                    }
                }
            if (ipass == 1) { xcorr += s1; msq += s2; }
            else chi2 += s3;
        }
        if (ipass == 1) {
            if (msq == 0) break;
            flux = xcorr / msq;
        }
    }
    double dof = effpix - 1;
    if (dof <= 0) dof = 1.0e-6;
    double chi2per = 9999.0;
    if (dof > 0) chi2per = chi2 / dof;
    else flags |= (1 << 15);
    if (msq == 0 || totpix == 1) {
        flags |= (1 << 14);                     /* DIV_ZERO */
    } else {
        double arg = chi2 / msq / (totpix - 1);
        if (arg >= 0.0) flux_err = sqrt(arg);
        else flags |= (1 << 6);                 /* BAD_VAR */
    }
    res->flags = flags;
    res->flux = flux; res->flux_err = flux_err;
    res->chi2per = chi2per; res->dof = dof;
    free(gm); free(norms);
    return ORA_OK;
}
