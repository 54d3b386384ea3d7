/*
 * TEST INFRASTRUCTURE (see oracle.h).  CPU restatement of
 *   em_run + variants     ngmix/em/em_nb.py:15-1257
 *   EMFitter.go           ngmix/em/em.py:228-317, prep_image :95-120
 *   admom                 ngmix/admom/admom_nb.py:13-243
 *   AdmomFitter.go        ngmix/admom/admom.py:325-360
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

#define GMIX_LOW_DETVAL 1.0e-200

static void gset(ora_gauss *g, double p, double row, double col, double irr,
                 double irc, double icc) {
    g->norm_set = 0;
    g->drr = NAN; g->drc = NAN; g->dcc = NAN; g->norm = NAN; g->pnorm = NAN;
    g->p = p; g->row = row; g->col = col;
    g->irr = irr; g->irc = irc; g->icc = icc;
    g->det = irr * icc - irc * irc;
}

/* gmix_get_moms, em_nb.py:1260-1294 */
static void gmix_get_moms(const ora_gauss *gmix, int n, double *irr,
                          double *irc, double *icc) {
    double row = 0.0, col = 0.0, psum = 0.0;
    for (int i = 0; i < n; i++) {
        double p = gmix[i].p;
        row += p * gmix[i].row;
        col += p * gmix[i].col;
        psum += p;
    }
    row /= psum;
    col /= psum;
    double a = 0.0, b = 0.0, c = 0.0;
    for (int i = 0; i < n; i++) {
        double rowdiff = gmix[i].row - row;
        double coldiff = gmix[i].col - col;
        double p = gmix[i].p;
        a += p * (gmix[i].irr + rowdiff * rowdiff);
        b += p * (gmix[i].irc + rowdiff * coldiff);
        c += p * (gmix[i].icc + coldiff * coldiff);
    }
    *irr = a / psum; *irc = b / psum; *icc = c / psum;
}

typedef struct {
    double gi, tvsum, tusum, tu2sum, tuvsum, tv2sum;
    double pnew, vsum, usum, u2sum, uvsum, v2sum;
} em_sums;

/*
 * one function for the four variants; the variant only selects which sums
 * exist and which fields the M step updates.
 */
int ora_em_run(int variant, const ora_em_conf *conf, ora_pixel *pixels,
               int npix, ora_gauss *gmix, int ngauss, ora_gauss *gmix_psf,
               int npsf, ora_gauss *gmix_conv, int fill_zero_weight,
               int32_t *numiter_out, double *frac_diff_out, double *sky_out) {
    int nconv = ngauss * npsf;
    if (ora_gmix_set_norms(gmix_conv, nconv)) return ORA_RANGE_ERROR;
    em_sums *sums = (em_sums *)calloc((size_t)ngauss, sizeof(em_sums));
    double *logtau = (double *)malloc(sizeof(double) * 2 * (size_t)nconv);
    double *logdet = logtau + nconv;
    double tol = conf->tol;
    double sky = conf->sky;
    double elogL_last = -9999.9e9;
    double frac_diff = NAN;
    int numiter = 0;
    int status = ORA_OK;
    double p_last = 0.0;
    if (variant == ORA_EM_FLUXONLY)
        for (int i = 0; i < ngauss; i++) p_last += gmix[i].p;

    for (int it = 0; it < conf->maxiter; it++) {
        double elogL = 0.0, skysum = 0.0;
        memset(sums, 0, sizeof(em_sums) * (size_t)ngauss);
        if (variant != ORA_EM_FLUXONLY)
            for (int i = 0; i < nconv; i++) {      /* set_logtau_logdet */
                logtau[i] = log(gmix_conv[i].p);
                logdet[i] = log(gmix_conv[i].det);
            }
        if (fill_zero_weight)                       /* em_nb.py:1297-1315 */
            for (int ip = 0; ip < npix; ip++)
                if (pixels[ip].ierr <= 0.0)
                    pixels[ip].val = sky + ora_gmix_eval_fast(
                        gmix_conv, nconv, pixels[ip].v, pixels[ip].u,
                        pixels[ip].area);

        for (int ip = 0; ip < npix; ip++) {
            const ora_pixel *px = &pixels[ip];
            double v = px->v, u = px->u;
            double gsum = 0.0, logL = 0.0;
            /* do_scratch_sums*, em_nb.py:160-246 */
            for (int ii = 0; ii < ngauss; ii++) {
                em_sums *ts = &sums[ii];
                ts->gi = 0.0; ts->tvsum = 0.0; ts->tusum = 0.0;
                ts->tv2sum = 0.0; ts->tuvsum = 0.0; ts->tu2sum = 0.0;
                for (int i = ii * npsf; i < (ii + 1) * npsf; i++) {
                    const ora_gauss *g = &gmix_conv[i];
                    double vdiff = v - g->row;
                    double udiff = u - g->col;
                    double u2 = udiff * udiff;
                    double v2 = vdiff * vdiff;
                    double uv = udiff * vdiff;
                    double chi2 = g->dcc * v2 + g->drr * u2
                        - 2.0 * g->drc * uv;
                    double val;
                    if (chi2 < 25.0 && chi2 >= 0.0)
                        val = g->pnorm * ora_fexp(-0.5 * chi2) * px->area;
                    else
                        val = 0.0;
                    ts->gi += val;
                    gsum += val;
                    if (variant == ORA_EM_FULL || variant == ORA_EM_FIXCOV) {
                        ts->tvsum += v * val;
                        ts->tusum += u * val;
                    }
                    if (variant == ORA_EM_FULL || variant == ORA_EM_FIXCEN) {
                        ts->tv2sum += v2 * val;
                        ts->tuvsum += uv * val;
                        ts->tu2sum += u2 * val;
                    }
                    if (variant != ORA_EM_FLUXONLY)
                        logL += val * (logtau[i] - 0.5 * logdet[i]
                                       - 0.5 * chi2);
                }
            }
            if (variant != ORA_EM_FLUXONLY) {
                if (gsum == 0.0) logL = 0.0;
                else logL *= 1.0 / gsum;
            }
            double gtot = gsum + sky;
            if (gtot == 0.0) { status = ORA_RANGE_ERROR; goto done; }
            elogL += logL;
            skysum += sky * px->val / gtot;
            /* do_sums*, em_nb.py:249-281 */
            double factor = px->val / gtot;
            for (int i = 0; i < ngauss; i++) {
                em_sums *ts = &sums[i];
                double wtau = ts->gi * factor;
                ts->pnew += wtau;
                if (variant == ORA_EM_FULL || variant == ORA_EM_FIXCOV) {
                    ts->usum += ts->tusum * factor;
                    ts->vsum += ts->tvsum * factor;
                }
                if (variant == ORA_EM_FULL || variant == ORA_EM_FIXCEN) {
                    ts->u2sum += ts->tu2sum * factor;
                    ts->uvsum += ts->tuvsum * factor;
                    ts->v2sum += ts->tv2sum * factor;
                }
            }
        }

        /* gmix_set_from_sums*, em_nb.py:284-354, 592-655, 923-972, 1173-1216 */
        {
            double minval = 1.0e-4;
            double psf_irr = 0, psf_irc = 0, psf_icc = 0;
            if (variant == ORA_EM_FULL || variant == ORA_EM_FIXCEN)
                gmix_get_moms(gmix_psf, npsf, &psf_irr, &psf_irc, &psf_icc);
            for (int i = 0; i < ngauss; i++) {
                em_sums *ts = &sums[i];
                ora_gauss *g = &gmix[i];
                double p = ts->pnew;
                /* numba raises ZeroDivisionError on 1.0/0.0, which
                 * EMFitter.go maps to EM_RANGE_ERROR (em.py:307-315) */
                if (p == 0.0 && variant != ORA_EM_FLUXONLY) {
                    status = ORA_ZERO_DIV;
                    goto done;
                }
                double pinv = 1.0 / p;
                double v = g->row, u = g->col;
                double irr = g->irr, irc = g->irc, icc = g->icc;
                if (variant == ORA_EM_FULL || variant == ORA_EM_FIXCOV) {
                    v = ts->vsum * pinv;
                    u = ts->usum * pinv;
                }
                if (variant == ORA_EM_FULL || variant == ORA_EM_FIXCEN) {
                    irr = ts->v2sum * pinv;
                    irc = ts->uvsum * pinv;
                    icc = ts->u2sum * pinv;
                    irr = irr - psf_irr;
                    irc = irc - psf_irc;
                    icc = icc - psf_icc;
                    if (irr < 0.0 || icc < 0.0) {
                        irr = minval; irc = 0.0; icc = minval;
                    }
                    double det = irr * icc - irc * irc;
                    if (det < GMIX_LOW_DETVAL) {
                        double T = irr + icc;
                        irr = icc = T / 2;
                        irc = 0.0;
                    }
                }
                gset(g, p, v, u, irr, irc, icc);
            }
            ora_gmix_convolve_fill(gmix_conv, gmix, ngauss, gmix_psf, npsf);
            if (ora_gmix_set_norms(gmix_conv, nconv)) {
                status = ORA_RANGE_ERROR;
                goto done;
            }
        }
        if (conf->vary_sky) sky = skysum / npix;

        numiter = it + 1;
        if (variant == ORA_EM_FLUXONLY) {
            double psum = 0.0;
            for (int i = 0; i < ngauss; i++) psum += gmix[i].p;
            if (numiter >= conf->miniter) {
                frac_diff = fabs(psum / p_last - 1);
                if (frac_diff < tol) break;
            }
            p_last = psum;
        } else {
            if (numiter >= conf->miniter) {
                if (elogL == 0.0) { status = ORA_RANGE_ERROR; goto done; }
                frac_diff = fabs((elogL - elogL_last) / elogL);
                if (frac_diff < tol) break;
            }
            elogL_last = elogL;
        }
    }
    for (int i = 0; i < ngauss; i++) gmix[i].norm_set = 0;
done:
    *numiter_out = numiter;
    *frac_diff_out = frac_diff;
    *sky_out = sky;
    free(sums);
    free(logtau);
    return status;
}

/* EMFitter.go, em.py:228-317 */
int ora_em_go(int variant, const double *image, const double *weight, int nrow,
              int ncol, const ora_jac *jac, const double *guess_pars,
              int ngauss, const double *psf_pars, int npsf, double tol,
              int miniter, int maxiter, int vary_sky, int have_sky,
              double sky_in, int ignore_zero_weight, double *pars_out,
              ora_em_result *res) {
    int ntot = nrow * ncol;
    double *im = (double *)malloc(sizeof(double) * (size_t)ntot);
    memcpy(im, image, sizeof(double) * (size_t)ntot);
    double sky = sky_in;
    if (!have_sky) {                       /* prep_image, em.py:95-120 */
        double im_min = im[0], im_max = im[0];
        for (int i = 1; i < ntot; i++) {
            if (im[i] < im_min) im_min = im[i];
            if (im[i] > im_max) im_max = im[i];
        }
        double desired_minval = 0.001 * (im_max - im_min);
        sky = desired_minval - im_min;
        for (int i = 0; i < ntot; i++) im[i] += sky;
    }
    int npixv = ntot;
    if (ignore_zero_weight) {
        npixv = 0;
        for (int i = 0; i < ntot; i++) if (weight[i] > 0.0) npixv++;
    }
    ora_pixel *pixels = (ora_pixel *)malloc(sizeof(ora_pixel) * (size_t)ntot);
    ora_fill_pixels(pixels, npixv, im, weight, nrow, ncol, jac,
                    ignore_zero_weight);

    ora_gauss gpsf[16], gm[16];
    int npsf_eff = npsf;
    if (npsf <= 0) {                       /* delta psf, em.py:254-256 */
        double dp[6] = {0., 0., 0., 0., 0., 1.0};
        ora_gmix_fill(ORA_GAUSS, dp, 6, gpsf, 1);
        npsf_eff = 1;
    } else {
        ora_gmix_fill(ORA_FULL, psf_pars, 6 * npsf, gpsf, npsf);
        double psum0 = 0.0;                /* set_flux(1.0), gmix.py:299-311 */
        for (int i = 0; i < npsf; i++) psum0 += gpsf[i].p;
        double rat = 1.0 / psum0;
        for (int i = 0; i < npsf; i++) gpsf[i].p *= rat;
    }
    ora_gmix_fill(ORA_FULL, guess_pars, 6 * ngauss, gm, ngauss);
    ora_gauss *gconv = (ora_gauss *)calloc((size_t)ngauss * npsf_eff,
                                           sizeof(ora_gauss));
    ora_gmix_convolve_fill(gconv, gm, ngauss, gpsf, npsf_eff);

    ora_em_conf conf;
    conf.tol = tol; conf.maxiter = maxiter; conf.miniter = miniter;
    conf.sky = sky; conf.vary_sky = vary_sky; conf.pad = 0;
    int fill_zero_weight = 0;
    for (int i = 0; i < npixv; i++)
        if (pixels[i].ierr <= 0.0) fill_zero_weight = 1;

    int32_t numiter = 0;
    double fdiff = NAN, sky_out = sky;
    int st = ora_em_run(variant, &conf, pixels, npixv, gm, ngauss, gpsf,
                        npsf_eff, gconv, fill_zero_weight, &numiter, &fdiff,
                        &sky_out);
    res->flags = 0;
    res->numiter = numiter;
    res->fdiff = fdiff;
    res->sky = sky_out;
    if (st) {
        res->flags = ORA_GMIX_RANGE_ERROR;
    } else {
        if (numiter >= maxiter) res->flags = ORA_MAXITER;
        for (int i = 0; i < ngauss; i++) {
            double *q = pars_out + 6 * i;
            q[0] = gm[i].p; q[1] = gm[i].row; q[2] = gm[i].col;
            q[3] = gm[i].irr; q[4] = gm[i].irc; q[5] = gm[i].icc;
        }
    }
    free(gconv);
    free(pixels);
    free(im);
    return st;
}

/* ------------------------------------------------------------------ */
/* admom_nb.py:229-239 */
static void clear_result(ora_admom_result *res) {
    res->npix = 0;
    res->wsum = 0.0;
    for (int i = 0; i < 7; i++) res->sums[i] = 0.0;
    for (int i = 0; i < 49; i++) res->sums_cov[i] = 0.0;
    for (int i = 0; i < 6; i++) res->pars[i] = NAN;
    res->rho4 = NAN;
}

/* admom_nb.py:178-226 */
static void deweight_moments(ora_gauss *wt, double Irr, double Irc, double Icc,
                             ora_admom_result *res) {
    double detm = Irr * Icc - Irc * Irc;
    if (detm <= GMIX_LOW_DETVAL) { res->flags = ORA_LOW_DET; return; }
    double Wrr = wt->irr, Wrc = wt->irc, Wcc = wt->icc;
    double detw = Wrr * Wcc - Wrc * Wrc;
    if (detw <= GMIX_LOW_DETVAL) { res->flags = ORA_LOW_DET; return; }
    double idetw = 1.0 / detw;
    double idetm = 1.0 / detm;
    double Nrr = Icc * idetm - Wcc * idetw;
    double Ncc = Irr * idetm - Wrr * idetw;
    double Nrc = -Irc * idetm + Wrc * idetw;
    double detn = Nrr * Ncc - Nrc * Nrc;
    if (detn <= GMIX_LOW_DETVAL) { res->flags = ORA_LOW_DET; return; }
    double idetn = 1. / detn;
    wt->irr = Ncc * idetn;
    wt->icc = Nrr * idetn;
    wt->irc = -Nrc * idetn;
    wt->det = wt->irr * wt->icc - wt->irc * wt->irc;
}

/* admom_nb.py:13-108 */
void ora_admom(const ora_admom_conf *conf, ora_gauss *wt,
               const ora_pixel *pixels, int npix, ora_admom_result *res) {
    double roworig = wt->row, colorig = wt->col;
    double e1old = NAN, e2old = NAN, Told = NAN;
    int i = 0;
    for (i = 0; i < conf->maxiter; i++) {
        if (wt->det < GMIX_LOW_DETVAL) { res->flags = ORA_LOW_DET; break; }
        ora_gmix_set_norms(wt, 1);
        clear_result(res);
        /* admom_censums :111-128 */
        for (int ip = 0; ip < npix; ip++) {
            const ora_pixel *px = &pixels[ip];
            double weight = ora_gmix_eval_fast(wt, 1, px->v, px->u, px->area);
            double wdata = weight * px->val;
            res->npix += 1;
            res->sums[0] += wdata * px->v;
            res->sums[1] += wdata * px->u;
            res->sums[5] += wdata;
        }
        if (res->sums[5] <= 0.0) { res->flags = ORA_NONPOS_FLUX; break; }
        wt->row = res->sums[0] / res->sums[5];
        wt->col = res->sums[1] / res->sums[5];
        if (fabs(wt->row - roworig) > conf->shiftmax
            || fabs(wt->col - colorig) > conf->shiftmax) {
            res->flags = ORA_CEN_SHIFT;
            break;
        }
        clear_result(res);
        /* admom_momsums :131-175 */
        {
            double vcen = wt->row, ucen = wt->col;
            double *F = res->F;
            for (int ip = 0; ip < npix; ip++) {
                const ora_pixel *px = &pixels[ip];
                double weight = ora_gmix_eval_fast(wt, 1, px->v, px->u,
                                                   px->area);
                double var = 1.0 / (px->ierr * px->ierr);
                double vmod = px->v - vcen;
                double umod = px->u - ucen;
                double wdata = weight * px->val;
                double w2 = weight * weight;
                double chi2 = wt->dcc * vmod * vmod + wt->drr * umod * umod
                    - 2.0 * wt->drc * vmod * umod;
                F[0] = px->v;
                F[1] = px->u;
                F[2] = umod * umod - vmod * vmod;
                F[3] = 2 * vmod * umod;
                F[4] = umod * umod + vmod * vmod;
                F[5] = 1.0;
                F[6] = chi2 * chi2;
                res->wsum += weight;
                res->npix += 1;
                for (int a = 0; a < 7; a++) {
                    res->sums[a] += wdata * F[a];
                    for (int b = 0; b < 7; b++)
                        res->sums_cov[a * 7 + b] += w2 * var * F[a] * F[b];
                }
            }
        }
        if (res->sums[5] <= 0.0) { res->flags = ORA_NONPOS_FLUX; break; }
        double finv = 1.0 / res->sums[5];
        double M1 = res->sums[2] * finv;
        double M2 = res->sums[3] * finv;
        double T = res->sums[4] * finv;
        double Irr = 0.5 * (T - M1);
        double Icc = 0.5 * (T + M1);
        double Irc = 0.5 * M2;
        if (T <= 0.0) { res->flags = ORA_NONPOS_SIZE; break; }
        double e1 = (Icc - Irr) / T;
        double e2 = 2 * Irc / T;
        if (fabs(e1 - e1old) < conf->etol && fabs(e2 - e2old) < conf->etol
            && fabs(T / Told - 1.) < conf->Ttol) {
            res->pars[0] = wt->row;
            res->pars[1] = wt->col;
            res->pars[2] = wt->icc - wt->irr;
            res->pars[3] = 2.0 * wt->irc;
            res->pars[4] = wt->icc + wt->irr;
            res->pars[5] = 1.0;
            res->rho4 = res->sums[6] / res->sums[5];
            break;
        } else {
            if (!conf->cenonly) {
                deweight_moments(wt, Irr, Irc, Icc, res);
                if (res->flags != 0) break;
            }
            e1old = e1; e2old = e2; Told = T;
        }
    }
    /* python leaves i at maxiter-1 when the loop is exhausted */
    if (i == conf->maxiter) i = conf->maxiter - 1;
    res->numiter = i + 1;
    if (res->numiter == conf->maxiter) res->flags = ORA_MAXITER;
}

/* AdmomFitter.go on raw arrays, admom.py:325-360 (guess is a GMix) */
int ora_admom_go(const double *image, const double *weight, int nrow,
                 int ncol, const ora_jac *jac, const double *guess_pars,
                 const ora_admom_conf *conf, ora_admom_result *res,
                 double *wt_norm_out) {
    int ntot = nrow * ncol;
    int npixv = 0;
    for (int i = 0; i < ntot; i++) if (weight[i] > 0.0) npixv++;
    if (npixv == 0) return -2;
    ora_pixel *pixels = (ora_pixel *)malloc(sizeof(ora_pixel) * (size_t)ntot);
    ora_fill_pixels(pixels, npixv, image, weight, nrow, ncol, jac, 1);
    ora_gauss wt;
    ora_gmix_fill(ORA_FULL, guess_pars, 6, &wt, 1);
    memset(res, 0, sizeof(*res));
    ora_admom(conf, &wt, pixels, npixv, res);
    if (wt_norm_out) *wt_norm_out = wt.norm;
    free(pixels);
    return 0;
}

/* ---------------- threaded batch drivers (cpu_baseline) ---------------- */
typedef struct {
    int lo, hi, kind;
    int variant, nrow, ncol, ngauss, miniter, maxiter;
    double tol;
    const double *images, *weights, *jacs, *guess;
    double *pars_out;
    ora_em_result *em_res;
    const ora_admom_conf *am_conf;
    ora_admom_result *am_res;
} ea_job;

static void *ea_worker(void *arg) {
    ea_job *j = (ea_job *)arg;
    size_t npx = (size_t)j->nrow * j->ncol;
    for (int i = j->lo; i < j->hi; i++) {
        const ora_jac *jac = (const ora_jac *)(j->jacs + 8 * (size_t)i);
        if (j->kind == 0)
            ora_em_go(j->variant, j->images + i * npx, j->weights + i * npx,
                      j->nrow, j->ncol, jac,
                      j->guess + (size_t)i * 6 * j->ngauss, j->ngauss, NULL, 0,
                      j->tol, j->miniter, j->maxiter, 0, 0, 0.0, 1,
                      j->pars_out + (size_t)i * 6 * j->ngauss, &j->em_res[i]);
        else
            ora_admom_go(j->images + i * npx, j->weights + i * npx, j->nrow,
                         j->ncol, jac, j->guess + (size_t)i * 6, j->am_conf,
                         &j->am_res[i], NULL);
    }
    return NULL;
}

static int run_jobs(int n, int nthreads, ea_job *tmpl) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > n) nthreads = n;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    ea_job *jobs = (ea_job *)malloc(sizeof(ea_job) * nthreads);
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = *tmpl;
        jobs[t].lo = (int)((long long)n * t / nthreads);
        jobs[t].hi = (int)((long long)n * (t + 1) / nthreads);
        pthread_create(&th[t], NULL, ea_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
    return 0;
}

int ora_em_go_batch(int n, int nthreads, int variant, const double *images,
                    const double *weights, int nrow, int ncol,
                    const double *jacs, const double *guess_pars, int ngauss,
                    double tol, int miniter, int maxiter, double *pars_out,
                    ora_em_result *res) {
    ea_job j;
    memset(&j, 0, sizeof(j));
    j.kind = 0; j.variant = variant; j.nrow = nrow; j.ncol = ncol;
    j.ngauss = ngauss; j.miniter = miniter; j.maxiter = maxiter; j.tol = tol;
    j.images = images; j.weights = weights; j.jacs = jacs;
    j.guess = guess_pars; j.pars_out = pars_out; j.em_res = res;
    return run_jobs(n, nthreads, &j);
}

int ora_admom_go_batch(int n, int nthreads, const double *images,
                       const double *weights, int nrow, int ncol,
                       const double *jacs, const double *guess_pars,
                       const ora_admom_conf *conf, ora_admom_result *res) {
    ea_job j;
    memset(&j, 0, sizeof(j));
    j.kind = 1; j.nrow = nrow; j.ncol = ncol;
    j.images = images; j.weights = weights; j.jacs = jacs;
    j.guess = guess_pars; j.am_conf = conf; j.am_res = res;
    return run_jobs(n, nthreads, &j);
}
