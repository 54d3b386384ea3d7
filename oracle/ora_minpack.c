/*
 * TEST INFRASTRUCTURE (see oracle.h).  Restatement of MINPACK-1 lmder /
 * lmdif, the Levenberg-Marquardt driver the reference reaches through
 * scipy.optimize.leastsq -> scipy.optimize._minpack._lmder/_lmdif
 * (ngmix/fitting/leastsqbound.py:17,445-447,475,491).
 *
 * The arithmetic lives in scipy's binary (scipy 1.18.1 in the build
 * container; the reference pins no version, setup.py:13-19) and is NOT in
 * /root/reference.  This file restates the published MINPACK-1 algorithm
 * (More, Garbow, Hillstrom 1980: enorm, qrfac, qrsolv, lmpar, fdjac2, lmder,
 * lmdif) and is pinned against live scipy runs: tests/golden/lm_*.npz hold
 * (ier, nfev, x, cov_x) of the reference's own fits and
 * tests/test_oracle_golden.py requires identical ier/nfev and x to 1e-12.
 *
 * Matrices are column major with leading dimension m, ipvt is 0-based.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

#define EPSMCH DBL_EPSILON           /* dpmpar(1) */
#define DWARF DBL_MIN                /* dpmpar(2) */

static double dmax(double a, double b) { return a >= b ? a : b; }
static double dmin(double a, double b) { return a <= b ? a : b; }

/* euclidean norm with the three-accumulator scaling of MINPACK enorm */
double ora_enorm(int n, const double *x) {
    const double rdwarf = 3.834e-20, rgiant = 1.304e19;
    double s1 = 0.0, s2 = 0.0, s3 = 0.0, x1max = 0.0, x3max = 0.0;
    double agiant = rgiant / (double)n;
    for (int i = 0; i < n; i++) {
        double xabs = fabs(x[i]);
        if (xabs > rdwarf && xabs < agiant) {
            s2 += xabs * xabs;
        } else if (xabs <= rdwarf) {
            if (xabs > x3max) {
                double r = x3max / xabs;
                s3 = 1.0 + s3 * r * r;
                x3max = xabs;
            } else if (xabs != 0.0) {
                double r = xabs / x3max;
                s3 += r * r;
            }
        } else {
            if (xabs > x1max) {
                double r = x1max / xabs;
                s1 = 1.0 + s1 * r * r;
                x1max = xabs;
            } else {
                double r = xabs / x1max;
                s1 += r * r;
            }
        }
    }
    if (s1 != 0.0) return x1max * sqrt(s1 + (s2 / x1max) / x1max);
    if (s2 != 0.0) {
        if (s2 >= x3max)
            return sqrt(s2 * (1.0 + (x3max / s2) * (x3max * s3)));
        return sqrt(x3max * ((s2 / x3max) + (x3max * s3)));
    }
    return x3max * sqrt(s3);
}

/* Householder QR with column pivoting */
static void qrfac(int m, int n, double *a, int lda, int pivot, int *ipvt,
                  double *rdiag, double *acnorm, double *wa) {
    for (int j = 0; j < n; j++) {
        acnorm[j] = ora_enorm(m, a + (size_t)j * lda);
        rdiag[j] = acnorm[j];
        wa[j] = rdiag[j];
        if (pivot) ipvt[j] = j;
    }
    int minmn = m < n ? m : n;
    for (int j = 0; j < minmn; j++) {
        if (pivot) {
            int kmax = j;
            for (int k = j; k < n; k++)
                if (rdiag[k] > rdiag[kmax]) kmax = k;
            if (kmax != j) {
                for (int i = 0; i < m; i++) {
                    double t = a[i + (size_t)j * lda];
                    a[i + (size_t)j * lda] = a[i + (size_t)kmax * lda];
                    a[i + (size_t)kmax * lda] = t;
                }
                rdiag[kmax] = rdiag[j];
                wa[kmax] = wa[j];
                int k = ipvt[j]; ipvt[j] = ipvt[kmax]; ipvt[kmax] = k;
            }
        }
        double ajnorm = ora_enorm(m - j, a + j + (size_t)j * lda);
        if (ajnorm != 0.0) {
            if (a[j + (size_t)j * lda] < 0.0) ajnorm = -ajnorm;
            for (int i = j; i < m; i++) a[i + (size_t)j * lda] /= ajnorm;
            a[j + (size_t)j * lda] += 1.0;
            for (int k = j + 1; k < n; k++) {
                double sum = 0.0;
                for (int i = j; i < m; i++)
                    sum += a[i + (size_t)j * lda] * a[i + (size_t)k * lda];
                double temp = sum / a[j + (size_t)j * lda];
                for (int i = j; i < m; i++)
                    a[i + (size_t)k * lda] -= temp * a[i + (size_t)j * lda];
                if (pivot && rdiag[k] != 0.0) {
                    temp = a[j + (size_t)k * lda] / rdiag[k];
                    rdiag[k] *= sqrt(dmax(0.0, 1.0 - temp * temp));
                    double r = rdiag[k] / wa[k];
                    if (0.05 * (r * r) <= EPSMCH) {
                        rdiag[k] = ora_enorm(m - j - 1,
                                             a + j + 1 + (size_t)k * lda);
                        wa[k] = rdiag[k];
                    }
                }
            }
        }
        rdiag[j] = -ajnorm;
    }
}

static void qrsolv(int n, double *r, int ldr, const int *ipvt,
                   const double *diag, const double *qtb, double *x,
                   double *sdiag, double *wa) {
    for (int j = 0; j < n; j++) {
        for (int i = j; i < n; i++)
            r[i + (size_t)j * ldr] = r[j + (size_t)i * ldr];
        x[j] = r[j + (size_t)j * ldr];
        wa[j] = qtb[j];
    }
    for (int j = 0; j < n; j++) {
        int l = ipvt[j];
        if (diag[l] != 0.0) {
            for (int k = j; k < n; k++) sdiag[k] = 0.0;
            sdiag[j] = diag[l];
            double qtbpj = 0.0;
            for (int k = j; k < n; k++) {
                if (sdiag[k] == 0.0) continue;
                double c, s;
                double rkk = r[k + (size_t)k * ldr];
                if (fabs(rkk) < fabs(sdiag[k])) {
                    double cotan = rkk / sdiag[k];
                    s = 0.5 / sqrt(0.25 + 0.25 * (cotan * cotan));
                    c = s * cotan;
                } else {
                    double tn = sdiag[k] / rkk;
                    c = 0.5 / sqrt(0.25 + 0.25 * (tn * tn));
                    s = c * tn;
                }
                r[k + (size_t)k * ldr] = c * rkk + s * sdiag[k];
                double temp = c * wa[k] + s * qtbpj;
                qtbpj = -s * wa[k] + c * qtbpj;
                wa[k] = temp;
                for (int i = k + 1; i < n; i++) {
                    temp = c * r[i + (size_t)k * ldr] + s * sdiag[i];
                    sdiag[i] = -s * r[i + (size_t)k * ldr] + c * sdiag[i];
                    r[i + (size_t)k * ldr] = temp;
                }
            }
        }
        sdiag[j] = r[j + (size_t)j * ldr];
        r[j + (size_t)j * ldr] = x[j];
    }
    int nsing = n;
    for (int j = 0; j < n; j++) {
        if (sdiag[j] == 0.0 && nsing == n) nsing = j;
        if (nsing < n) wa[j] = 0.0;
    }
    for (int k = 1; k <= nsing; k++) {
        int j = nsing - k;
        double sum = 0.0;
        for (int i = j + 1; i < nsing; i++)
            sum += r[i + (size_t)j * ldr] * wa[i];
        wa[j] = (wa[j] - sum) / sdiag[j];
    }
    for (int j = 0; j < n; j++) x[ipvt[j]] = wa[j];
}

static void lmpar(int n, double *r, int ldr, const int *ipvt,
                  const double *diag, const double *qtb, double delta,
                  double *par, double *x, double *sdiag, double *wa1,
                  double *wa2) {
    int nsing = n;
    for (int j = 0; j < n; j++) {
        wa1[j] = qtb[j];
        if (r[j + (size_t)j * ldr] == 0.0 && nsing == n) nsing = j;
        if (nsing < n) wa1[j] = 0.0;
    }
    for (int k = 1; k <= nsing; k++) {
        int j = nsing - k;
        wa1[j] /= r[j + (size_t)j * ldr];
        double temp = wa1[j];
        for (int i = 0; i < j; i++) wa1[i] -= r[i + (size_t)j * ldr] * temp;
    }
    for (int j = 0; j < n; j++) x[ipvt[j]] = wa1[j];

    int iter = 0;
    for (int j = 0; j < n; j++) wa2[j] = diag[j] * x[j];
    double dxnorm = ora_enorm(n, wa2);
    double fp = dxnorm - delta;
    if (fp <= 0.1 * delta) {
        *par = 0.0;
        return;
    }
    double parl = 0.0;
    if (nsing >= n) {
        for (int j = 0; j < n; j++) {
            int l = ipvt[j];
            wa1[j] = diag[l] * (wa2[l] / dxnorm);
        }
        for (int j = 0; j < n; j++) {
            double sum = 0.0;
            for (int i = 0; i < j; i++) sum += r[i + (size_t)j * ldr] * wa1[i];
            wa1[j] = (wa1[j] - sum) / r[j + (size_t)j * ldr];
        }
        double temp = ora_enorm(n, wa1);
        parl = ((fp / delta) / temp) / temp;
    }
    for (int j = 0; j < n; j++) {
        double sum = 0.0;
        for (int i = 0; i <= j; i++) sum += r[i + (size_t)j * ldr] * qtb[i];
        wa1[j] = sum / diag[ipvt[j]];
    }
    double gnorm = ora_enorm(n, wa1);
    double paru = gnorm / delta;
    if (paru == 0.0) paru = DWARF / dmin(delta, 0.1);
    *par = dmax(*par, parl);
    *par = dmin(*par, paru);
    if (*par == 0.0) *par = gnorm / dxnorm;
    for (;;) {
        iter++;
        if (*par == 0.0) *par = dmax(DWARF, 0.001 * paru);
        double temp = sqrt(*par);
        for (int j = 0; j < n; j++) wa1[j] = temp * diag[j];
        qrsolv(n, r, ldr, ipvt, wa1, qtb, x, sdiag, wa2);
        for (int j = 0; j < n; j++) wa2[j] = diag[j] * x[j];
        dxnorm = ora_enorm(n, wa2);
        temp = fp;
        fp = dxnorm - delta;
        if (fabs(fp) <= 0.1 * delta
            || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10)
            break;
        for (int j = 0; j < n; j++) {
            int l = ipvt[j];
            wa1[j] = diag[l] * (wa2[l] / dxnorm);
        }
        for (int j = 0; j < n; j++) {
            wa1[j] /= sdiag[j];
            temp = wa1[j];
            for (int i = j + 1; i < n; i++)
                wa1[i] -= r[i + (size_t)j * ldr] * temp;
        }
        temp = ora_enorm(n, wa1);
        double parc = ((fp / delta) / temp) / temp;
        if (fp > 0.0) parl = dmax(parl, *par);
        if (fp < 0.0) paru = dmin(paru, *par);
        *par = dmax(parl, *par + parc);
    }
}

/* forward-difference jacobian */
static int fdjac2(ora_fcn fcn, void *ctx, int m, int n, double *x,
                  const double *fvec, double *fjac, double epsfcn,
                  double *wa) {
    double eps = sqrt(dmax(epsfcn, EPSMCH));
    for (int j = 0; j < n; j++) {
        double temp = x[j];
        double h = eps * fabs(temp);
        if (h == 0.0) h = eps;
        x[j] = temp + h;
        int st = fcn(ctx, m, n, x, wa);
        x[j] = temp;
        if (st < 0) return st;
        for (int i = 0; i < m; i++)
            fjac[i + (size_t)j * m] = (wa[i] - fvec[i]) / h;
    }
    return 0;
}

/*
 * shared driver: jfcn != NULL -> lmder, else lmdif with fdjac2.
 * mode = 1 (internal scaling), nprint = 0, as scipy.leastsq(diag=None).
 */
static int lm_driver(ora_fcn fcn, ora_jacfcn jfcn, void *ctx, int m, int n,
                     double *x, double *fvec, double *fjac, double ftol,
                     double xtol, double gtol, int maxfev, double epsfcn,
                     double factor, int *nfev, int *njev, int *ipvt,
                     double *qtf) {
    int info = 0;
    *nfev = 0;
    *njev = 0;
    if (n <= 0 || m < n || ftol < 0.0 || xtol < 0.0 || gtol < 0.0
        || maxfev <= 0 || factor <= 0.0)
        return 0;
    double *diag = (double *)malloc(sizeof(double) * (5 * n + m));
    double *wa1 = diag + n, *wa2 = wa1 + n, *wa3 = wa2 + n, *wa4 = wa3 + n;
    double *sdiag_unused = NULL;
    (void)sdiag_unused;

    if (fcn(ctx, m, n, x, fvec) < 0) { free(diag); return -1; }
    *nfev = 1;
    double fnorm = ora_enorm(m, fvec);
    double par = 0.0, delta = 0.0, xnorm = 0.0, gnorm = 0.0;
    int iter = 1;

    for (;;) {
        if (jfcn) {
            if (jfcn(ctx, m, n, x, fjac) < 0) { info = -1; break; }
            *njev += 1;
        } else {
            if (fdjac2(fcn, ctx, m, n, x, fvec, fjac, epsfcn, wa4) < 0) {
                info = -1;
                break;
            }
            *nfev += n;
        }
        qrfac(m, n, fjac, m, 1, ipvt, wa1, wa2, wa3);
        if (iter == 1) {
            for (int j = 0; j < n; j++) {
                diag[j] = wa2[j];
                if (wa2[j] == 0.0) diag[j] = 1.0;
            }
            for (int j = 0; j < n; j++) wa3[j] = diag[j] * x[j];
            xnorm = ora_enorm(n, wa3);
            delta = factor * xnorm;
            if (delta == 0.0) delta = factor;
        }
        for (int i = 0; i < m; i++) wa4[i] = fvec[i];
        for (int j = 0; j < n; j++) {
            double fjj = fjac[j + (size_t)j * m];
            if (fjj != 0.0) {
                double sum = 0.0;
                for (int i = j; i < m; i++)
                    sum += fjac[i + (size_t)j * m] * wa4[i];
                double temp = -sum / fjj;
                for (int i = j; i < m; i++)
                    wa4[i] += fjac[i + (size_t)j * m] * temp;
            }
            fjac[j + (size_t)j * m] = wa1[j];
            qtf[j] = wa4[j];
        }
        gnorm = 0.0;
        if (fnorm != 0.0) {
            for (int j = 0; j < n; j++) {
                int l = ipvt[j];
                if (wa2[l] != 0.0) {
                    double sum = 0.0;
                    for (int i = 0; i <= j; i++)
                        sum += fjac[i + (size_t)j * m] * (qtf[i] / fnorm);
                    gnorm = dmax(gnorm, fabs(sum / wa2[l]));
                }
            }
        }
        if (gnorm <= gtol) { info = 4; break; }
        for (int j = 0; j < n; j++) diag[j] = dmax(diag[j], wa2[j]);

        int done = 0;
        for (;;) {
            lmpar(n, fjac, m, ipvt, diag, qtf, delta, &par, wa1, wa2, wa3,
                  wa4);
            for (int j = 0; j < n; j++) {
                wa1[j] = -wa1[j];
                wa2[j] = x[j] + wa1[j];
                wa3[j] = diag[j] * wa1[j];
            }
            double pnorm = ora_enorm(n, wa3);
            if (iter == 1) delta = dmin(delta, pnorm);
            if (fcn(ctx, m, n, wa2, wa4) < 0) { info = -1; done = 1; break; }
            *nfev += 1;
            double fnorm1 = ora_enorm(m, wa4);
            double actred = -1.0;
            if (0.1 * fnorm1 < fnorm) {
                double r = fnorm1 / fnorm;
                actred = 1.0 - r * r;
            }
            for (int j = 0; j < n; j++) {
                wa3[j] = 0.0;
                int l = ipvt[j];
                double temp = wa1[l];
                for (int i = 0; i <= j; i++)
                    wa3[i] += fjac[i + (size_t)j * m] * temp;
            }
            double temp1 = ora_enorm(n, wa3) / fnorm;
            double temp2 = (sqrt(par) * pnorm) / fnorm;
            double prered = temp1 * temp1 + temp2 * temp2 / 0.5;
            double dirder = -(temp1 * temp1 + temp2 * temp2);
            double ratio = 0.0;
            if (prered != 0.0) ratio = actred / prered;
            if (ratio <= 0.25) {
                double temp = 0.5;
                if (actred >= 0.0) temp = 0.5;
                if (actred < 0.0) temp = 0.5 * dirder / (dirder + 0.5 * actred);
                if (0.1 * fnorm1 >= fnorm || temp < 0.1) temp = 0.1;
                delta = temp * dmin(delta, pnorm / 0.1);
                par /= temp;
            } else if (par == 0.0 || ratio >= 0.75) {
                delta = pnorm / 0.5;
                par *= 0.5;
            }
            if (ratio >= 1.0e-4) {
                for (int j = 0; j < n; j++) {
                    x[j] = wa2[j];
                    wa2[j] = diag[j] * x[j];
                }
                for (int i = 0; i < m; i++) fvec[i] = wa4[i];
                xnorm = ora_enorm(n, wa2);
                fnorm = fnorm1;
                iter++;
            }
            if (fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0)
                info = 1;
            if (delta <= xtol * xnorm) info = 2;
            if (fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0
                && info == 2)
                info = 3;
            if (info != 0) { done = 1; break; }
            if (*nfev >= maxfev) info = 5;
            if (fabs(actred) <= EPSMCH && prered <= EPSMCH
                && 0.5 * ratio <= 1.0)
                info = 6;
            if (delta <= EPSMCH * xnorm) info = 7;
            if (gnorm <= EPSMCH) info = 8;
            if (info != 0) { done = 1; break; }
            if (ratio >= 1.0e-4) break;
        }
        if (done) break;
    }
    free(diag);
    return info;
}

int ora_lmder(ora_fcn fcn, ora_jacfcn jfcn, void *ctx, int m, int n,
              double *x, double *fvec, double *fjac, double ftol, double xtol,
              double gtol, int maxfev, double factor, int *nfev, int *njev,
              int *ipvt, double *qtf) {
    return lm_driver(fcn, jfcn, ctx, m, n, x, fvec, fjac, ftol, xtol, gtol,
                     maxfev, 0.0, factor, nfev, njev, ipvt, qtf);
}

int ora_lmdif(ora_fcn fcn, void *ctx, int m, int n, double *x, double *fvec,
              double *fjac, double ftol, double xtol, double gtol, int maxfev,
              double epsfcn, double factor, int *nfev, int *ipvt,
              double *qtf) {
    int njev;
    return lm_driver(fcn, NULL, ctx, m, n, x, fvec, fjac, ftol, xtol, gtol,
                     maxfev, epsfcn, factor, nfev, &njev, ipvt, qtf);
}
