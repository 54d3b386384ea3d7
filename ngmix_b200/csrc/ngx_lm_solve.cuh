// The per-object Levenberg-Marquardt iteration of the engine, written as a
// state machine that is advanced by ONE thread of the CTA between pixel
// passes.  It reproduces MINPACK lmder/lmdif as driven by the reference
// (ngmix/fitting/leastsqbound.py:445-447: ftol/xtol from DEFAULT_LM_PARS,
// gtol = 0, factor = 100, mode 1 scaling), but it never sees the m x n
// Jacobian: the pixel passes reduce it to the sufficient statistics
//     A = J^T J (n x n),  b = J^T f (n),  |f|^2
// and a diagonally pivoted Cholesky factor of A stands in for qrfac's R
// (same pivot rule: largest remaining column norm; R differs from MINPACK's
// only by row signs, which lmpar/qrsolv are invariant to).
//
// The file compiles as plain C++ as well (NGX_HD empty), which is how
// tests/test_lm_solve_host.py drives it against the oracle on the CPU.
#pragma once
#include <math.h>
#include <float.h>

#ifdef __CUDACC__
#define NGX_HD __host__ __device__
// the solver runs on one thread per CTA: keep ONE compact, non-unrolled copy
// of every routine so it does not evict the pixel loops from the 32 KB
// instruction cache (ncu: 50% of its stall samples were no_instruction)
#define NGX_NI __noinline__
#else
#define NGX_HD
#define NGX_NI
#endif

namespace ngx {

// workspace size: the galaxy models and multi-band fits with up to 3 bands
// need 8; ngx_lm_wide.cu compiles the same code with 14 (5 coelliptical
// gaussians = 4 + 2 * 5 parameters, multi-band fits with up to 5 + 7)
#ifndef LM_MAXP
#define LM_MAXP 8
#endif
#define LM_EPSMCH 2.220446049250313e-16
#define LM_DWARF 2.2250738585072014e-308

// actions requested from the pixel passes
enum { LM_ACT_DONE = 0, LM_ACT_EVAL_F = 1, LM_ACT_EVAL_J = 2 };

struct LMState {
    int n;                     // number of parameters
    int action;                // what the CTA must evaluate next
    int info;                  // MINPACK info / scipy ier
    int nfev, njev, iter, maxfev;
    int inner;                 // 1 while inside the inner (lmpar) loop
    double ftol, xtol, gtol, factor;
    double x[LM_MAXP];         // current accepted iterate
    double xtrial[LM_MAXP];    // point the next EVAL_F is wanted at
    double diag[LM_MAXP], qtf[LM_MAXP], acnorm[LM_MAXP];
    double step[LM_MAXP];      // wa1 = -p
    double R[LM_MAXP * LM_MAXP];   // column major, upper triangle = R
    int ipvt[LM_MAXP];
    double fnorm, xnorm, delta, par, gnorm, pnorm;
    // predicted reduction and directional derivative of the proposed step
    // (lmder's prered / dirder): they depend only on the factorisation and
    // the step, so lm_propose computes them and the caller can tell a likely
    // final evaluation (prered <= ftol) before running it
    double prered, dirder;
    // workspace.  The state lives in shared memory on the device; keeping the
    // temporaries here (not in per-thread local arrays) matters: local memory
    // is interleaved by lane, so one thread walking a local array touches a
    // new 128-byte line per element (ncu: 42% of the solver's stalls were
    // long-scoreboard, 1.1 GB of DRAM traffic per launch).
    double A[LM_MAXP * LM_MAXP], b[LM_MAXP];     // J^T J, J^T f
    double T1[LM_MAXP * LM_MAXP], T2[LM_MAXP * LM_MAXP];
    double v[10][LM_MAXP];
};

NGX_HD inline double lm_max(double a, double b) { return a >= b ? a : b; }
NGX_HD inline double lm_min(double a, double b) { return a <= b ? a : b; }

// euclidean norm of a short vector.  MINPACK's enorm rescales to dodge
// overflow/underflow; for n <= 8 parameter-space vectors of sane magnitude a
// plain sum of squares rounds within an ulp or two of it.
NGX_HD inline double lm_enorm(int n, const double *x) {
    double s = 0.0;
#pragma unroll 1
    for (int i = 0; i < n; i++) s += x[i] * x[i];
    return sqrt(s);
}

// Pivoted Cholesky A P = P R^T R  (A symmetric n x n, column major, ld
// LM_MAXP).  Pivot rule of qrfac: at step j take the first column k >= j with
// the largest remaining norm.  A non-positive pivot leaves a zero row (rank
// deficiency), as qrfac leaves a zero diagonal.
NGX_HD NGX_NI static void lm_pchol(int n, const double *A, double *R, int *ipvt,
                            double *acnorm, double *d) {
#pragma unroll 1
    for (int j = 0; j < n; j++) {
        ipvt[j] = j;
        d[j] = A[j + LM_MAXP * j];
        acnorm[j] = sqrt(d[j] > 0.0 ? d[j] : 0.0);
#pragma unroll 1
        for (int i = 0; i < n; i++) R[i + LM_MAXP * j] = 0.0;
    }
#pragma unroll 1
    for (int j = 0; j < n; j++) {
        int kmax = j;
#pragma unroll 1
        for (int k = j; k < n; k++)
            if (d[k] > d[kmax]) kmax = k;
        if (kmax != j) {
            // swap columns j and kmax of R (rows < j), d and ipvt
#pragma unroll 1
            for (int i = 0; i < j; i++) {
                double t = R[i + LM_MAXP * j];
                R[i + LM_MAXP * j] = R[i + LM_MAXP * kmax];
                R[i + LM_MAXP * kmax] = t;
            }
            double t = d[j]; d[j] = d[kmax]; d[kmax] = t;
            int ti = ipvt[j]; ipvt[j] = ipvt[kmax]; ipvt[kmax] = ti;
        }
        double piv = d[j];
        if (!(piv > 0.0)) {
            // remaining columns are numerically dependent
#pragma unroll 1
            for (int k = j; k < n; k++) R[j + LM_MAXP * k] = 0.0;
#pragma unroll 1
            for (int jj = j; jj < n; jj++)
#pragma unroll 1
                for (int k = jj; k < n; k++) R[jj + LM_MAXP * k] = 0.0;
            return;
        }
        double rjj = sqrt(piv);
        R[j + LM_MAXP * j] = rjj;
        // one reciprocal per row instead of a division per element: a double
        // division is a ~30-instruction dependent sequence on this thread
        double irjj = 1.0 / rjj;
        int pj = ipvt[j];
#pragma unroll 1
        for (int k = j + 1; k < n; k++) {
            int pk = ipvt[k];
            double s = A[pj + LM_MAXP * pk];
#pragma unroll 1
            for (int i = 0; i < j; i++)
                s -= R[i + LM_MAXP * j] * R[i + LM_MAXP * k];
            double rjk = s * irjj;
            R[j + LM_MAXP * k] = rjk;
            d[k] -= rjk * rjk;
        }
    }
}

// MINPACK qrsolv on an n x n R (column major, ld LM_MAXP)
NGX_HD NGX_NI static void lm_qrsolv(int n, double *r, const int *ipvt,
                             const double *diag, const double *qtb, double *x,
                             double *sdiag, double *wa) {
    const int ld = LM_MAXP;
#pragma unroll 1
    for (int j = 0; j < n; j++) {
#pragma unroll 1
        for (int i = j; i < n; i++) r[i + ld * j] = r[j + ld * i];
        x[j] = r[j + ld * j];
        wa[j] = qtb[j];
    }
#pragma unroll 1
    for (int j = 0; j < n; j++) {
        int l = ipvt[j];
        if (diag[l] != 0.0) {
#pragma unroll 1
            for (int k = j; k < n; k++) sdiag[k] = 0.0;
            sdiag[j] = diag[l];
            double qtbpj = 0.0;
#pragma unroll 1
            for (int k = j; k < n; k++) {
                if (sdiag[k] == 0.0) continue;
                double c, s;
                double rkk = r[k + ld * k];
                if (fabs(rkk) < fabs(sdiag[k])) {
                    double cotan = rkk / sdiag[k];
                    s = 0.5 / sqrt(0.25 + 0.25 * (cotan * cotan));
                    c = s * cotan;
                } else {
                    double tn = sdiag[k] / rkk;
                    c = 0.5 / sqrt(0.25 + 0.25 * (tn * tn));
                    s = c * tn;
                }
                r[k + ld * k] = c * rkk + s * sdiag[k];
                double temp = c * wa[k] + s * qtbpj;
                qtbpj = -s * wa[k] + c * qtbpj;
                wa[k] = temp;
#pragma unroll 1
                for (int i = k + 1; i < n; i++) {
                    temp = c * r[i + ld * k] + s * sdiag[i];
                    sdiag[i] = -s * r[i + ld * k] + c * sdiag[i];
                    r[i + ld * k] = temp;
                }
            }
        }
        sdiag[j] = r[j + ld * j];
        r[j + ld * j] = x[j];
    }
    int nsing = n;
#pragma unroll 1
    for (int j = 0; j < n; j++) {
        if (sdiag[j] == 0.0 && nsing == n) nsing = j;
        if (nsing < n) wa[j] = 0.0;
    }
#pragma unroll 1
    for (int k = 1; k <= nsing; k++) {
        int j = nsing - k;
        double sum = 0.0;
#pragma unroll 1
        for (int i = j + 1; i < nsing; i++) sum += r[i + ld * j] * wa[i];
        wa[j] = (wa[j] - sum) / sdiag[j];
    }
#pragma unroll 1
    for (int j = 0; j < n; j++) x[ipvt[j]] = wa[j];
}

// MINPACK lmpar
NGX_HD NGX_NI static void lm_lmpar(int n, double *r, const int *ipvt,
                            const double *diag, const double *qtb,
                            double delta, double *par, double *x,
                            double *sdiag, double *wa1, double *wa2) {
    const int ld = LM_MAXP;
    int nsing = n;
#pragma unroll 1
    for (int j = 0; j < n; j++) {
        wa1[j] = qtb[j];
        if (r[j + ld * j] == 0.0 && nsing == n) nsing = j;
        if (nsing < n) wa1[j] = 0.0;
    }
#pragma unroll 1
    for (int k = 1; k <= nsing; k++) {
        int j = nsing - k;
        wa1[j] /= r[j + ld * j];
        double temp = wa1[j];
#pragma unroll 1
        for (int i = 0; i < j; i++) wa1[i] -= r[i + ld * j] * temp;
    }
#pragma unroll 1
    for (int j = 0; j < n; j++) x[ipvt[j]] = wa1[j];

    int iter = 0;
#pragma unroll 1
    for (int j = 0; j < n; j++) wa2[j] = diag[j] * x[j];
    double dxnorm = lm_enorm(n, wa2);
    double fp = dxnorm - delta;
    if (fp <= 0.1 * delta) {
        *par = 0.0;
        return;
    }
    double parl = 0.0;
    if (nsing >= n) {
#pragma unroll 1
        for (int j = 0; j < n; j++) {
            int l = ipvt[j];
            wa1[j] = diag[l] * (wa2[l] / dxnorm);
        }
#pragma unroll 1
        for (int j = 0; j < n; j++) {
            double sum = 0.0;
#pragma unroll 1
            for (int i = 0; i < j; i++) sum += r[i + ld * j] * wa1[i];
            wa1[j] = (wa1[j] - sum) / r[j + ld * j];
        }
        double temp = lm_enorm(n, wa1);
        parl = ((fp / delta) / temp) / temp;
    }
#pragma unroll 1
    for (int j = 0; j < n; j++) {
        double sum = 0.0;
#pragma unroll 1
        for (int i = 0; i <= j; i++) sum += r[i + ld * j] * qtb[i];
        wa1[j] = sum / diag[ipvt[j]];
    }
    double gnorm = lm_enorm(n, wa1);
    double paru = gnorm / delta;
    if (paru == 0.0) paru = LM_DWARF / lm_min(delta, 0.1);
    *par = lm_max(*par, parl);
    *par = lm_min(*par, paru);
    if (*par == 0.0) *par = gnorm / dxnorm;
#pragma unroll 1
    for (;;) {
        iter++;
        if (*par == 0.0) *par = lm_max(LM_DWARF, 0.001 * paru);
        double temp = sqrt(*par);
#pragma unroll 1
        for (int j = 0; j < n; j++) wa1[j] = temp * diag[j];
        lm_qrsolv(n, r, ipvt, wa1, qtb, x, sdiag, wa2);
#pragma unroll 1
        for (int j = 0; j < n; j++) wa2[j] = diag[j] * x[j];
        dxnorm = lm_enorm(n, wa2);
        temp = fp;
        fp = dxnorm - delta;
        if (fabs(fp) <= 0.1 * delta
            || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10)
            break;
#pragma unroll 1
        for (int j = 0; j < n; j++) {
            int l = ipvt[j];
            wa1[j] = diag[l] * (wa2[l] / dxnorm);
        }
#pragma unroll 1
        for (int j = 0; j < n; j++) {
            wa1[j] /= sdiag[j];
            temp = wa1[j];
#pragma unroll 1
            for (int i = j + 1; i < n; i++) wa1[i] -= r[i + ld * j] * temp;
        }
        temp = lm_enorm(n, wa1);
        double parc = ((fp / delta) / temp) / temp;
        if (fp > 0.0) parl = lm_max(parl, *par);
        if (fp < 0.0) paru = lm_min(paru, *par);
        *par = lm_max(parl, *par + parc);
    }
}

// ---------------------------------------------------------------------
NGX_HD NGX_NI static void lm_init(LMState *s, int n, const double *x0, int maxfev,
                           double ftol, double xtol) {
    s->n = n;
    s->maxfev = maxfev;
    s->ftol = ftol;
    s->xtol = xtol;
    s->gtol = 0.0;
    s->factor = 100.0;
    s->info = 0;
    s->nfev = 0;
    s->njev = 0;
    s->iter = 1;
    s->inner = 0;
    s->par = 0.0;
    s->delta = 0.0;
    s->xnorm = 0.0;
    s->gnorm = 0.0;
    s->pnorm = 0.0;
    s->fnorm = 0.0;
    s->prered = 1.0;
    s->dirder = 0.0;
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        s->x[i] = x0[i];
        s->xtrial[i] = x0[i];
        s->diag[i] = 0.0;
    }
    s->action = LM_ACT_EVAL_F;
}

// compute the next trial point from the current factorisation (lmpar)
NGX_HD NGX_NI static void lm_propose(LMState *s) {
    int n = s->n;
    double *p = s->v[4], *wa3 = s->v[5];
    lm_lmpar(n, s->R, s->ipvt, s->diag, s->qtf, s->delta, &s->par, p,
             s->v[1], s->v[2], s->v[3]);
#pragma unroll 1
    for (int j = 0; j < n; j++) {
        s->step[j] = -p[j];
        s->xtrial[j] = s->x[j] + s->step[j];
        wa3[j] = s->diag[j] * s->step[j];
    }
    s->pnorm = lm_enorm(n, wa3);
    if (s->iter == 1) s->delta = lm_min(s->delta, s->pnorm);
    // |R P^T p|
#pragma unroll 1
    for (int j = 0; j < n; j++) wa3[j] = 0.0;
#pragma unroll 1
    for (int j = 0; j < n; j++) {
        double temp = s->step[s->ipvt[j]];
#pragma unroll 1
        for (int i = 0; i <= j; i++) wa3[i] += s->R[i + LM_MAXP * j] * temp;
    }
    double temp1 = lm_enorm(n, wa3) / s->fnorm;
    double temp2 = (sqrt(s->par) * s->pnorm) / s->fnorm;
    s->prered = temp1 * temp1 + temp2 * temp2 / 0.5;
    s->dirder = -(temp1 * temp1 + temp2 * temp2);
    s->action = LM_ACT_EVAL_F;
}

// Feed the result of an EVAL_F pass: fnorm_new = |f(xtrial)| (NaN when the
// model or prior was out of range: the reference's -inf residual vector has
// a NaN MINPACK norm).  Returns 1 when the trial point was accepted (or was
// the initial point), i.e. when the caller must adopt the pass's statistics.
NGX_HD NGX_NI static int lm_feed_f(LMState *s, double fnorm_new) {
    int n = s->n;
    s->nfev += 1;
    if (s->nfev == 1) {
        // initial evaluation
        s->fnorm = fnorm_new;
        s->action = LM_ACT_EVAL_J;
        return 1;
    }
    double fnorm1 = fnorm_new;
    double actred = -1.0;
    if (0.1 * fnorm1 < s->fnorm) {
        double r = fnorm1 / s->fnorm;
        actred = 1.0 - r * r;
    }
    const double prered = s->prered, dirder = s->dirder;
    double ratio = 0.0;
    if (prered != 0.0) ratio = actred / prered;
    if (ratio <= 0.25) {
        double temp = 0.5;
        if (actred < 0.0) temp = 0.5 * dirder / (dirder + 0.5 * actred);
        if (0.1 * fnorm1 >= s->fnorm || temp < 0.1) temp = 0.1;
        s->delta = temp * lm_min(s->delta, s->pnorm / 0.1);
        s->par /= temp;
    } else if (s->par == 0.0 || ratio >= 0.75) {
        s->delta = s->pnorm / 0.5;
        s->par *= 0.5;
    }
    int accepted = 0;
    if (ratio >= 1.0e-4) {
        double *wa2 = s->v[6];
#pragma unroll 1
        for (int j = 0; j < n; j++) {
            s->x[j] = s->xtrial[j];
            wa2[j] = s->diag[j] * s->x[j];
        }
        s->xnorm = lm_enorm(n, wa2);
        s->fnorm = fnorm1;
        s->iter += 1;
        accepted = 1;
    }
    int info = 0;
    bool small = fabs(actred) <= s->ftol && prered <= s->ftol
        && 0.5 * ratio <= 1.0;
    if (small) info = 1;
    if (s->delta <= s->xtol * s->xnorm) info = 2;
    if (small && info == 2) info = 3;
    if (info == 0) {
        if (s->nfev >= s->maxfev) info = 5;
        if (fabs(actred) <= LM_EPSMCH && prered <= LM_EPSMCH
            && 0.5 * ratio <= 1.0)
            info = 6;
        if (s->delta <= LM_EPSMCH * s->xnorm) info = 7;
        if (s->gnorm <= LM_EPSMCH) info = 8;
    }
    if (info != 0) {
        s->info = info;
        s->action = LM_ACT_DONE;
        return accepted;
    }
    if (ratio >= 1.0e-4) {
        s->action = LM_ACT_EVAL_J;      // back to the outer loop
    } else {
        lm_propose(s);                  // stay in the inner loop
    }
    return accepted;
}

// Feed the result of an EVAL_J pass at s->x: A = J^T J (column major, ld
// LM_MAXP), b = J^T f.  extra_nfev is the number of residual evaluations the
// Jacobian cost (n for forward differences, 0 for the analytic one).
NGX_HD NGX_NI static void lm_feed_j(LMState *s, const double *A, const double *b,
                             int extra_nfev) {
    int n = s->n;
    s->njev += 1;
    s->nfev += extra_nfev;
    lm_pchol(n, A, s->R, s->ipvt, s->acnorm, s->v[0]);
    if (s->iter == 1) {
        double *wa3 = s->v[5];
#pragma unroll 1
        for (int j = 0; j < n; j++) {
            s->diag[j] = s->acnorm[j];
            if (s->acnorm[j] == 0.0) s->diag[j] = 1.0;
            wa3[j] = s->diag[j] * s->x[j];
        }
        s->xnorm = lm_enorm(n, wa3);
        s->delta = s->factor * s->xnorm;
        if (s->delta == 0.0) s->delta = s->factor;
    }
    // qtf = R^-T P^T b
#pragma unroll 1
    for (int j = 0; j < n; j++) {
        double rjj = s->R[j + LM_MAXP * j];
        double v = b[s->ipvt[j]];
#pragma unroll 1
        for (int i = 0; i < j; i++) v -= s->R[i + LM_MAXP * j] * s->qtf[i];
        s->qtf[j] = (rjj != 0.0) ? v / rjj : 0.0;
    }
    // gnorm = max_j |(J^T f)_j| / (acnorm_j |f|).  MINPACK forms (J^T f)_j as
    // sum_i R[i,j] qtf[i]/|f|; that is b[ipvt_j]/|f| exactly, and gnorm only
    // feeds the gnorm <= gtol (= 0) and gnorm <= eps tests
    double gnorm = 0.0;
    if (s->fnorm != 0.0) {
        double ifnorm = 1.0 / s->fnorm;
#pragma unroll 1
        for (int j = 0; j < n; j++) {
            int l = s->ipvt[j];
            if (s->acnorm[l] != 0.0)
                gnorm = lm_max(gnorm, fabs((b[l] * ifnorm) / s->acnorm[l]));
        }
    }
    s->gnorm = gnorm;
    if (gnorm <= s->gtol) {
        s->info = 4;
        s->action = LM_ACT_DONE;
        return;
    }
#pragma unroll 1
    for (int j = 0; j < n; j++) s->diag[j] = lm_max(s->diag[j], s->acnorm[j]);
    lm_propose(s);
}

// cov0 = (J^T J)^-1 = P R^-1 R^-T P^T from the last factorisation (row
// major n x n out).  Returns 0 when R is singular (scipy: cov_x None).
NGX_HD NGX_NI static int lm_cov0(LMState *s, double *cov) {
    int n = s->n;
    const int ld = LM_MAXP;
    double *rinv = s->T1;
#pragma unroll 1
    for (int j = 0; j < n; j++)
        if (s->R[j + ld * j] == 0.0) return 0;
#pragma unroll 1
    for (int i = 0; i < n * LM_MAXP; i++) rinv[i] = 0.0;
#pragma unroll 1
    for (int j = 0; j < n; j++) {
        rinv[j + ld * j] = 1.0 / s->R[j + ld * j];
#pragma unroll 1
        for (int i = j - 1; i >= 0; i--) {
            double sum = 0.0;
#pragma unroll 1
            for (int k = i + 1; k <= j; k++)
                sum += s->R[i + ld * k] * rinv[k + ld * j];
            rinv[i + ld * j] = -sum / s->R[i + ld * i];
        }
    }
    // row ipvt[j] of P R^-1 is row j of R^-1
#pragma unroll 1
    for (int a = 0; a < n; a++)
#pragma unroll 1
        for (int b = 0; b <= a; b++) {
            double sum = 0.0;
#pragma unroll 1
            for (int k = 0; k < n; k++)
                sum += rinv[a + ld * k] * rinv[b + ld * k];
            cov[s->ipvt[a] * n + s->ipvt[b]] = sum;
            cov[s->ipvt[b] * n + s->ipvt[a]] = sum;
        }
    return 1;
}

// _test_cov (leastsqbound.py:165-180) for a matrix that is s_sq * B B^T by
// construction: a negative eigenvalue exists iff the matrix is not positive
// semi-definite, decided by a Cholesky sweep with a rounding-level floor.
// Returns flag bits (NEG_COV_EIG 1024, NEG_COV_DIAG 2048, EIG_NOTFINITE 8192)
NGX_HD NGX_NI static int lm_test_cov(int n, const double *cov, double *L) {
    int flags = 0;
#pragma unroll 1
    for (int i = 0; i < n * n; i++)
        if (!(fabs(cov[i]) <= DBL_MAX)) return 8192;
#pragma unroll 1
    for (int i = 0; i < n; i++)
        if (cov[i * n + i] < 0.0) flags |= 2048;
    double scale = 0.0;
#pragma unroll 1
    for (int i = 0; i < n; i++) scale = lm_max(scale, fabs(cov[i * n + i]));
    double tol = 64.0 * LM_EPSMCH * scale;
    int neg = 0;
#pragma unroll 1
    for (int j = 0; j < n && !neg; j++) {
        double dsum = cov[j * n + j];
#pragma unroll 1
        for (int k = 0; k < j; k++) dsum -= L[j * n + k] * L[j * n + k];
        if (dsum < -tol) { neg = 1; break; }
        double ljj = sqrt(dsum > 0.0 ? dsum : 0.0);
        L[j * n + j] = ljj;
#pragma unroll 1
        for (int i = j + 1; i < n; i++) {
            double sum = cov[i * n + j];
#pragma unroll 1
            for (int k = 0; k < j; k++) sum -= L[i * n + k] * L[j * n + k];
            L[i * n + j] = ljj > 0.0 ? sum / ljj : 0.0;
        }
    }
    if (neg || (flags & 2048)) flags |= 1024;
    return flags;
}

}  // namespace ngx
