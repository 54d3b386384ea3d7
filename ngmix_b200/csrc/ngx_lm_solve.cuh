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

// ---------------------------------------------------------------------
// Warp-cooperative execution.  On the device the routines below (lm_pchol,
// lm_qrsolv, lm_lmpar, lm_propose, lm_feed_f, lm_feed_j) are entered by ALL
// 32 lanes of one converged warp; the state lives in shared memory.
//   LM_FOR(k, lo, hi)  one iteration per lane (n <= LM_MAXP < 32): loops whose
//                      iterations are independent (a column each)
//   LM_ONE             lane 0 only: scalar writes to the shared state
//   LM_SYNC            __syncwarp(): orders the shared-memory accesses of the
//                      lanes between a region's writes and the next reads
// Everything else is computed redundantly (and identically) by every lane.
// Every per-element operation sequence is the one of the scalar algorithm
// (each sum is formed by one lane in the original order), so the results are
// bit-identical to a one-thread run; on the host the macros degenerate to
// plain loops, which is how tests/test_lm_solve_host.py drives this file.
// A one-thread solver held each CTA for ~30% of a fit (ncu r1m: 1.94 barrier
// stalls per issue); the O(n^3) / O(n^2) loops now take O(n^2) / O(n) steps.
#ifdef __CUDA_ARCH__
#define LM_LANE ((int)(threadIdx.x & 31))
#define LM_FOR(k, lo, hi) \
    for (int k = (lo) + LM_LANE; k < (hi); k += 32)
#define LM_ONE if (LM_LANE == 0)
#define LM_SYNC __syncwarp()
#else
#define LM_FOR(k, lo, hi) for (int k = (lo); k < (hi); k++)
#define LM_ONE
#define LM_SYNC
#endif

// Pivoted Cholesky A P = P R^T R  (A symmetric n x n, column major, ld
// LM_MAXP).  Pivot rule of qrfac: at step j take the first column k >= j with
// the largest remaining norm.  A non-positive pivot leaves a zero row (rank
// deficiency), as qrfac leaves a zero diagonal.  Lane k owns column k.
NGX_HD NGX_NI static void lm_pchol(int n, const double *A, double *R, int *ipvt,
                            double *acnorm, double *d) {
    LM_FOR(j, 0, n) {
        ipvt[j] = j;
        double dj = A[j + LM_MAXP * j];
        d[j] = dj;
        acnorm[j] = sqrt(dj > 0.0 ? dj : 0.0);
#pragma unroll 1
        for (int i = 0; i < n; i++) R[i + LM_MAXP * j] = 0.0;
    }
    LM_SYNC;
#pragma unroll 1
    for (int j = 0; j < n; j++) {
        int kmax = j;
        double dmax = d[j];
#pragma unroll 1
        for (int k = j + 1; k < n; k++) {
            double dk = d[k];
            if (dk > dmax) { kmax = k; dmax = dk; }
        }
        LM_SYNC;
        if (kmax != j) {
            // swap columns j and kmax of R (rows < j), d and ipvt
            LM_FOR(i, 0, j) {
                double t = R[i + LM_MAXP * j];
                R[i + LM_MAXP * j] = R[i + LM_MAXP * kmax];
                R[i + LM_MAXP * kmax] = t;
            }
            LM_ONE {
                double t = d[j]; d[j] = d[kmax]; d[kmax] = t;
                int ti = ipvt[j]; ipvt[j] = ipvt[kmax]; ipvt[kmax] = ti;
            }
            LM_SYNC;
        }
        double piv = dmax;
        if (!(piv > 0.0)) {
            // remaining columns are numerically dependent
            LM_FOR(k, j, n) {
#pragma unroll 1
                for (int jj = j; jj <= k; jj++) R[jj + LM_MAXP * k] = 0.0;
            }
            LM_SYNC;
            return;
        }
        double rjj = sqrt(piv);
        // one reciprocal per row instead of a division per element: a double
        // division is a ~30-instruction dependent sequence
        double irjj = 1.0 / rjj;
        int pj = ipvt[j];
        LM_ONE R[j + LM_MAXP * j] = rjj;
        LM_FOR(k, j + 1, n) {
            int pk = ipvt[k];
            double s = A[pj + LM_MAXP * pk];
#pragma unroll 1
            for (int i = 0; i < j; i++)
                s -= R[i + LM_MAXP * j] * R[i + LM_MAXP * k];
            double rjk = s * irjj;
            R[j + LM_MAXP * k] = rjk;
            d[k] -= rjk * rjk;
        }
        LM_SYNC;
    }
}

// MINPACK qrsolv on an n x n R (column major, ld LM_MAXP)
NGX_HD NGX_NI static void lm_qrsolv(int n, double *r, const int *ipvt,
                             const double *diag, const double *qtb, double *x,
                             double *sdiag, double *wa) {
    const int ld = LM_MAXP;
    LM_FOR(j, 0, n) {
        // lower triangle <- upper triangle (disjoint), lane j its column
#pragma unroll 1
        for (int i = j; i < n; i++) r[i + ld * j] = r[j + ld * i];
        x[j] = r[j + ld * j];
        wa[j] = qtb[j];
    }
    LM_SYNC;
#pragma unroll 1
    for (int j = 0; j < n; j++) {
        int l = ipvt[j];
        const double dl = diag[l];
        if (dl != 0.0) {
            LM_FOR(k, j, n) sdiag[k] = (k == j) ? dl : 0.0;
            LM_SYNC;
            double qtbpj = 0.0;
#pragma unroll 1
            for (int k = j; k < n; k++) {
                const double sk = sdiag[k];
                if (sk == 0.0) continue;
                double c, s;
                const double rkk = r[k + ld * k];
                if (fabs(rkk) < fabs(sk)) {
                    double cotan = rkk / sk;
                    s = 0.5 / sqrt(0.25 + 0.25 * (cotan * cotan));
                    c = s * cotan;
                } else {
                    double tn = sk / rkk;
                    c = 0.5 / sqrt(0.25 + 0.25 * (tn * tn));
                    s = c * tn;
                }
                const double wak = wa[k];
                LM_SYNC;
                LM_ONE {
                    r[k + ld * k] = c * rkk + s * sk;
                    wa[k] = c * wak + s * qtbpj;
                }
                qtbpj = -s * wak + c * qtbpj;
                LM_FOR(i, k + 1, n) {
                    const double rik = r[i + ld * k], sdi = sdiag[i];
                    r[i + ld * k] = c * rik + s * sdi;
                    sdiag[i] = -s * rik + c * sdi;
                }
                LM_SYNC;
            }
        }
        const double rjj = r[j + ld * j], xj = x[j];
        LM_SYNC;
        LM_ONE {
            sdiag[j] = rjj;
            r[j + ld * j] = xj;
        }
        LM_SYNC;
    }
    int nsing = n;
#pragma unroll 1
    for (int j = 0; j < n; j++)
        if (sdiag[j] == 0.0 && nsing == n) nsing = j;
    LM_FOR(j, nsing, n) wa[j] = 0.0;
    LM_SYNC;
    // back substitution: each sum runs over ascending i as in MINPACK, so
    // the steps stay sequential (the divisions dominate them anyway)
#pragma unroll 1
    for (int k = 1; k <= nsing; k++) {
        int j = nsing - k;
        double sum = 0.0;
#pragma unroll 1
        for (int i = j + 1; i < nsing; i++) sum += r[i + ld * j] * wa[i];
        const double wj = (wa[j] - sum) / sdiag[j];
        LM_SYNC;
        LM_ONE wa[j] = wj;
        LM_SYNC;
    }
    LM_FOR(j, 0, n) x[ipvt[j]] = wa[j];
    LM_SYNC;
}

// MINPACK lmpar
NGX_HD NGX_NI static void lm_lmpar(int n, double *r, const int *ipvt,
                            const double *diag, const double *qtb,
                            double delta, double *par_io, double *x,
                            double *sdiag, double *wa1, double *wa2) {
    const int ld = LM_MAXP;
    double par = *par_io;
    int nsing = n;
#pragma unroll 1
    for (int j = 0; j < n; j++)
        if (r[j + ld * j] == 0.0 && nsing == n) nsing = j;
    LM_FOR(j, 0, n) wa1[j] = j < nsing ? qtb[j] : 0.0;
    LM_SYNC;
#pragma unroll 1
    for (int k = 1; k <= nsing; k++) {
        int j = nsing - k;
        const double temp = wa1[j] / r[j + ld * j];
        LM_SYNC;
        LM_ONE wa1[j] = temp;
        LM_FOR(i, 0, j) wa1[i] -= r[i + ld * j] * temp;
        LM_SYNC;
    }
    LM_FOR(j, 0, n) x[ipvt[j]] = wa1[j];
    LM_SYNC;

    int iter = 0;
    LM_FOR(j, 0, n) wa2[j] = diag[j] * x[j];
    LM_SYNC;
    double dxnorm = lm_enorm(n, wa2);
    double fp = dxnorm - delta;
    if (fp <= 0.1 * delta) {
        LM_SYNC;
        LM_ONE *par_io = 0.0;
        LM_SYNC;
        return;
    }
    double parl = 0.0;
    if (nsing >= n) {
        LM_FOR(j, 0, n) {
            int l = ipvt[j];
            wa1[j] = diag[l] * (wa2[l] / dxnorm);
            sdiag[j] = 0.0;     // running sums of the forward substitution
        }
        LM_SYNC;
        // wa1[j] = (wa1[j] - sum_{i<j} r[i,j] wa1[i]) / r[j,j]: lane j adds
        // its sum's terms as the wa1[i] become final, in ascending i
#pragma unroll 1
        for (int j = 0; j < n; j++) {
            const double wj = (wa1[j] - sdiag[j]) / r[j + ld * j];
            LM_SYNC;
            LM_ONE wa1[j] = wj;
            LM_FOR(i, j + 1, n) sdiag[i] += r[j + ld * i] * wj;
            LM_SYNC;
        }
        double temp = lm_enorm(n, wa1);
        parl = ((fp / delta) / temp) / temp;
        LM_SYNC;
    }
    LM_FOR(j, 0, n) {
        double sum = 0.0;
#pragma unroll 1
        for (int i = 0; i <= j; i++) sum += r[i + ld * j] * qtb[i];
        wa1[j] = sum / diag[ipvt[j]];
    }
    LM_SYNC;
    double gnorm = lm_enorm(n, wa1);
    double paru = gnorm / delta;
    if (paru == 0.0) paru = LM_DWARF / lm_min(delta, 0.1);
    par = lm_max(par, parl);
    par = lm_min(par, paru);
    if (par == 0.0) par = gnorm / dxnorm;
#pragma unroll 1
    for (;;) {
        iter++;
        if (par == 0.0) par = lm_max(LM_DWARF, 0.001 * paru);
        double temp = sqrt(par);
        LM_SYNC;
        LM_FOR(j, 0, n) wa1[j] = temp * diag[j];
        LM_SYNC;
        lm_qrsolv(n, r, ipvt, wa1, qtb, x, sdiag, wa2);
        LM_FOR(j, 0, n) wa2[j] = diag[j] * x[j];
        LM_SYNC;
        dxnorm = lm_enorm(n, wa2);
        temp = fp;
        fp = dxnorm - delta;
        if (fabs(fp) <= 0.1 * delta
            || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10)
            break;
        LM_FOR(j, 0, n) {
            int l = ipvt[j];
            wa1[j] = diag[l] * (wa2[l] / dxnorm);
        }
        LM_SYNC;
#pragma unroll 1
        for (int j = 0; j < n; j++) {
            const double wj = wa1[j] / sdiag[j];
            LM_SYNC;
            LM_ONE wa1[j] = wj;
            LM_FOR(i, j + 1, n) wa1[i] -= r[i + ld * j] * wj;
            LM_SYNC;
        }
        temp = lm_enorm(n, wa1);
        double parc = ((fp / delta) / temp) / temp;
        if (fp > 0.0) parl = lm_max(parl, par);
        if (fp < 0.0) paru = lm_min(paru, par);
        par = lm_max(parl, par + parc);
    }
    LM_SYNC;
    LM_ONE *par_io = par;
    LM_SYNC;
}

// ---------------------------------------------------------------------
// (one thread)
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
// (whole warp)
NGX_HD NGX_NI static void lm_propose(LMState *s) {
    int n = s->n;
    double *p = s->v[4], *wa3 = s->v[5];
    lm_lmpar(n, s->R, s->ipvt, s->diag, s->qtf, s->delta, &s->par, p,
             s->v[1], s->v[2], s->v[3]);
    LM_FOR(j, 0, n) {
        const double st = -p[j];
        s->step[j] = st;
        s->xtrial[j] = s->x[j] + st;
        wa3[j] = s->diag[j] * st;
    }
    LM_SYNC;
    const double pnorm = lm_enorm(n, wa3);
    const double par = s->par, fnorm = s->fnorm;
    const int first = s->iter == 1;
    const double delta0 = s->delta;
    LM_SYNC;
    // |R P^T p|: wa3[i] += R[i,j] step[ipvt[j]] over ascending j, lane i
    LM_FOR(j, 0, n) wa3[j] = 0.0;
    LM_SYNC;
#pragma unroll 1
    for (int j = 0; j < n; j++) {
        const double temp = s->step[s->ipvt[j]];
        LM_FOR(i, 0, j + 1) wa3[i] += s->R[i + LM_MAXP * j] * temp;
    }
    LM_SYNC;
    double temp1 = lm_enorm(n, wa3) / fnorm;
    double temp2 = (sqrt(par) * pnorm) / fnorm;
    LM_ONE {
        s->pnorm = pnorm;
        if (first) s->delta = lm_min(delta0, pnorm);
        s->prered = temp1 * temp1 + temp2 * temp2 / 0.5;
        s->dirder = -(temp1 * temp1 + temp2 * temp2);
        s->action = LM_ACT_EVAL_F;
    }
    LM_SYNC;
}

// Feed the result of an EVAL_F pass: fnorm_new = |f(xtrial)| (NaN when the
// model or prior was out of range: the reference's -inf residual vector has
// a NaN MINPACK norm).  Returns 1 when the trial point was accepted (or was
// the initial point), i.e. when the caller must adopt the pass's statistics.
// (whole warp; every lane returns the same value)
NGX_HD NGX_NI static int lm_feed_f(LMState *s, double fnorm_new) {
    int n = s->n;
    const int nfev = s->nfev + 1;
    if (nfev == 1) {
        // initial evaluation
        LM_SYNC;
        LM_ONE {
            s->nfev = nfev;
            s->fnorm = fnorm_new;
            s->action = LM_ACT_EVAL_J;
        }
        LM_SYNC;
        return 1;
    }
    double fnorm1 = fnorm_new;
    const double fnorm0 = s->fnorm;
    double delta = s->delta, par = s->par, xnorm = s->xnorm;
    const double pnorm = s->pnorm, gnorm = s->gnorm;
    const double ftol = s->ftol, xtol = s->xtol;
    const int maxfev = s->maxfev;
    double actred = -1.0;
    if (0.1 * fnorm1 < fnorm0) {
        double r = fnorm1 / fnorm0;
        actred = 1.0 - r * r;
    }
    const double prered = s->prered, dirder = s->dirder;
    double ratio = 0.0;
    if (prered != 0.0) ratio = actred / prered;
    if (ratio <= 0.25) {
        double temp = 0.5;
        if (actred < 0.0) temp = 0.5 * dirder / (dirder + 0.5 * actred);
        if (0.1 * fnorm1 >= fnorm0 || temp < 0.1) temp = 0.1;
        delta = temp * lm_min(delta, pnorm / 0.1);
        par /= temp;
    } else if (par == 0.0 || ratio >= 0.75) {
        delta = pnorm / 0.5;
        par *= 0.5;
    }
    int accepted = 0;
    if (ratio >= 1.0e-4) {
        double *wa2 = s->v[6];
        LM_FOR(j, 0, n) {
            const double xj = s->xtrial[j];
            s->x[j] = xj;
            wa2[j] = s->diag[j] * xj;
        }
        LM_SYNC;
        xnorm = lm_enorm(n, wa2);
        accepted = 1;
    }
    int info = 0;
    bool small = fabs(actred) <= ftol && prered <= ftol
        && 0.5 * ratio <= 1.0;
    if (small) info = 1;
    if (delta <= xtol * xnorm) info = 2;
    if (small && info == 2) info = 3;
    if (info == 0) {
        if (nfev >= maxfev) info = 5;
        if (fabs(actred) <= LM_EPSMCH && prered <= LM_EPSMCH
            && 0.5 * ratio <= 1.0)
            info = 6;
        if (delta <= LM_EPSMCH * xnorm) info = 7;
        if (gnorm <= LM_EPSMCH) info = 8;
    }
    LM_SYNC;
    LM_ONE {
        s->nfev = nfev;
        s->delta = delta;
        s->par = par;
        if (accepted) {
            s->xnorm = xnorm;
            s->fnorm = fnorm1;
            s->iter += 1;
        }
        if (info != 0) {
            s->info = info;
            s->action = LM_ACT_DONE;
        } else if (accepted) {
            s->action = LM_ACT_EVAL_J;      // back to the outer loop
        }
    }
    LM_SYNC;
    if (info == 0 && !accepted)
        lm_propose(s);                      // stay in the inner loop
    return accepted;
}

// Feed the result of an EVAL_J pass at s->x: A = J^T J (column major, ld
// LM_MAXP), b = J^T f.  extra_nfev is the number of residual evaluations the
// Jacobian cost (n for forward differences, 0 for the analytic one).
// (whole warp)
NGX_HD NGX_NI static void lm_feed_j(LMState *s, const double *A, const double *b,
                             int extra_nfev) {
    int n = s->n;
    const int first = s->iter == 1;
    const double fnorm = s->fnorm, gtol = s->gtol, factor = s->factor;
    LM_SYNC;
    LM_ONE {
        s->njev += 1;
        s->nfev += extra_nfev;
    }
    lm_pchol(n, A, s->R, s->ipvt, s->acnorm, s->v[0]);
    double *wa3 = s->v[5], *vv = s->v[6];
    if (first) {
        LM_FOR(j, 0, n) {
            double dj = s->acnorm[j];
            if (dj == 0.0) dj = 1.0;
            s->diag[j] = dj;
            wa3[j] = dj * s->x[j];
        }
        LM_SYNC;
        const double xnorm = lm_enorm(n, wa3);
        double delta = factor * xnorm;
        if (delta == 0.0) delta = factor;
        LM_ONE {
            s->xnorm = xnorm;
            s->delta = delta;
        }
        LM_SYNC;
    }
    // qtf = R^-T P^T b: v[k] loses R[j,k] qtf[j] over ascending j (the order
    // of the scalar dot-product form), lane k its element
    LM_FOR(k, 0, n) vv[k] = b[s->ipvt[k]];
    LM_SYNC;
#pragma unroll 1
    for (int j = 0; j < n; j++) {
        const double rjj = s->R[j + LM_MAXP * j];
        const double qj = (rjj != 0.0) ? vv[j] / rjj : 0.0;
        LM_ONE s->qtf[j] = qj;
        LM_FOR(k, j + 1, n) vv[k] -= s->R[j + LM_MAXP * k] * qj;
        LM_SYNC;
    }
    // gnorm = max_j |(J^T f)_j| / (acnorm_j |f|).  MINPACK forms (J^T f)_j as
    // sum_i R[i,j] qtf[i]/|f|; that is b[ipvt_j]/|f| exactly, and gnorm only
    // feeds the gnorm <= gtol (= 0) and gnorm <= eps tests
    double gnorm = 0.0;
    if (fnorm != 0.0) {
        double ifnorm = 1.0 / fnorm;
        LM_FOR(j, 0, n) {
            int l = s->ipvt[j];
            const double an = s->acnorm[l];
            wa3[j] = an != 0.0 ? fabs((b[l] * ifnorm) / an) : 0.0;
        }
        LM_SYNC;
#pragma unroll 1
        for (int j = 0; j < n; j++) gnorm = lm_max(gnorm, wa3[j]);
    }
    LM_SYNC;
    LM_ONE {
        s->gnorm = gnorm;
        if (gnorm <= gtol) {
            s->info = 4;
            s->action = LM_ACT_DONE;
        }
    }
    if (gnorm <= gtol) {
        LM_SYNC;
        return;
    }
    LM_FOR(j, 0, n) s->diag[j] = lm_max(s->diag[j], s->acnorm[j]);
    LM_SYNC;
    lm_propose(s);
}

// cov0 = (J^T J)^-1 = P R^-1 R^-T P^T from the last factorisation (row
// major n x n out).  Returns 0 when R is singular (scipy: cov_x None).
// (whole warp: lane j inverts column j of R, then one lane per matrix entry)
NGX_HD NGX_NI static int lm_cov0(LMState *s, double *cov) {
    int n = s->n;
    const int ld = LM_MAXP;
    double *rinv = s->T1;
#pragma unroll 1
    for (int j = 0; j < n; j++)
        if (s->R[j + ld * j] == 0.0) return 0;
    LM_FOR(j, 0, n) {
#pragma unroll 1
        for (int i = 0; i < n; i++) rinv[i + ld * j] = 0.0;
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
    LM_SYNC;
    // row ipvt[j] of P R^-1 is row j of R^-1
    LM_FOR(q, 0, n * (n + 1) / 2) {
        // q -> (a, b), b <= a
        int a = 0;
        while ((a + 1) * (a + 2) / 2 <= q) a++;
        int b = q - a * (a + 1) / 2;
        double sum = 0.0;
#pragma unroll 1
        for (int k = 0; k < n; k++)
            sum += rinv[a + ld * k] * rinv[b + ld * k];
        cov[s->ipvt[a] * n + s->ipvt[b]] = sum;
        cov[s->ipvt[b] * n + s->ipvt[a]] = sum;
    }
    LM_SYNC;
    return 1;
}

// _test_cov (leastsqbound.py:165-180) for a matrix that is s_sq * B B^T by
// construction: a negative eigenvalue exists iff the matrix is not positive
// semi-definite, decided by a Cholesky sweep with a rounding-level floor.
// Returns flag bits (NEG_COV_EIG 1024, NEG_COV_DIAG 2048, EIG_NOTFINITE 8192)
// (whole warp: lane i owns row i of L)
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
        LM_SYNC;
        LM_ONE L[j * n + j] = ljj;
        LM_FOR(i, j + 1, n) {
            double sum = cov[i * n + j];
#pragma unroll 1
            for (int k = 0; k < j; k++) sum -= L[i * n + k] * L[j * n + k];
            L[i * n + j] = ljj > 0.0 ? sum / ljj : 0.0;
        }
        LM_SYNC;
    }
    if (neg || (flags & 2048)) flags |= 1024;
    return flags;
}

}  // namespace ngx
