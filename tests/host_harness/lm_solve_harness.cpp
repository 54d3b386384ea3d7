// TEST INFRASTRUCTURE: drives the engine's LM state machine
// (ngmix_b200/csrc/ngx_lm_solve.cuh, compiled as plain C++) on the CPU, with
// the oracle's calc_fdiff / calc_jacobian standing in for the pixel passes.
// It checks that feeding only (J^T J, J^T f, |f|) reproduces MINPACK.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../ngmix_b200/csrc/ngx_lm_solve.cuh"
extern "C" {
#include "../../oracle/oracle.h"
}

extern "C" int harness_lm_fit(int model, const ora_pixel *pixels, int npix,
                              const double *psf_pars, int npsf,
                              const ora_prior *prior, const double *guess,
                              int maxfev, double ftol, double xtol,
                              double *pars_out, int *nfev, int *ier,
                              double *cov0, int *have_cov) {
    using namespace ngx;
    int np = ora_model_npars(model);
    int nprior = (prior && prior->present) ? (model == ORA_BD ? 7 : 6) : 0;
    int m = npix + nprior;
    std::vector<double> fvec(m), jac((size_t)m * np);
    LMState s;
    lm_init(&s, np, guess, maxfev, ftol, xtol);
    while (s.action != LM_ACT_DONE) {
        if (s.action == LM_ACT_EVAL_F) {
            int st = ora_lm_calc_fdiff(model, pixels, npix, psf_pars, npsf,
                                       prior, s.xtrial, fvec.data(), m);
            double fn;
            if (st != 0) {
                fn = NAN;
            } else {
                double ss = 0;
                for (int i = 0; i < m; i++) ss += fvec[i] * fvec[i];
                fn = std::sqrt(ss);
            }
            lm_feed_f(&s, fn);
        } else {
            // f at the accepted point and the jacobian there
            ora_lm_calc_fdiff(model, pixels, npix, psf_pars, npsf, prior, s.x,
                              fvec.data(), m);
            ora_lm_calc_jacobian(model, pixels, npix, psf_pars, npsf, prior,
                                 s.x, jac.data(), m);
            double A[LM_MAXP * LM_MAXP] = {0}, b[LM_MAXP] = {0};
            bool finite = std::isfinite(fvec[0]);
            for (int a = 0; a < np; a++) {
                for (int c = 0; c <= a; c++) {
                    double sum = 0;
                    for (int i = 0; i < m; i++)
                        sum += jac[i + (size_t)a * m] * jac[i + (size_t)c * m];
                    A[a + LM_MAXP * c] = sum;
                    A[c + LM_MAXP * a] = sum;
                }
                double sum = 0;
                if (finite)
                    for (int i = 0; i < m; i++)
                        sum += jac[i + (size_t)a * m] * fvec[i];
                b[a] = sum;
            }
            lm_feed_j(&s, A, b, 0);
        }
    }
    for (int i = 0; i < np; i++) pars_out[i] = s.x[i];
    *nfev = s.nfev;
    *ier = s.info;
    *have_cov = lm_cov0(&s, cov0);
    return 0;
}

extern "C" int harness_test_cov(int n, const double *cov) {
    double L[64];
    return ngx::lm_test_cov(n, cov, L);
}
