/*
 * TEST INFRASTRUCTURE - CPU restatement ("oracle") of the ngmix fitting hot
 * path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use this.  It is never linked into, imported by
 * or called from the product (ngmix_b200/), which fails loudly without its
 * CUDA library.
 *
 * Parity status: PINNED.  Every function here is checked in
 * tests/test_oracle_golden.py against vectors produced by running the
 * unmodified reference (/root/reference/ngmix, numba+scipy) in the build
 * container (oracle/gen_golden.py -> tests/golden/ npz files).
 *
 * Record layouts are byte-identical to the reference's numpy dtypes so the
 * Python wrapper passes structured arrays straight through:
 *   ora_gauss  = _gauss2d_dtype   ngmix/gmix/gmix.py:1196-1210   (104 B)
 *   ora_pixel  = _pixels_dtype    ngmix/pixels/pixels.py:72-79   (48 B)
 *   ora_coord  = _coords_dtype    ngmix/pixels/pixels.py:82-86   (24 B)
 *   ora_jac    = _jacobian_dtype  ngmix/jacobian/jacobian.py:406-414 (64 B)
 */
#ifndef NGMIX_ORACLE_H
#define NGMIX_ORACLE_H

#include <stdint.h>

typedef struct {
    double p, row, col, irr, irc, icc, det;
    int64_t norm_set;
    double drr, drc, dcc, norm, pnorm;
} ora_gauss;

typedef struct { double u, v, area, val, ierr, fdiff; } ora_pixel;
typedef struct { double u, v, area; } ora_coord;
typedef struct {
    double row0, col0, dvdrow, dvdcol, dudrow, dudcol, det, scale;
} ora_jac;

/* model ids = ngmix/gmix/gmix.py:1100-1110 */
enum {
    ORA_FULL = 0, ORA_GAUSS = 1, ORA_TURB = 2, ORA_EXP = 3, ORA_DEV = 4,
    ORA_BDF = 6, ORA_COELLIP = 7, ORA_CM = 9, ORA_BD = 10
};

/* status codes of oracle functions (exceptions of the reference) */
enum { ORA_OK = 0, ORA_RANGE_ERROR = 1, ORA_ZERO_DIV = 2 };

/* flag bits = ngmix/flags.py:3-23 */
enum {
    ORA_NO_ATTEMPT = 1, ORA_CEN_SHIFT = 2, ORA_NONPOS_FLUX = 4,
    ORA_NONPOS_SIZE = 8, ORA_LOW_DET = 16, ORA_MAXITER = 32,
    ORA_NONPOS_VAR = 64, ORA_GMIX_RANGE_ERROR = 128,
    ORA_NONPOS_SHAPE_VAR = 256, ORA_LM_SINGULAR_MATRIX = 512,
    ORA_LM_NEG_COV_EIG = 1024, ORA_LM_NEG_COV_DIAG = 2048,
    ORA_LM_FUNC_NOTFINITE = 4096, ORA_EIG_NOTFINITE = 8192,
    ORA_DIV_ZERO = 16384, ORA_ZERO_DOF = 32768
};

/* ---- fastexp (ngmix/fastexp_nb.py) ---- */
double ora_fexp(double x);
double ora_apod_window(double chi2);
double ora_apod_window_deriv(double chi2);
void ora_fexp_arr(int n, const double *x, double *out);
void ora_apod_arr(int n, const double *chi2, double *w, double *dw);

/* ---- mixtures (ngmix/gmix/gmix_nb.py) ---- */
int ora_model_ngauss(int model);
int ora_model_npars(int model);
int ora_gmix_fill(int model, const double *pars, int npars, ora_gauss *gmix,
                  int ngauss);
int ora_gmix_fill_cm(double fracdev, double TdByTe, double Tfactor,
                     const double *pars, ora_gauss *gmix);
double ora_get_cm_Tfactor(double fracdev, double TdByTe);
int ora_gmix_set_norms(ora_gauss *gmix, int n);
void ora_gmix_convolve_fill(ora_gauss *self, const ora_gauss *gmix, int n,
                            const ora_gauss *psf, int npsf);
int ora_g1g2_to_e1e2(double g1, double g2, double *e1, double *e2);

/* ---- pixels / jacobian ---- */
int ora_fill_pixels(ora_pixel *pixels, int npixels, const double *image,
                    const double *weight, int nrow, int ncol,
                    const ora_jac *jac, int ignore_zero_weight);
void ora_fill_coords(ora_coord *coords, int nrow, int ncol,
                     const ora_jac *jac);

/* ---- pixel passes ---- */
double ora_gmix_eval_fast(const ora_gauss *gmix, int n, double v, double u,
                          double area);
int ora_render(ora_gauss *gmix, int n, const ora_coord *coords, int ncoords,
               double *image, int fast_exp);
int ora_fill_fdiff(ora_gauss *gmix, int n, const ora_pixel *pixels, int npix,
                   double *fdiff, int start);
int ora_get_loglike(ora_gauss *gmix, int n, const ora_pixel *pixels, int npix,
                    double *out4);
int ora_get_model_s2n_sum(ora_gauss *gmix, int n, const ora_pixel *pixels,
                          int npix, double *s2n_sum);
void ora_deriv_images(const double *gpars, const double *dcov, int ngauss,
                      const double *vv, const double *uu, const double *area,
                      int npix, double *out);

/* ---- priors for the LM residual vector ---- */
enum { ORA_P_NONE = 0, ORA_P_FLAT = 1, ORA_P_ERF = 2, ORA_P_NORMAL = 3,
       ORA_P_LOGNORMAL = 4, ORA_P_TRUNCGAUSS = 5 };
typedef struct { int32_t kind; int32_t pad; double a, b, c, d; } ora_prior1d;
typedef struct {
    int32_t present;       /* 0: no prior */
    int32_t g_kind;        /* 1: GPriorBA(sigma) */
    double cen1, cen2, sigma1, sigma2;
    double g_sigma;
    ora_prior1d T;
    ora_prior1d fracdev;   /* bdf, bd */
    ora_prior1d F;
    ora_prior1d logTratio; /* bd only */
    int32_t ncoellip;      /* coellip fits: ngauss (read even without prior) */
    int32_t pad2;
} ora_prior;

/* ---- LM fit = Fitter.go (fitting/fitters.py:64-112) ---- */
#define ORA_MAXP 20     /* 7 shared + 12 band fluxes (multi-band bd); coellip needs 14;
                         * matrices are packed npars x npars */
typedef struct {
    int32_t flags, nfev, ier, npix, dof, njev;
    double lnprob, s2n_numer, s2n_denom, chi2per, s2n, T, T_err, flux,
        flux_err;
    double pars[ORA_MAXP], pars_err[ORA_MAXP], pars_cov0[ORA_MAXP * ORA_MAXP],
        pars_cov[ORA_MAXP * ORA_MAXP];
} ora_lm_result;

/* prior.bounds of the reference (joint_prior.py:37-61): type 0 none, 1 lower
 * bound only, 2 upper only, 3 both */
typedef struct {
    int32_t present;
    int32_t type[ORA_MAXP];
    double lo[ORA_MAXP], hi[ORA_MAXP];
} ora_bounds;

int ora_lm_fit_bounded(int model, const double *image, const double *weight,
                       int nrow, int ncol, const ora_jac *jac,
                       const double *psf_pars, int npsf, const double *guess,
                       const ora_prior *prior, const ora_bounds *bounds,
                       int maxfev, double ftol, double xtol,
                       int analytic_jacobian, int ignore_zero_weight,
                       ora_lm_result *res);
int ora_lm_fit(int model, const double *image, const double *weight,
               int nrow, int ncol, const ora_jac *jac, const double *psf_pars,
               int npsf, const double *guess, const ora_prior *prior,
               int maxfev, double ftol, double xtol, int analytic_jacobian,
               int ignore_zero_weight, ora_lm_result *res);

/* Fitter.go on an ObsList / MultiBandObsList: nstamp stamps of one object,
 * ordered band by band (bands[] non-decreasing), 5 + nband parameters,
 * Fp[nband] flux priors (used when prior->present) */
int ora_lm_fit_multi(int model, int nstamp, const double *images,
                     const double *weights, int nrow, int ncol,
                     const ora_jac *jacs, const double *psf_pars, int npsf,
                     const int32_t *bands, int nband, const double *guess,
                     const ora_prior *prior, const ora_prior1d *Fp, int maxfev,
                     double ftol, double xtol, int ignore_zero_weight,
                     ora_lm_result *res);
int ora_lm_fit_multi2(int model, int nstamp, const double *images,
                     const double *weights, int nrow, int ncol,
                     const ora_jac *jacs, const double *psf_pars, int npsf,
                     const int32_t *bands, int nband, const double *guess,
                     const ora_prior *prior, const ora_prior1d *Fp, int maxfev,
                     double ftol, double xtol, int analytic_jacobian, int ignore_zero_weight,
                     ora_lm_result *res);

int ora_lm_fit_multi_bounded(int model, int nstamp, const double *images,
                             const double *weights, int nrow, int ncol,
                             const ora_jac *jacs, const double *psf_pars,
                             int npsf, const int32_t *bands, int nband,
                             const double *guess, const ora_prior *prior,
                             const ora_prior1d *Fp, const ora_bounds *bounds,
                             int maxfev, double ftol, double xtol,
                             int analytic_jacobian, int ignore_zero_weight,
                             ora_lm_result *res);

/* residual vector and jacobian of FitModel (results.py:439-570) */
int ora_lm_calc_fdiff(int model, const ora_pixel *pixels, int npix,
                      const double *psf_pars, int npsf, const ora_prior *prior,
                      const double *pars, double *fdiff, int fdiff_size);
int ora_lm_calc_jacobian(int model, const ora_pixel *pixels, int npix,
                         const double *psf_pars, int npsf,
                         const ora_prior *prior, const double *pars,
                         double *jac, int fdiff_size);

/* batch drivers used by bench.py's cpu_baseline leg (threads over objects) */
int ora_lm_fit_batch(int n, int nthreads, int model, const double *images,
                     const double *weights, int nrow, int ncol,
                     const double *jacs, const double *psf_pars, int npsf,
                     const double *guess, const ora_prior *prior, int maxfev,
                     double ftol, double xtol, int analytic_jacobian,
                     ora_lm_result *res);

/* ---- MINPACK restatement (scipy.optimize._minpack lmder / lmdif) ---- */
typedef int (*ora_fcn)(void *ctx, int m, int n, const double *x,
                       double *fvec);
typedef int (*ora_jacfcn)(void *ctx, int m, int n, const double *x,
                          double *fjac /* m x n, column major */);
int ora_lmder(ora_fcn fcn, ora_jacfcn jfcn, void *ctx, int m, int n,
              double *x, double *fvec, double *fjac, double ftol, double xtol,
              double gtol, int maxfev, double factor, int *nfev, int *njev,
              int *ipvt, double *qtf);
int ora_lmdif(ora_fcn fcn, void *ctx, int m, int n, double *x, double *fvec,
              double *fjac, double ftol, double xtol, double gtol, int maxfev,
              double epsfcn, double factor, int *nfev, int *ipvt,
              double *qtf);

/* ---- EM (ngmix/em/em_nb.py) ---- */
typedef struct {
    double tol;
    int32_t maxiter, miniter;
    double sky;
    int32_t vary_sky, pad;
} ora_em_conf;
enum { ORA_EM_FULL = 0, ORA_EM_FIXCEN = 1, ORA_EM_FIXCOV = 2,
       ORA_EM_FLUXONLY = 3 };
int ora_em_run(int variant, const ora_em_conf *conf, ora_pixel *pixels,
               int npix, ora_gauss *gmix, int ngauss, ora_gauss *gmix_psf,
               int npsf, ora_gauss *gmix_conv, int fill_zero_weight,
               int32_t *numiter, double *frac_diff, double *sky);

/* EMFitter.go (em/em.py:228-317) on raw arrays */
typedef struct {
    int32_t flags, numiter;
    double fdiff, sky;
} ora_em_result;
int ora_em_go(int variant, const double *image, const double *weight, int nrow,
              int ncol, const ora_jac *jac, const double *guess_pars,
              int ngauss, const double *psf_pars, int npsf, double tol,
              int miniter, int maxiter, int vary_sky, int have_sky,
              double sky_in, int ignore_zero_weight, double *pars_out,
              ora_em_result *res);
int ora_em_go_batch(int n, int nthreads, int variant, const double *images,
                    const double *weights, int nrow, int ncol,
                    const double *jacs, const double *guess_pars, int ngauss,
                    double tol, int miniter, int maxiter, double *pars_out,
                    ora_em_result *res);

/* ---- admom (ngmix/admom/admom_nb.py) ---- */
typedef struct {
    int32_t maxiter, pad;
    double shiftmax, etol, Ttol;
    int32_t cenonly, pad2;
} ora_admom_conf;
typedef struct {
    int32_t flags, numiter, npix, pad;
    double wsum;
    double sums[7];
    double sums_cov[49];
    double pars[6];
    double rho4;
    double F[7];
} ora_admom_result;
void ora_admom(const ora_admom_conf *conf, ora_gauss *wt,
               const ora_pixel *pixels, int npix, ora_admom_result *res);
int ora_admom_go(const double *image, const double *weight, int nrow,
                 int ncol, const ora_jac *jac, const double *guess_pars,
                 const ora_admom_conf *conf, ora_admom_result *res,
                 double *wt_norm_out);
int ora_admom_go_batch(int n, int nthreads, const double *images,
                       const double *weights, int nrow, int ncol,
                       const double *jacs, const double *guess_pars,
                       const ora_admom_conf *conf, ora_admom_result *res);

/* ---- template (psf) flux: PSFFluxFitModel.go, fitting/results.py:676-805 ----
 * nstamp stamps of one object; gm_pars [nstamp, ngauss, 6] full parameters of
 * the template mixture of each stamp (the psf mixture for a psf flux) */
typedef struct {
    int32_t flags, pad;
    double flux, flux_err, chi2per, dof;
} ora_flux_result;
int ora_template_flux(int nstamp, const double *images, const double *weights,
                      int nrow, int ncol, const ora_jac *jacs,
                      const double *gm_pars, int ngauss, int normalize,
                      int ignore_zero_weight, ora_flux_result *res);

/* ---- weighted moments (ngmix/gmix/gmix_nb.py:681-821) ---- */
#define ORA_NMOM_MAX 17
typedef struct {
    int32_t flags, npix;
    double wsum;
    double sums[ORA_NMOM_MAX];
    double sums_cov[ORA_NMOM_MAX * ORA_NMOM_MAX];   /* [nmom][nmom] in the
                                                       leading nmom*nmom */
} ora_mom_result;
/* adds to res (the reference accumulates into the caller's record);
 * nmom = 6 (get_weighted_sums) or 17 (get_higher_order_weighted_sums) */
int ora_weighted_sums(ora_gauss *wt, int ngauss, const ora_pixel *pixels,
                      int npix, double maxrad, int nmom, ora_mom_result *res);
/* from image/weight/jacobian: fill_pixels + the sums, fresh record */
int ora_weighted_sums_go(const double *image, const double *weight, int nrow,
                         int ncol, const ora_jac *jac, const double *wt_pars,
                         int ngauss, double maxrad, int nmom,
                         int ignore_zero_weight, ora_mom_result *res);

#endif
