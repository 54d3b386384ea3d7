/*
 * ngx.h - C ABI of the B200-native batched Gaussian-mixture fitting engine.
 *
 * This is the drop-in boundary for the fitting hot path of esheldon/ngmix.
 * The reference has no FFI: its boundary is "Python method -> @njit function
 * on numpy structured arrays".  Every entry point below names the @njit
 * function(s) (reference file:line) it replaces.  Record layouts are byte
 * identical to the reference's numpy dtypes, so a ctypes binding can hand the
 * reference's own arrays (GMix._data, Observation.pixels, Jacobian._data)
 * straight to the *_host entry points.
 *
 * Conventions
 *  - Every function processes a BATCH of n independent objects (stamps);
 *    n = 1 reproduces the reference's per-object call.
 *  - ngx_xxx(...)      : all array pointers are DEVICE pointers, work is
 *                        enqueued on `stream` (a cudaStream_t, NULL = legacy
 *                        default stream) and the call returns without
 *                        synchronising.  No allocation on the hot path.
 *  - ngx_xxx_host(...) : all array pointers are HOST pointers; the library
 *                        stages them through its own device buffers and
 *                        returns when the results are in the host arrays.
 *  - Small configuration structs (ngx_prior, ngx_lm_pars, ngx_em_conf,
 *    ngx_admom_conf) are always HOST pointers, passed by value to kernels.
 *  - Return value: NGX_OK or an NGX_ERR_* code for API misuse / CUDA errors
 *    (text via ngx_last_error()).  Numerical outcomes never use the return
 *    value: they are per-object int32 flags with the reference's bit values
 *    (ngmix/flags.py:3-23) or per-object status codes (NGX_ST_*), standing in
 *    for the GMixRangeError / ZeroDivisionError exceptions numba raises.
 *  - All arithmetic is IEEE fp64.  No exceptions cross this boundary.
 */
#ifndef NGX_H
#define NGX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NGX_VERSION 100

/* return codes */
#define NGX_OK 0
#define NGX_ERR_ARG 1       /* bad size / null pointer / unsupported model */
#define NGX_ERR_CUDA 2      /* a CUDA runtime call failed */
#define NGX_ERR_NODEVICE 3  /* no CUDA device: there is no CPU fallback */

/* per-object status (the reference's exceptions) */
#define NGX_ST_OK 0
#define NGX_ST_RANGE_ERROR 1   /* GMixRangeError, gexceptions.py:24 */
#define NGX_ST_ZERO_DIV 2      /* ZeroDivisionError raised by numba */
#define NGX_ST_FATAL 3         /* GMixFatalError("no weights > 0") */

/* model ids, ngmix/gmix/gmix.py:1100-1110 */
#define NGX_FULL 0
#define NGX_GAUSS 1
#define NGX_TURB 2
#define NGX_EXP 3
#define NGX_DEV 4
#define NGX_BDF 6
#define NGX_COELLIP 7
#define NGX_CM 9
#define NGX_BD 10

/* flag bits, ngmix/flags.py:3-23 */
#define NGX_NO_ATTEMPT 1
#define NGX_CEN_SHIFT 2
#define NGX_NONPOS_FLUX 4
#define NGX_NONPOS_SIZE 8
#define NGX_LOW_DET 16
#define NGX_MAXITER 32
#define NGX_NONPOS_VAR 64
#define NGX_GMIX_RANGE_ERROR 128
#define NGX_NONPOS_SHAPE_VAR 256
#define NGX_LM_SINGULAR_MATRIX 512
#define NGX_LM_NEG_COV_EIG 1024
#define NGX_LM_NEG_COV_DIAG 2048
#define NGX_LM_FUNC_NOTFINITE 4096
#define NGX_EIG_NOTFINITE 8192
#define NGX_DIV_ZERO 16384
#define NGX_ZERO_DOF 32768
#define NGX_EM_RANGE_ERROR NGX_GMIX_RANGE_ERROR
#define NGX_EM_MAXITER NGX_MAXITER

/* _gauss2d_dtype, ngmix/gmix/gmix.py:1196-1210 (104 bytes) */
typedef struct {
    double p, row, col, irr, irc, icc, det;
    int64_t norm_set;
    double drr, drc, dcc, norm, pnorm;
} ngx_gauss;

/* _pixels_dtype, ngmix/pixels/pixels.py:72-79 (48 bytes) */
typedef struct { double u, v, area, val, ierr, fdiff; } ngx_pixel;

/* _coords_dtype, ngmix/pixels/pixels.py:82-86 (24 bytes) */
typedef struct { double u, v, area; } ngx_coord;

/* _jacobian_dtype, ngmix/jacobian/jacobian.py:406-414 (64 bytes) */
typedef struct {
    double row0, col0, dvdrow, dvdcol, dudrow, dudcol, det, scale;
} ngx_jacobian;

/* ------------------------------------------------------------------ */
/* library / device management */
int ngx_version(void);
const char *ngx_last_error(void);
/* select the CUDA device for the calling thread; NGX_ERR_NODEVICE if none */
int ngx_init(int device);
/* sm count, shared memory per block, device name (name may be NULL) */
int ngx_device_info(int *sm_count, int *smem_optin, char *name, int name_len);

/* page-locked host memory for the *_host entry points (their chunked H2D
 * copies overlap the fits only from pinned memory): allocate pinned arrays,
 * or pin arrays the caller already owns for the duration of the calls */
int ngx_host_alloc(size_t bytes, void **out);
int ngx_host_free(void *p);
int ngx_host_register(void *p, size_t bytes);
int ngx_host_unregister(void *p);

/* ------------------------------------------------------------------ */
/* mixtures: replace gmix_fill_<model> (gmix_nb.py:307-558), gmix_set_norms
 * (:176-218), gmix_convolve_fill (:609-649).  status[n] receives NGX_ST_*.
 * pars is [n, npars]; gmix is [n, ngauss]. */
int ngx_gmix_fill(int n, int model, int ngauss, const double *pars, int npars,
                  ngx_gauss *gmix, int32_t *status, void *stream);
/* composite model with fixed fracdev, TdByTe (gmix_nb.py:430-466);
 * cm is [n, 3] = fracdev, TdByTe, Tfactor (get_cm_Tfactor :561-593) */
int ngx_gmix_fill_cm(int n, const double *cm, const double *pars,
                     ngx_gauss *gmix, int32_t *status, void *stream);
int ngx_gmix_set_norms(int n, int ngauss, ngx_gauss *gmix, int32_t *status,
                       void *stream);
int ngx_gmix_convolve(int n, int nobj, int npsf, const ngx_gauss *obj,
                      const ngx_gauss *psf, ngx_gauss *out, void *stream);

/* pixels: replace fill_pixels / fill_coords (pixels_nb.py:6-94) and
 * jacobian_get_vu (jacobian_nb.py:4-16).  pixels is [n, nrow*ncol]; the
 * first npix[i] records of object i are valid (zero-weight pixels are
 * dropped, in row-major order, when ignore_zero_weight != 0). */
int ngx_make_pixels(int n, int nrow, int ncol, const double *images,
                    const double *weights, const ngx_jacobian *jac,
                    int ignore_zero_weight, ngx_pixel *pixels, int32_t *npix,
                    void *stream);
int ngx_make_coords(int n, int nrow, int ncol, const ngx_jacobian *jac,
                    ngx_coord *coords, void *stream);

/* render: replaces render (render_nb.py:9-36) + fill_coords.  Adds the
 * mixture into images [n, nrow, ncol]; fast_exp != 0 uses the apodized
 * fexp evaluation (gmix_nb.py:28-63), else exact exp (gmix_nb.py:66-92). */
int ngx_render(int n, int ngauss, const ngx_gauss *gmix,
               const ngx_jacobian *jac, int nrow, int ncol, int fast_exp,
               double *images, int32_t *status, void *stream);

/* one fused pixel pass replacing fill_fdiff (gmix_nb.py:877-900),
 * get_loglike (:824-874) and get_model_s2n_sum (:903-937):
 *   out4 [n,4] = loglike, s2n_numer, s2n_denom, npix
 *   fdiff [n, nrow*ncol] (nullable) = (model-val)*ierr for the valid pixels
 *   of object i, compacted in the reference's pixel order. */
int ngx_loglike(int n, int ngauss, const ngx_gauss *gmix, const double *images,
                const double *weights, const ngx_jacobian *jac, int nrow,
                int ncol, int ignore_zero_weight, double *out4, double *fdiff,
                int32_t *status, void *stream);

/* same pass on a pre-built pixel array (the reference's own calling
 * convention: fill_fdiff(gmix, pixels, fdiff, start)); pixels [n, maxpix],
 * npix[n] valid records each; fdiff [n, maxpix] nullable */
int ngx_loglike_pixels(int n, int ngauss, const ngx_gauss *gmix,
                       const ngx_pixel *pixels, int maxpix,
                       const int32_t *npix, double *out4, double *fdiff,
                       int32_t *status, void *stream);

/* deriv_images (fitting/derivs_nb.py:40-127): gpars [n,ngauss,6], dcov
 * [n,ngauss,3,3], evaluated on the full nrow x ncol grid of jac;
 * out [n,6,nrow*ncol] is ADDED to (zero it first, as the reference does) */
int ngx_deriv_images(int n, int ngauss, const double *gpars,
                     const double *dcov, const ngx_jacobian *jac, int nrow,
                     int ncol, double *out, void *stream);

/* calc_noise_cov (fitting/noise_cov.py:90-137), the matrix B of the sandwich
 * covariance pars_cov0 B pars_cov0 of a fit under stationary correlated noise:
 * B_ab = sum_q conj(G_a) G_b |fft2(noise)|^2 / N^2 with G_a = fft2(weight *
 * dmodel/dp_a), evaluated in real space (Parseval) as the Gram matrix of the
 * circular cross-correlations of weight * dmodel/dp_a with the noise image.
 * dimages [n,nd,nrow*ncol] holds the derivative images, parameter a uses
 * image rows[a] (NULL: a) times scale[n,nk] (NULL: 1); work [n,nk,nrow*ncol]
 * is scratch; B [n,nk,nk]; nk <= 8. */
int ngx_noise_b(int n, int nk, const int32_t *rows, const double *dimages,
                int nd, const double *scale, const double *weights,
                const double *noise, int nrow, int ncol, double *work,
                double *B, void *stream);

/* ------------------------------------------------------------------ */
/* priors of the LM residual vector (PriorSimpleSep joint_prior.py:5-236,
 * PriorBDFSep :484-676) over CenPrior, GPriorBA, FlatPrior, TwoSidedErf,
 * Normal */
#define NGX_P_NONE 0
#define NGX_P_FLAT 1      /* a=minval b=maxval */
#define NGX_P_ERF 2       /* a=minval b=width_at_min c=maxval d=width_at_max */
#define NGX_P_NORMAL 3    /* a=mean b=sigma */
#define NGX_P_LOGNORMAL 4 /* a=logmean b=logivar c=lnprob_max d=shift
                           * (priors/priors.py:674-840) */
#define NGX_P_TRUNCGAUSS 5 /* a=mean b=sigma c=minval d=maxval
                            * (priors/priors.py:1046-1150) */
typedef struct { int32_t kind; int32_t pad; double a, b, c, d; } ngx_prior1d;
typedef struct {
    int32_t present;      /* 0 = fit without prior */
    int32_t g_kind;       /* 1 = GPriorBA(g_sigma) */
    double cen1, cen2, sigma1, sigma2;   /* CenPrior */
    double g_sigma;
    ngx_prior1d T;
    ngx_prior1d fracdev;  /* bdf only */
    ngx_prior1d F;        /* flux prior (band 0) */
    int32_t nband;        /* 0 or 1: single band; 2, 3: ngx_lm_fit_multi */
    int32_t ncoellip;     /* ngauss of PriorCoellipSame (ngx_lm_fit_coellip) */
    ngx_prior1d F_more[2];    /* flux priors of bands 1 and 2 */
    ngx_prior1d logTratio;    /* bd only: PriorBDSep, joint_prior.py:267-399 */
    ngx_prior1d F_extra[4];   /* flux priors of bands 3..6 (nband up to 7) */
    ngx_prior1d F_xwide[5];   /* flux priors of bands 7..11
                               * (ngx_lm_fit_multi_xwide, nband up to 12) */
} ngx_prior;

/* DEFAULT_LM_PARS, ngmix/defaults.py:17 */
#define NGX_MAX_PARS 20
#define NGX_MAX_BANDS 12
/* bounds kinds of one parameter (prior.bounds of the reference,
 * joint_prior.py:37-61; transforms of leastsqbound.py:183-263) */
#define NGX_B_NONE 0
#define NGX_B_LOWER 1     /* (min, None) */
#define NGX_B_UPPER 2     /* (None, max) */
#define NGX_B_BOTH 3      /* (min, max) */
typedef struct {
    int32_t maxfev;             /* 4000 */
    int32_t analytic_jacobian;  /* Fitter(analytic_jacobian=True) */
    double ftol, xtol;          /* 1e-5, 1e-5 */
    /* bounded fit (run_leastsq(bounds=prior.bounds), leastsqbound.py:445-552):
     * MINPACK iterates on internal, unconstrained parameters; residuals are
     * evaluated at the external ones.  nbounds = 0: unbounded fit. */
    int32_t nbounds;            /* 0 or the number of parameters */
    int32_t btype[NGX_MAX_PARS];
    int32_t pad_;
    double blo[NGX_MAX_PARS], bhi[NGX_MAX_PARS];
} ngx_lm_pars;

/* result record of one fit: the keys of run_leastsq
 * (fitting/leastsqbound.py:118-126) + FitModel.set_fit_result
 * (fitting/results.py:45-72).  Matrices are row major npars x npars. */
typedef struct {
    int32_t flags, nfev, ier, npix, dof, njev;
    double lnprob, s2n_numer, s2n_denom, chi2per, s2n, T, T_err, flux,
        flux_err;
    double pars[8], pars_err[8], pars_cov0[64], pars_cov[64];
} ngx_lm_result;

/* the same record with room for 14 parameters: multi-band fits with 4 to 7
 * bands (ngx_lm_fit_multi_wide), coelliptical fits (ngx_lm_fit_coellip) */
typedef struct {
    int32_t flags, nfev, ier, npix, dof, njev;
    double lnprob, s2n_numer, s2n_denom, chi2per, s2n, T, T_err, flux,
        flux_err;
    double pars[14], pars_err[14], pars_cov0[196], pars_cov[196];
} ngx_lm_result_wide;

/* ... and for 20: multi-band fits with 8 to 12 bands
 * (ngx_lm_fit_multi_xwide) */
typedef struct {
    int32_t flags, nfev, ier, npix, dof, njev;
    double lnprob, s2n_numer, s2n_denom, chi2per, s2n, T, T_err, flux,
        flux_err;
    double pars[20], pars_err[20], pars_cov0[400], pars_cov[400];
} ngx_lm_result_xwide;

/* Fitter(model, prior, fit_pars).go(obs, guess) for n stamps
 * (fitting/fitters.py:64-112): FitModel.calc_fdiff/calc_jacobian
 * (results.py:439-625), the MINPACK lmder/lmdif iteration reached through
 * scipy.optimize.leastsq (leastsqbound.py:445-447), run_leastsq's epilogue
 * (:70-126) and set_fit_result, all on the device.
 * images/weights [n,nrow,ncol]; psf_pars [n,npsf,6] full gaussian pars of
 * obs.psf.gmix (npsf = 0: no psf); guess [n,npars]. */
int ngx_lm_fit(int n, int model, const double *images, const double *weights,
               const ngx_jacobian *jac, int nrow, int ncol,
               const double *psf_pars, int npsf, const double *guess,
               const ngx_prior *prior, const ngx_lm_pars *lm_pars,
               int ignore_zero_weight, ngx_lm_result *res, void *stream);

/* ------------------------------------------------------------------ */
/* Fitter.go on ObsList / MultiBandObsList input (fitting/results.py:212-347,
 * 439-570: every epoch of every band is a stamp of the object; the residual
 * vector stacks the prior rows and the stamps' pixel rows, the parameters are
 * (cen1, cen2, g1, g2, T, F_band0 .. F_band{nband-1})).  Object i owns stamps
 * stamp_start[i] .. stamp_start[i+1]-1 of the stamp arrays, ordered band by
 * band; stamp_band[s] is the band of stamp s; guess is [nobj, 5 + nband].
 * Simple models (gauss, exp, dev) with the analytic Jacobian, nband <= 3.
 * The result record is the single-stamp one: pars / pars_cov carry 5 + nband
 * entries, flux / flux_err are those of band 0. */
int ngx_lm_fit_multi(int nobj, int model, const int32_t *stamp_start,
                     const int32_t *stamp_band, int nband,
                     const double *images, const double *weights,
                     const ngx_jacobian *jac, int nrow, int ncol,
                     const double *psf_pars, int npsf, const double *guess,
                     const ngx_prior *prior, const ngx_lm_pars *lm_pars,
                     int ignore_zero_weight, ngx_lm_result *res, void *stream);
/* CoellipFitter.go (fitting/fitters.py:120-141, results.py:628-674): ngauss
 * coelliptical gaussians, guess [n, 4 + 2 ngauss] = (cen1, cen2, g1, g2,
 * T_1..T_n, F_1..F_n), forward-difference Jacobian, optional
 * PriorCoellipSame (joint_prior.py:874-1031: cen, g, the T prior on every T_i,
 * the F prior on every F_i); ngauss <= 5 */
int ngx_lm_fit_coellip(int n, int ngauss, const double *images,
                       const double *weights, const ngx_jacobian *jac,
                       int nrow, int ncol, const double *psf_pars, int npsf,
                       const double *guess, const ngx_prior *prior,
                       const ngx_lm_pars *lm_pars, int ignore_zero_weight,
                       ngx_lm_result_wide *res, void *stream);
int ngx_lm_fit_coellip_host(int n, int ngauss, const double *images,
                            const double *weights, const ngx_jacobian *jac,
                            int nrow, int ncol, const double *psf_pars,
                            int npsf, const double *guess,
                            const ngx_prior *prior,
                            const ngx_lm_pars *lm_pars,
                            int ignore_zero_weight,
                            ngx_lm_result_wide *res);

/* the same for 4 to 7 bands (up to 12 parameters, 14-slot records); 1 to 3
 * bands are accepted too */
int ngx_lm_fit_multi_wide(int nobj, int model, const int32_t *stamp_start,
                     const int32_t *stamp_band, int nband,
                     const double *images, const double *weights,
                     const ngx_jacobian *jac, int nrow, int ncol,
                     const double *psf_pars, int npsf, const double *guess,
                     const ngx_prior *prior, const ngx_lm_pars *lm_pars,
                     int ignore_zero_weight, ngx_lm_result_wide *res, void *stream);

/* ... and for 8 to 12 bands (the reference has no limit on the number of
 * bands, results.py:212-260; up to 19 parameters, 20-slot records); fewer
 * bands are accepted too */
int ngx_lm_fit_multi_xwide(int nobj, int model, const int32_t *stamp_start,
                     const int32_t *stamp_band, int nband,
                     const double *images, const double *weights,
                     const ngx_jacobian *jac, int nrow, int ncol,
                     const double *psf_pars, int npsf, const double *guess,
                     const ngx_prior *prior, const ngx_lm_pars *lm_pars,
                     int ignore_zero_weight, ngx_lm_result_xwide *res, void *stream);

/* EM: EMFitter*.go (em/em.py:228-317) = prep_image (:95-120) + em_run /
 * em_run_fixcen / em_run_fixcov / em_run_fluxonly (em/em_nb.py:15-1257) */
#define NGX_EM_FULL 0
#define NGX_EM_FIXCEN 1
#define NGX_EM_FIXCOV 2
#define NGX_EM_FLUXONLY 3
typedef struct {
    double tol;          /* 1e-5 */
    int32_t maxiter;     /* 500 */
    int32_t miniter;     /* 40 (20 for fluxonly) */
    double sky;          /* used when have_sky != 0 */
    int32_t vary_sky;
    int32_t have_sky;    /* 0: run prep_image to choose the sky */
} ngx_em_conf;
typedef struct {
    int32_t flags, numiter;
    double fdiff, sky;
} ngx_em_result;
/* guess [n,ngauss,6] full pars before psf convolution; psf_pars [n,npsf,6]
 * (npsf = 0 -> delta psf); pars_out [n,ngauss,6] */
int ngx_em(int n, int variant, const double *images, const double *weights,
           const ngx_jacobian *jac, int nrow, int ncol, const double *guess,
           int ngauss, const double *psf_pars, int npsf,
           const ngx_em_conf *conf, int ignore_zero_weight, double *pars_out,
           ngx_em_result *res, void *stream);

/* admom (admom/admom_nb.py:13-243); result = _admom_result_dtype
 * (admom/admom.py:571-583, aligned) */
typedef struct {
    int32_t maxiter;     /* 200 */
    int32_t cenonly;
    double shiftmax;     /* 5.0 */
    double etol;         /* 1e-5 */
    double Ttol;         /* 1e-3 */
} ngx_admom_conf;
typedef struct {
    int32_t flags, numiter, npix, pad;
    double wsum;
    double sums[7];
    double sums_cov[49];
    double pars[6];
    double rho4;
    double F[7];
    double wt_norm;      /* wt_gmix["norm"][0] at exit, for get_result */
} ngx_admom_result;
/* guess [n,6] full pars of the 1-gaussian weight */
int ngx_admom(int n, const double *images, const double *weights,
              const ngx_jacobian *jac, int nrow, int ncol, const double *guess,
              const ngx_admom_conf *conf, ngx_admom_result *res, void *stream);

/* weighted moments: get_weighted_sums / get_higher_order_weighted_sums
 * (gmix/gmix_nb.py:681-821), the pass behind GMix.get_weighted_sums
 * (gmix/gmix.py:706-754) and GaussMom.go (gaussmom.py:49-73).
 * gmix [n,ngauss] is the weight mixture (norms are set on a private copy; a
 * range error sets flags = NGX_GMIX_RANGE_ERROR), maxrad [n] the radius cut
 * in (v,u) units, nmom 6 or 17.  sums_cov holds [nmom][nmom] row major in its
 * leading nmom*nmom entries (the layout of get_moments_result_dtype). */
#define NGX_NMOM_MAX 17
typedef struct {
    int32_t flags, npix;
    double wsum;
    double sums[NGX_NMOM_MAX];
    double sums_cov[NGX_NMOM_MAX * NGX_NMOM_MAX];
} ngx_mom_result;
int ngx_weighted_sums(int n, int ngauss, const ngx_gauss *gmix,
                      const double *images, const double *weights,
                      const ngx_jacobian *jac, int nrow, int ncol,
                      const double *maxrad, int nmom, int ignore_zero_weight,
                      ngx_mom_result *res, void *stream);

/* template (psf) flux: PSFFluxFitModel.go (fitting/results.py:676-805),
 * zero-lag cross-correlation of the exact-exp render of each stamp's template
 * mixture with the image, chi^2 at the fitted flux.  Object i owns stamps
 * stamp_start[i] .. stamp_start[i+1]-1; gm_pars [nstamps, ngauss, 6] are the
 * full parameters of each stamp's template (the psf mixture for a psf flux);
 * normalize != 0 sets each template's flux to 1 first. */
typedef struct {
    int32_t flags, pad;    /* DIV_ZERO 2^14, BAD_VAR 2^6, GMIX_RANGE_ERROR */
    double flux, flux_err, chi2per, dof;
} ngx_flux_result;
int ngx_template_flux(int nobj, const int32_t *stamp_start,
                      const double *images, const double *weights,
                      const ngx_jacobian *jac, int nrow, int ncol,
                      const double *gm_pars, int ngauss, int normalize,
                      int ignore_zero_weight, ngx_flux_result *res,
                      void *stream);

/* the same with pixel templates [ns, nrow, ncol] in place of mixtures
 * (obs.psf.template / obs.template, results.py:834-864) */
int ngx_template_flux_images(int nobj, const int32_t *stamp_start,
                             const double *images, const double *weights,
                             const double *templates, int nrow, int ncol,
                             int normalize, int ignore_zero_weight,
                             ngx_flux_result *res, void *stream);
int ngx_template_flux_images_host(int nobj, const int32_t *stamp_start,
                                  const double *images, const double *weights,
                                  const double *templates, int nrow, int ncol,
                                  int normalize, int ignore_zero_weight,
                                  ngx_flux_result *res);

/* ------------------------------------------------------------------ */
/* The production chain in one device pass (bootstrap.py:43-102 with
 * PSFRunner(EMFitter, GMixPSFGuesser(ngauss), ntry) and Runner(Fitter(model),
 * TPSFFluxGuesser(T), ntry); runners.py:116-149, guessers.py:78-144, 205-261,
 * 767-1051): per object an EM fit of its psf stamp, the psf-template flux of
 * its image (PSFFluxFitter), a guess drawn from that flux and T, and the LM
 * fit, each stage retried with fresh guesses while flags != 0.  Guesses come
 * from Philox4x32-10 keyed by `seed` with counter (object index, try, draw
 * block, stage): they depend on nothing but (seed, index, try), so results
 * do not change with batch size or sharding (ngx_chain_draws exposes the
 * stream).  All device pointers; everything is enqueued on `stream`.
 *   T_guess [n] central T of the object guesses (NULL: T of the psf mixture)
 *   psf_pars [n,psf_ngauss,6], psf_res [n], psf_flux [n], guess [n,npars],
 *   res [n], ntry_used [2,n] (psf, object) are outputs;
 *   workspace of ngx_bootstrap_workspace_bytes(n, psf_ngauss) bytes.
 * Objects whose psf fit or psf flux fails get res.flags = NGX_NO_ATTEMPT. */
size_t ngx_bootstrap_workspace_bytes(int n, int psf_ngauss);
int ngx_bootstrap(int n, const double *psf_images, const double *psf_weights,
                  const ngx_jacobian *psf_jac, int psf_nrow, int psf_ncol,
                  int psf_ngauss, int psf_ntry, const ngx_em_conf *em_conf,
                  const double *images, const double *weights,
                  const ngx_jacobian *jac, int nrow, int ncol, int model,
                  const double *T_guess, int ntry, const ngx_prior *prior,
                  const ngx_lm_pars *lm_pars, unsigned long long seed,
                  double *psf_pars, ngx_em_result *psf_res,
                  ngx_flux_result *psf_flux, double *guess,
                  ngx_lm_result *res, int32_t *ntry_used, void *workspace,
                  size_t workspace_bytes, void *stream);
/* the first ndraw uniform doubles of stream (object i, try, stage) for
 * i < n: out [n, ndraw]; stage 1 = psf guesses, 2 = object guesses */
int ngx_chain_draws(int n, unsigned long long seed, int tr, int stage,
                    int ndraw, double *out, void *stream);
/* ngx_em / ngx_lm_fit restricted to the objects with todo[i] != 0 (device
 * array; the records of the others are left untouched): the retry rounds */
int ngx_em_todo(int n, int variant, const double *images,
                const double *weights, const ngx_jacobian *jac, int nrow,
                int ncol, const double *guess, int ngauss,
                const double *psf_pars, int npsf, const ngx_em_conf *conf,
                int ignore_zero_weight, const int32_t *todo, double *pars_out,
                ngx_em_result *res, void *stream);
int ngx_lm_fit_todo(int n, int model, const double *images,
                    const double *weights, const ngx_jacobian *jac, int nrow,
                    int ncol, const double *psf_pars, int npsf,
                    const double *guess, const ngx_prior *prior,
                    const ngx_lm_pars *lm_pars, int ignore_zero_weight,
                    const int32_t *todo, ngx_lm_result *res, void *stream);

/* ------------------------------------------------------------------ */
/* host-buffer variants (synchronous; H2D/D2H inside) */
int ngx_gmix_fill_host(int n, int model, int ngauss, const double *pars,
                       int npars, ngx_gauss *gmix, int32_t *status);
int ngx_gmix_fill_cm_host(int n, const double *cm, const double *pars,
                          ngx_gauss *gmix, int32_t *status);
int ngx_gmix_set_norms_host(int n, int ngauss, ngx_gauss *gmix,
                            int32_t *status);
int ngx_gmix_convolve_host(int n, int nobj, int npsf, const ngx_gauss *obj,
                           const ngx_gauss *psf, ngx_gauss *out);
int ngx_make_pixels_host(int n, int nrow, int ncol, const double *images,
                         const double *weights, const ngx_jacobian *jac,
                         int ignore_zero_weight, ngx_pixel *pixels,
                         int32_t *npix);
int ngx_make_coords_host(int n, int nrow, int ncol, const ngx_jacobian *jac,
                         ngx_coord *coords);
int ngx_render_host(int n, int ngauss, const ngx_gauss *gmix,
                    const ngx_jacobian *jac, int nrow, int ncol, int fast_exp,
                    double *images, int32_t *status);
int ngx_loglike_host(int n, int ngauss, const ngx_gauss *gmix,
                     const double *images, const double *weights,
                     const ngx_jacobian *jac, int nrow, int ncol,
                     int ignore_zero_weight, double *out4, double *fdiff,
                     int32_t *status);
int ngx_loglike_pixels_host(int n, int ngauss, const ngx_gauss *gmix,
                            const ngx_pixel *pixels, int maxpix,
                            const int32_t *npix, double *out4, double *fdiff,
                            int32_t *status);
int ngx_deriv_images_host(int n, int ngauss, const double *gpars,
                          const double *dcov, const ngx_jacobian *jac,
                          int nrow, int ncol, double *out);
/* B from caller-supplied derivative images [n,nk,nrow*ncol] (the
 * central-difference images of models without analytic derivatives,
 * noise_cov.py:203-226) */
int ngx_noise_b_host(int n, int nk, const double *dimages,
                     const double *weights, const double *noise, int nrow,
                     int ncol, double *B);
/* B [n,6,6] in parameter order (cen1, cen2, g1, g2, T, flux) for the models
 * with analytic derivative images (noise_cov.py:160-201): gpars / dcov as for
 * ngx_deriv_images, flux [n] */
int ngx_noise_cov_host(int n, int ngauss, const double *gpars,
                       const double *dcov, const double *flux,
                       const ngx_jacobian *jac, const double *weights,
                       const double *noise, int nrow, int ncol, double *B);
int ngx_lm_fit_host(int n, int model, const double *images,
                    const double *weights, const ngx_jacobian *jac, int nrow,
                    int ncol, const double *psf_pars, int npsf,
                    const double *guess, const ngx_prior *prior,
                    const ngx_lm_pars *lm_pars, int ignore_zero_weight,
                    ngx_lm_result *res);
/* the same with the stamps as floats on the host (survey pixels are float32 on
 * disk; the reference widens them in Observation.__init__): half the PCIe
 * bytes of the call, widened exactly on the device */
int ngx_lm_fit_host_f32(int n, int model, const float *images,
                        const float *weights, const ngx_jacobian *jac,
                        int nrow, int ncol, const double *psf_pars, int npsf,
                        const double *guess, const ngx_prior *prior,
                        const ngx_lm_pars *lm_pars, int ignore_zero_weight,
                        ngx_lm_result *res);
/* the compact survey format: float32 pixels, ONE weight per stamp and an
 * optional good-pixel mask of one bit per pixel (bit k of a stamp = pixel k
 * in row-major order, most significant bit of each byte first, as
 * numpy.packbits writes it; NULL = every pixel good).  It stands for the
 * Observation the reference would build from image.astype(f8) and
 * weight = where(mask, weight_scalar, 0) (observation.py:114-126,
 * pixels/pixels.py:6-52: ierr = sqrt(weight), zero-weight pixels dropped);
 * 4.1 instead of 16 bytes per pixel cross PCIe and the results are
 * bit-identical to ngx_lm_fit_host on the expanded arrays. */
int ngx_lm_fit_host_compact(int n, int model, const float *images,
                            const double *weight_scalar,
                            const unsigned char *maskbits,
                            const ngx_jacobian *jac, int nrow, int ncol,
                            const double *psf_pars, int npsf,
                            const double *guess, const ngx_prior *prior,
                            const ngx_lm_pars *lm_pars,
                            int ignore_zero_weight, ngx_lm_result *res);
/* device-pointer expansion of the compact format into the engine's float64
 * planes (ObservationBatch.from_arrays with float32 pixels, one weight per
 * stamp and a mask): images32 [n,nrow,ncol], weight_scalar [n], maskbits
 * [n, ceil(nrow*ncol/8)] or NULL -> images, weights [n,nrow,ncol] */
int ngx_expand_compact(int n, int nrow, int ncol, const float *images32,
                       const double *weight_scalar,
                       const unsigned char *maskbits, double *images,
                       double *weights, void *stream);
int ngx_em_host(int n, int variant, const double *images,
                const double *weights, const ngx_jacobian *jac, int nrow,
                int ncol, const double *guess, int ngauss,
                const double *psf_pars, int npsf, const ngx_em_conf *conf,
                int ignore_zero_weight, double *pars_out, ngx_em_result *res);
int ngx_admom_host(int n, const double *images, const double *weights,
                   const ngx_jacobian *jac, int nrow, int ncol,
                   const double *guess, const ngx_admom_conf *conf,
                   ngx_admom_result *res);
int ngx_weighted_sums_host(int n, int ngauss, const ngx_gauss *gmix,
                           const double *images, const double *weights,
                           const ngx_jacobian *jac, int nrow, int ncol,
                           const double *maxrad, int nmom,
                           int ignore_zero_weight, ngx_mom_result *res);
int ngx_template_flux_host(int nobj, const int32_t *stamp_start,
                           const double *images, const double *weights,
                           const ngx_jacobian *jac, int nrow, int ncol,
                           const double *gm_pars, int ngauss, int normalize,
                           int ignore_zero_weight, ngx_flux_result *res);
int ngx_lm_fit_multi_host(int nobj, int model, const int32_t *stamp_start,
                          const int32_t *stamp_band, int nband,
                          const double *images, const double *weights,
                          const ngx_jacobian *jac, int nrow, int ncol,
                          const double *psf_pars, int npsf,
                          const double *guess, const ngx_prior *prior,
                          const ngx_lm_pars *lm_pars, int ignore_zero_weight,
                          ngx_lm_result *res);
int ngx_lm_fit_multi_wide_host(int nobj, int model, const int32_t *stamp_start,
                          const int32_t *stamp_band, int nband,
                          const double *images, const double *weights,
                          const ngx_jacobian *jac, int nrow, int ncol,
                          const double *psf_pars, int npsf,
                          const double *guess, const ngx_prior *prior,
                          const ngx_lm_pars *lm_pars, int ignore_zero_weight,
                          ngx_lm_result_wide *res);

int ngx_lm_fit_multi_xwide_host(int nobj, int model, const int32_t *stamp_start,
                          const int32_t *stamp_band, int nband,
                          const double *images, const double *weights,
                          const ngx_jacobian *jac, int nrow, int ncol,
                          const double *psf_pars, int npsf,
                          const double *guess, const ngx_prior *prior,
                          const ngx_lm_pars *lm_pars, int ignore_zero_weight,
                          ngx_lm_result_xwide *res);

/* FP64 pipe micro-benchmark used for the roofline denominator: runs a
 * register-resident chain of `iters` x 8 independent DFMAs per thread on a
 * full grid and returns the measured fp64 FMA instructions (lane ops) per
 * second in *dfma_per_s.  Device pointers are not needed. */
int ngx_peak_dfma(int iters, double *dfma_per_s, double *ms);
/* FP64 tensor-core probe: warp-level m8n8k4 DMMA instructions per second;
 * mode 1 interleaves 8 DFMA chains per thread (pipe-sharing check). */
int ngx_peak_dmma(int iters, int mode, double *dmma_per_s, double *ms);

#ifdef __cplusplus
}
#endif
#endif /* NGX_H */
