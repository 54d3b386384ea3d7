// The production fitting chain on the device, one pass over a batch:
//
//   psf stamps --EM (ntry)--> psf mixtures --template flux--> flux guesses
//   object stamps + psf mixtures --LM (ntry)--> fit records
//
// Replaces, for a batch, bootstrap (ngmix/bootstrap.py:43-102) with
// PSFRunner(EMFitter, GMixPSFGuesser) and Runner(Fitter, TPSFFluxGuesser)
// (runners.py:116-149, guessers.py:78-144, 205-261, 767-1051): guesses are
// drawn on the device from a counter-based generator, failed fits are retried
// with fresh guesses like run_fitter, and nothing returns to the host between
// the stages - every step is enqueued on one stream, and which objects a
// retry round refits is decided by a device-side mask.
//
// Random numbers: Philox4x32-10 keyed by the caller's 64-bit seed, counter
// (object index, try, draw block, stage).  Draw k of an (object, try, stage)
// is the k-th double of that stream, made from 2 x 32 bits exactly as numpy's
// random_sample does ((a >> 5) * 2^26 + (b >> 6)) / 2^53, and
// uniform(lo, hi) = lo + (hi - lo) * u.  The draws are made in the order of
// the reference guessers' rng.uniform calls, so "same draws given the same
// counter" is checkable on the host (tests/test_chain.py restates the
// generator in numpy).
#include "ngx_host.cuh"

namespace ngx {

#define NGX_STAGE_PSF 1u
#define NGX_STAGE_OBJ 2u

__device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1,
                                              unsigned c2, unsigned c3,
                                              unsigned k0, unsigned k1,
                                              unsigned out[4]) {
    const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const unsigned W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const unsigned hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const unsigned n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// the stream of one (object, try, stage): u(k), k = 0, 1, 2 ...
struct DrawStream {
    unsigned obj, tr, stage, k0, k1;
    unsigned block;         // cached block index, ~0u = none
    unsigned w[4];
    int next;
    __device__ DrawStream(unsigned long long seed, unsigned obj_, unsigned tr_,
                          unsigned stage_)
        : obj(obj_), tr(tr_), stage(stage_), k0((unsigned)seed),
          k1((unsigned)(seed >> 32)), block(~0u), next(0) {}
    __device__ double u01() {
        const unsigned b = (unsigned)next >> 1;
        if (b != block) {
            philox4x32_10(obj, tr, b, stage, k0, k1, w);
            block = b;
        }
        const int h = next & 1;
        next++;
        const unsigned a = w[2 * h] >> 5, c = w[2 * h + 1] >> 6;
        return ((double)a * 67108864.0 + (double)c) / 9007199254740992.0;
    }
    __device__ double uniform(double lo, double hi) {
        return lo + (hi - lo) * u01();
    }
};

// GMixPSFGuesser tables (guessers.py:1033-1051); the 4- and 5-gaussian
// recipes reuse entry 2 for the outer components, as the reference does
__constant__ double c_psf_p[6][5] = {
    {0, 0, 0, 0, 0}, {1, 0, 0, 0, 0},
    {0.596510042804182, 0.4034898268889178, 0, 0, 0},
    {0.596510042804182, 0.4034898268889178, 1.303069003078001e-07, 0, 0},
    {0.596510042804182, 0.4034898268889178, 1.303069003078001e-07, 1.0e-8, 0},
    {0.59453032, 0.35671819, 0.03567182, 0.01189061, 0.00118906}};
__constant__ double c_psf_f[6][5] = {
    {0, 0, 0, 0, 0}, {1, 0, 0, 0, 0},
    {0.5793612389470884, 1.621860687127999, 0, 0, 0},
    {0.5793612389470884, 1.621860687127999, 7.019347162356363, 0, 0},
    {0.5793612389470884, 1.621860687127999, 7.019347162356363, 16.0, 0},
    {0.5, 1.0, 3.0, 10.0, 20.0}};
__constant__ int c_psf_entry[6][5] = {
    {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 1, 0, 0, 0}, {0, 1, 2, 0, 0},
    {0, 1, 2, 2, 0}, {0, 1, 2, 2, 2}};

// sum of every pixel of each stamp (obs.image.sum(): masked pixels count)
__global__ void __launch_bounds__(128)
k_stamp_sums(int n, const double *__restrict__ images, int ntot,
             double *__restrict__ sums) {
    const int warps_per_block = blockDim.x >> 5;
    const int stamp = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    if (stamp >= n) return;
    const int lane = threadIdx.x & 31;
    const double *im = images + (size_t)stamp * ntot;
    // fixed association: lane partial sums in pixel order, then a butterfly
    double s = 0.0;
    for (int i = lane; i < ntot; i += 32) s += im[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) sums[stamp] = s;
}

// GMixPSFGuesser(rng, ngauss)(obs) for every stamp still to do
// (guessers.py:832-842 default T and flux, :879-1030 the draws)
__global__ void k_psf_guess(int n, const int32_t *__restrict__ todo,
                            unsigned long long seed, int tr, int ngauss,
                            const ngx_jacobian *__restrict__ jac,
                            const double *__restrict__ sums,
                            double *__restrict__ guess) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (todo && todo[i] == 0)) return;
    const double scale = jac[i].scale;
    // moments.fwhm_to_T: T = 2 (fwhm / 2.3548200450309493)^2
    const double sigma = scale * 3.5 / 2.3548200450309493;
    const double T = 2.0 * sigma * sigma;
    const double flux = sums[i];
    const double sigma2 = T / 2;
    DrawStream d(seed, (unsigned)i, (unsigned)tr, NGX_STAGE_PSF);
    double *g = guess + (size_t)i * ngauss * 6;
    if (ngauss == 1) {
        g[0] = flux * d.uniform(0.9, 1.1);
        g[1] = d.uniform(-0.1, 0.1);
        g[2] = d.uniform(-0.1, 0.1);
        g[3] = sigma2 * (1.0 + d.uniform(-0.1, 0.1));
        g[4] = d.uniform(-0.2 * sigma2, 0.2 * sigma2);
        g[5] = sigma2 * (1.0 + d.uniform(-0.1, 0.1));
        return;
    }
    for (int k = 0; k < ngauss; k++) {
        const int e = c_psf_entry[ngauss][k];
        const double pf = c_psf_p[ngauss][e], ff = c_psf_f[ngauss][e];
        double *q = g + 6 * k;
        q[0] = ngauss == 2 ? pf * flux
                           : flux * pf * (1.0 + d.uniform(-0.1, 0.1));
        q[1] = d.uniform(-0.1, 0.1);
        q[2] = d.uniform(-0.1, 0.1);
        q[3] = ff * sigma2 * (1.0 + d.uniform(-0.1, 0.1));
        q[4] = ngauss == 2 ? 0.0 : d.uniform(-0.01, 0.01);
        q[5] = ff * sigma2 * (1.0 + d.uniform(-0.1, 0.1));
    }
}

// TPSFFluxGuesser(rng, T)(obs) (guessers.py:78-144) for the simple models and
// its bdf / bd counterparts (BDFPSFFluxGuesser-style: fracdev ~ U(0.4, 0.6),
// logTratio ~ U(-0.1, 0.1)) for every object still to do.  T0: per-object
// central T (NULL: T of the psf mixture, sum p (irr + icc) / sum p).
__global__ void k_obj_guess(int n, const int32_t *__restrict__ todo,
                            unsigned long long seed, int tr, int model,
                            int np, const double *__restrict__ T0,
                            const double *__restrict__ psf_pars, int npsf,
                            const ngx_flux_result *__restrict__ pflux,
                            double *__restrict__ guess) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (todo && todo[i] == 0)) return;
    double T = 0.0;
    if (T0) {
        T = T0[i];
    } else {
        double ps = 0.0, ts = 0.0;
        for (int k = 0; k < npsf; k++) {
            const double *q = psf_pars + ((size_t)i * npsf + k) * 6;
            ps += q[0];
            ts += q[0] * (q[3] + q[5]);
        }
        T = ts / ps;
    }
    DrawStream d(seed, (unsigned)i, (unsigned)tr, NGX_STAGE_OBJ);
    double *g = guess + (size_t)i * np;
    g[0] = d.uniform(-0.01, 0.01);
    g[1] = d.uniform(-0.01, 0.01);
    g[2] = d.uniform(-0.02, 0.02);
    g[3] = d.uniform(-0.02, 0.02);
    g[4] = T * d.uniform(0.9, 1.1);
    int fi = 5;
    if (model == NGX_BD) {
        g[5] = d.uniform(-0.1, 0.1);
        g[6] = d.uniform(0.4, 0.6);
        fi = 7;
    } else if (model == NGX_BDF) {
        g[5] = d.uniform(0.4, 0.6);
        fi = 6;
    }
    g[fi] = pflux[i].flux * d.uniform(0.9, 1.1);
}

// after a psf round: objects whose fit succeeded are done
__global__ void k_todo_after_em(int n, int32_t *todo,
                                const ngx_em_result *__restrict__ res,
                                int32_t *ntry_used, int tr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || todo[i] == 0) return;
    ntry_used[i] = tr + 1;
    if (res[i].flags == 0) todo[i] = 0;
}

// before the object rounds: objects without a psf mixture or without a
// usable psf flux are not attempted (BootPSFFailure / PSFFluxFailure in the
// reference): their record carries NO_ATTEMPT
__global__ void k_todo_for_obj(int n, int32_t *todo, const int32_t *psf_todo,
                               const ngx_flux_result *__restrict__ pflux,
                               ngx_lm_result *res, int np) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool ok = psf_todo[i] == 0 && pflux[i].flags == 0
        && isfinite(pflux[i].flux);
    todo[i] = ok ? 1 : 0;
    if (!ok) {
        ngx_lm_result *r = &res[i];
        memset(r, 0, sizeof(*r));
        r->flags = NGX_NO_ATTEMPT;
        for (int k = 0; k < np; k++) {
            r->pars[k] = -9.999e9;
            r->pars_err[k] = 9.999e9;
        }
    }
}

__global__ void k_todo_after_lm(int n, int32_t *todo,
                                const ngx_lm_result *__restrict__ res,
                                int32_t *ntry_used, int tr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || todo[i] == 0) return;
    ntry_used[i] = tr + 1;
    if (res[i].flags == 0) todo[i] = 0;
}

__global__ void k_fill_i32(int n, int32_t *p, int32_t v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void k_iota_i32(int n, int32_t *p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

}  // namespace ngx

using namespace ngx;

extern "C" {

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
                  size_t workspace_bytes, void *stream) {
    NGX_REQUIRE(n >= 0 && psf_images && psf_weights && psf_jac && em_conf);
    NGX_REQUIRE(images && weights && jac && lm_pars && psf_pars && psf_res);
    NGX_REQUIRE(psf_flux && guess && res && ntry_used && workspace);
    NGX_REQUIRE(psf_ngauss >= 1 && psf_ngauss <= 5 && psf_ntry >= 1);
    NGX_REQUIRE(ntry >= 1);
    const int np = model == NGX_BD ? 8 : model == NGX_BDF ? 7 : 6;
    NGX_REQUIRE(model == NGX_GAUSS || model == NGX_EXP || model == NGX_DEV
                || model == NGX_TURB || model == NGX_BDF || model == NGX_BD);
    if (n == 0) return NGX_OK;
    // workspace: psf guesses, stamp sums, two masks, the iota of the
    // template-flux call
    const size_t need = ngx_bootstrap_workspace_bytes(n, psf_ngauss);
    NGX_REQUIRE(workspace_bytes >= need);
    char *w = (char *)workspace;
    double *psf_guess = (double *)w;
    w += sizeof(double) * (size_t)n * psf_ngauss * 6;
    double *sums = (double *)w;
    w += sizeof(double) * (size_t)n;
    int32_t *psf_todo = (int32_t *)w;
    w += sizeof(int32_t) * (size_t)n;
    int32_t *obj_todo = (int32_t *)w;
    w += sizeof(int32_t) * (size_t)n;
    int32_t *iota = (int32_t *)w;
    cudaStream_t st = as_stream(stream);
    const int B = 256, G = (n + B - 1) / B;

    // ---- psf stage: EM with up to psf_ntry guesses (run_psf_fitter)
    k_fill_i32<<<G, B, 0, st>>>(n, psf_todo, 1);
    k_fill_i32<<<(2 * n + B - 1) / B, B, 0, st>>>(2 * n, ntry_used, 0);
    k_stamp_sums<<<(n + 3) / 4, 128, 0, st>>>(n, psf_images,
                                              psf_nrow * psf_ncol, sums);
    NGX_LAUNCH_CHECK();
    for (int tr = 0; tr < psf_ntry; tr++) {
        k_psf_guess<<<G, B, 0, st>>>(n, psf_todo, seed, tr, psf_ngauss,
                                     psf_jac, sums, psf_guess);
        NGX_LAUNCH_CHECK();
        int rc = ngx_em_todo(n, NGX_EM_FULL, psf_images, psf_weights, psf_jac,
                             psf_nrow, psf_ncol, psf_guess, psf_ngauss,
                             nullptr, 0, em_conf, 1, psf_todo, psf_pars,
                             psf_res, stream);
        if (rc) return rc;
        k_todo_after_em<<<G, B, 0, st>>>(n, psf_todo, psf_res, ntry_used, tr);
        NGX_LAUNCH_CHECK();
    }
    // ---- psf flux of every object: the psf mixture as the template
    //      (PSFFluxFitter, fitting/results.py:703-767)
    k_iota_i32<<<(n + 1 + B - 1) / B, B, 0, st>>>(n + 1, iota);
    NGX_LAUNCH_CHECK();
    int rc = ngx_template_flux(n, iota, images, weights, jac, nrow, ncol,
                               psf_pars, psf_ngauss, 1, 1, psf_flux, stream);
    if (rc) return rc;
    // ---- object stage: LM with up to ntry guesses (run_fitter)
    k_todo_for_obj<<<G, B, 0, st>>>(n, obj_todo, psf_todo, psf_flux, res, np);
    NGX_LAUNCH_CHECK();
    for (int tr = 0; tr < ntry; tr++) {
        k_obj_guess<<<G, B, 0, st>>>(n, obj_todo, seed, tr, model, np,
                                     T_guess, psf_pars, psf_ngauss, psf_flux,
                                     guess);
        NGX_LAUNCH_CHECK();
        rc = ngx_lm_fit_todo(n, model, images, weights, jac, nrow, ncol,
                             psf_pars, psf_ngauss, guess, prior, lm_pars, 1,
                             obj_todo, res, stream);
        if (rc) return rc;
        k_todo_after_lm<<<G, B, 0, st>>>(n, obj_todo, res, ntry_used + n, tr);
        NGX_LAUNCH_CHECK();
    }
    return NGX_OK;
}

size_t ngx_bootstrap_workspace_bytes(int n, int psf_ngauss) {
    return sizeof(double) * ((size_t)n * psf_ngauss * 6 + (size_t)n)
        + sizeof(int32_t) * (3 * (size_t)n + 1) + 64;
}

// the first ndraw doubles of the stream (object i, try, stage), for tests of
// the "same draws given the same counter" contract: out [n, ndraw]
namespace ngx {
__global__ void k_chain_draws(int n, unsigned long long seed, int tr,
                              int stage, int ndraw, double *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    DrawStream d(seed, (unsigned)i, (unsigned)tr, (unsigned)stage);
    for (int k = 0; k < ndraw; k++) out[(size_t)i * ndraw + k] = d.u01();
}
}  // namespace ngx

int ngx_chain_draws(int n, unsigned long long seed, int tr, int stage,
                    int ndraw, double *out, void *stream) {
    NGX_REQUIRE(n >= 0 && ndraw >= 0 && out);
    if (n == 0 || ndraw == 0) return NGX_OK;
    k_chain_draws<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>(
        n, seed, tr, stage, ndraw, out);
    NGX_LAUNCH_CHECK();
    return NGX_OK;
}

}  // extern "C"
