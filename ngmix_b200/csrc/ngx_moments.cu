// ngx_weighted_sums: gaussian-mixture weighted moments of a stamp.
//
// Replaces get_weighted_sums / get_higher_order_weighted_sums
// (ngmix/gmix/gmix_nb.py:681-821), the single pixel pass behind
// GMix.get_weighted_sums / get_weighted_moments (gmix/gmix.py:674-754) and
// GaussMom.go (gaussmom.py:49-73).  One CTA per stamp; pixels are built on
// the fly from (image, weight, jacobian) exactly as fill_pixels does
// (pixels_nb.py:6-58), the weight mixture is evaluated with the exact
// exponential (gmix_eval_pixel, gmix_nb.py:95-105).
//
// Sums: nmom first moments F_i weighted by w * val, their covariance
// sum w^2 var F_i F_j (upper triangle, mirrored on output), wsum and npix.
// 6 moments need 29 accumulators and one pixel pass; 17 moments need 172,
// taken in seven row groups of at most 32 accumulators so that they stay in
// registers (the single weight gaussian is cheap to re-evaluate).  Reductions
// are halving butterflies in a fixed order: deterministic.
#include "ngx_core.cuh"
#include "ngx_host.cuh"

namespace ngx {

// the moment basis (gmix_nb.py:713-718, 784-807); v, u are offsets from the
// weight centre except in F[0], F[1], which are the pixel's own coordinates
template <int NMOM>
__device__ __forceinline__ void mom_basis(double pv, double pu, double v,
                                          double u, double r2,
                                          double (&F)[NMOM]) {
    F[0] = pv;
    F[1] = pu;
    if (NMOM == 6) {
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
}

// number of upper-triangle entries of covariance rows [I0, I1)
__host__ __device__ constexpr int tri_rows(int nmom, int i0, int i1) {
    int c = 0;
    for (int i = i0; i < i1; i++) c += nmom - i;
    return c;
}

struct MomArgs {
    int ngauss, nrow, ncol, ignore_zero_weight;
    const ngx_gauss *gmix;      // normalised copy in shared memory
    const double *im, *wt;
    ngx_jacobian jac;
    double maxrad2;
};

// One pixel pass accumulating covariance rows [I0, I1) (upper triangle)
// and, with FIRST, the nmom sums, wsum and npix.  Results go to shared
// memory: out[0 .. NACC).
template <int NMOM, int I0, int I1, bool FIRST>
__device__ __forceinline__ void mom_pass(const MomArgs &a, double *scratch,
                                         double *out) {
    constexpr int NCOV = tri_rows(NMOM, I0, I1);
    constexpr int NACC = NCOV + (FIRST ? NMOM + 2 : 0);
    static_assert(NACC <= 32, "row group too large");
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = 0.0;
    const int ntot = a.nrow * a.ncol;
    const double area = a.jac.scale * a.jac.scale;
    const double vcen = a.gmix[0].row, ucen = a.gmix[0].col;
    // four pixels of a thread at a time with their loads issued together (the
    // weight -> test -> pixel value chain per pixel left the pass waiting on
    // HBM latency: 5.3 long-scoreboard stalls per issue); the order in which
    // a thread adds its pixels is unchanged
    constexpr int PF = NMOM == 6 ? 4 : 1;
    for (int p0 = threadIdx.x; p0 < ntot; p0 += PF * blockDim.x) {
      double wpre[PF], dpre[PF];
#pragma unroll
      for (int c = 0; c < PF; c++) {
          const int pp = p0 + c * blockDim.x;
          wpre[c] = pp < ntot ? a.wt[pp] : 0.0;
          dpre[c] = (FIRST && pp < ntot) ? a.im[pp] : 0.0;
      }
#pragma unroll
      for (int c = 0; c < PF; c++) {
        const int p = p0 + c * blockDim.x;
        if (p >= ntot) break;
        double ivar = wpre[c];
        if (a.ignore_zero_weight && ivar <= 0.0) continue;   // fill_pixels
        if (ivar < 0.0) ivar = 0.0;
        double ierr = sqrt(ivar);
        int row = p / a.ncol, col = p - row * a.ncol;
        double rowdiff = row - a.jac.row0, coldiff = col - a.jac.col0;
        double pv = a.jac.dvdrow * rowdiff + a.jac.dvdcol * coldiff;
        double pu = a.jac.dudrow * rowdiff + a.jac.dudcol * coldiff;
        double v = pv - vcen, u = pu - ucen;
        double r2 = u * u + v * v;
        if (!(r2 < a.maxrad2)) continue;
        // get_weighted_sums also requires ierr > 0 (:708); the higher-order
        // version does not (:773)
        if (NMOM == 6 && !(ierr > 0.0)) continue;
        double weight = 0.0;
        for (int k = 0; k < a.ngauss; k++)
            weight += eval_exact(a.gmix[k], pv, pu, area);
        double var = 1.0 / (ierr * ierr);
        double F[NMOM];
        mom_basis<NMOM>(pv, pu, v, u, r2, F);
        double w2v = (weight * weight) * var;
        int q = 0;
#pragma unroll
        for (int i = I0; i < I1; i++) {
            double wf = w2v * F[i];
#pragma unroll
            for (int j = i; j < NMOM; j++) {
                acc[q] += wf * F[j];
                q++;
            }
        }
        if (FIRST) {
            double wdata = weight * dpre[c];
#pragma unroll
            for (int i = 0; i < NMOM; i++) acc[NCOV + i] += wdata * F[i];
            acc[NCOV + NMOM] += weight;
            acc[NCOV + NMOM + 1] += 1.0;
        }
      }
    }
    block_sum_to<NACC>(acc, scratch, out);
}

// copy a reduced row group into the result record (thread 0)
template <int NMOM, int I0, int I1, bool FIRST>
__device__ __forceinline__ void mom_store(const double *tot,
                                          ngx_mom_result *r) {
    constexpr int NCOV = tri_rows(NMOM, I0, I1);
    int q = 0;
    for (int i = I0; i < I1; i++)
        for (int j = i; j < NMOM; j++) {
            r->sums_cov[i * NMOM + j] = tot[q];
            r->sums_cov[j * NMOM + i] = tot[q];
            q++;
        }
    if (FIRST) {
        for (int i = 0; i < NMOM; i++) r->sums[i] = tot[NCOV + i];
        r->wsum = tot[NCOV + NMOM];
        r->npix = (int)(tot[NCOV + NMOM + 1] + 0.5);
    }
}

template <int NMOM, int I0, int I1, bool FIRST>
__device__ __forceinline__ void mom_group(const MomArgs &a, double *scratch,
                                          double *tot, ngx_mom_result *r) {
    mom_pass<NMOM, I0, I1, FIRST>(a, scratch, tot);
    if (threadIdx.x == 0) mom_store<NMOM, I0, I1, FIRST>(tot, r);
    __syncthreads();
}

template <int NMOM>
__global__ void __launch_bounds__(128)
k_weighted_sums(int ngauss, const ngx_gauss *__restrict__ gmix,
                const double *__restrict__ images,
                const double *__restrict__ weights,
                const ngx_jacobian *__restrict__ jac, int nrow, int ncol,
                const double *__restrict__ maxrad, int ignore_zero_weight,
                ngx_mom_result *__restrict__ res) {
    extern __shared__ double smem_d[];
    double *scratch = smem_d;                       // 32 * nwarps
    double *tot = scratch + 32 * (blockDim.x >> 5); // 32
    ngx_gauss *gm = (ngx_gauss *)(tot + 32);        // ngauss
    __shared__ int bad;
    const int obj = blockIdx.x;
    const size_t ntot = (size_t)nrow * ncol;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    for (int k = threadIdx.x; k < ngauss; k += blockDim.x) {
        ngx_gauss g = gmix[(size_t)obj * ngauss + k];
        if (gauss_set_norm(g) != NGX_ST_OK) bad = 1;
        gm[k] = g;
    }
    __syncthreads();
    ngx_mom_result *r = &res[obj];
    if (threadIdx.x == 0) {
        r->flags = bad ? NGX_GMIX_RANGE_ERROR : 0;
        r->npix = 0;
        r->wsum = 0.0;
    }
    for (int i = threadIdx.x; i < NGX_NMOM_MAX; i += blockDim.x)
        r->sums[i] = 0.0;
    for (int i = threadIdx.x; i < NGX_NMOM_MAX * NGX_NMOM_MAX;
         i += blockDim.x)
        r->sums_cov[i] = 0.0;
    if (bad) return;
    MomArgs a;
    a.ngauss = ngauss; a.nrow = nrow; a.ncol = ncol;
    a.ignore_zero_weight = ignore_zero_weight;
    a.gmix = gm;
    a.im = images + obj * ntot;
    a.wt = weights + obj * ntot;
    a.jac = jac[obj];
    a.maxrad2 = maxrad[obj] * maxrad[obj];
    __syncthreads();
    if (NMOM == 6) {
        mom_group<6, 0, 6, true>(a, scratch, tot, r);      // 21 + 8
    } else {
        mom_group<17, 0, 1, false>(a, scratch, tot, r);    // 17
        mom_group<17, 1, 2, false>(a, scratch, tot, r);    // 16
        mom_group<17, 2, 4, false>(a, scratch, tot, r);    // 29
        mom_group<17, 4, 6, false>(a, scratch, tot, r);    // 25
        mom_group<17, 6, 9, false>(a, scratch, tot, r);    // 30
        mom_group<17, 9, 13, false>(a, scratch, tot, r);   // 26
        mom_group<17, 13, 17, true>(a, scratch, tot, r);   // 10 + 19
    }
}

// ---------------------------------------------------------------------
// ngx_template_flux: PSFFluxFitModel.go (fitting/results.py:702-770).  One
// CTA per object; the object's stamps are walked twice: zero-lag
// cross-correlation of the exact-exp render of each stamp's template mixture
// with the image (all pixels; zero weights drop out), then chi^2 at the
// fitted flux.
struct FluxShared {
    ngx_gauss gm[16];
    double tot[8];
    double flux, norm;
    int bad;
};

__device__ __forceinline__ void flux_set_mixture(FluxShared *sh,
                                                 const double *pars, int ngauss,
                                                 int normalize, int ipass,
                                                 double flux) {
    // one thread: GMix(pars=...), set_flux(1.0) (gmix.py:299-311), and for
    // the second pass set_flux(flux * norm)
    double p[16], psum = 0.0;
    for (int k = 0; k < ngauss; k++) { p[k] = pars[6 * k]; psum += p[k]; }
    if (normalize) {
        double rat = 1.0 / psum;
        for (int k = 0; k < ngauss; k++) p[k] *= rat;
    }
    double nsum = 0.0;
    for (int k = 0; k < ngauss; k++) nsum += p[k];
    if (ipass == 1) {
        sh->norm = nsum;
    } else {
        double rat = (flux * nsum) / nsum;
        for (int k = 0; k < ngauss; k++) p[k] *= rat;
    }
    sh->bad = 0;
    for (int k = 0; k < ngauss; k++) {
        ngx_gauss g;
        const double *q = pars + 6 * k;
        gauss_set(g, p[k], q[1], q[2], q[3], q[4], q[5]);
        if (gauss_set_norm(g) != NGX_ST_OK) sh->bad = 1;
        sh->gm[k] = g;
    }
}

__global__ void __launch_bounds__(128)
k_template_flux(const int32_t *__restrict__ stamp_start,
                const double *__restrict__ images,
                const double *__restrict__ weights,
                const ngx_jacobian *__restrict__ jac, int nrow, int ncol,
                const double *__restrict__ gm_pars, int ngauss, int normalize,
                int ignore_zero_weight, ngx_flux_result *__restrict__ res) {
    __shared__ FluxShared sh;
    __shared__ double scratch[32 * 4];
    const int obj = blockIdx.x;
    const int s0 = stamp_start[obj], s1 = stamp_start[obj + 1];
    const int ntot = nrow * ncol;
    double xcorr = 0.0, msq = 0.0, chi2 = 0.0, totpix = 0.0, effpix = 0.0;
    int bad = 0;
    for (int ipass = 1; ipass <= 2 && !bad; ipass++) {
        for (int e = s0; e < s1; e++) {
            __syncthreads();
            if (threadIdx.x == 0)
                flux_set_mixture(&sh, gm_pars + (size_t)e * 6 * ngauss, ngauss,
                                 normalize, ipass, sh.flux);
            __syncthreads();
            if (sh.bad) { bad = 1; break; }
            const ngx_jacobian j = jac[e];
            const double area = j.scale * j.scale;
            const double *im = images + (size_t)e * ntot;
            const double *wt = weights + (size_t)e * ntot;
            double acc[5] = {0, 0, 0, 0, 0};
            for (int p = threadIdx.x; p < ntot; p += blockDim.x) {
                int row = p / ncol, col = p - row * ncol;
                double rowdiff = row - j.row0, coldiff = col - j.col0;
                double v = j.dvdrow * rowdiff + j.dvdcol * coldiff;
                double u = j.dudrow * rowdiff + j.dudcol * coldiff;
                double model = 0.0;
                for (int k = 0; k < ngauss; k++)
                    model += eval_exact(sh.gm[k], v, u, area);
                double d = im[p], w = wt[p];
                if (ipass == 1) {
                    acc[0] += model * d * w;
                    acc[1] += model * model * w;
                    if (!(ignore_zero_weight && w <= 0.0)) acc[2] += 1.0;
                    if (w > 0) acc[3] += 1.0;
                } else {
                    double diff = model - d;
                    acc[4] += diff * diff * w;
                }
            }
            block_sum_to<5>(acc, scratch, sh.tot);
            xcorr += sh.tot[0]; msq += sh.tot[1];
            totpix += sh.tot[2]; effpix += sh.tot[3];
            chi2 += sh.tot[4];
        }
        if (ipass == 1) {
            if (msq == 0) break;
            __syncthreads();
            if (threadIdx.x == 0) sh.flux = xcorr / msq;
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        ngx_flux_result r;
        r.pad = 0;
        int flags = 0;
        double flux = -9.999e9, flux_err = 9.999e9;
        if (bad) {
            flags = NGX_GMIX_RANGE_ERROR;
            r.chi2per = 9999.0;
            r.dof = 0.0;
        } else {
            if (msq != 0) flux = xcorr / msq;
            double dof = effpix - 1;
            if (dof <= 0) dof = 1.0e-6;
            r.chi2per = chi2 / dof;
            r.dof = dof;
            if (msq == 0 || totpix == 1) {
                flags |= 1 << 14;               // DIV_ZERO
            } else {
                double arg = chi2 / msq / (totpix - 1);
                if (arg >= 0.0) flux_err = sqrt(arg);
                else flags |= 1 << 6;           // BAD_VAR
            }
        }
        r.flags = flags;
        r.flux = flux;
        r.flux_err = flux_err;
        res[obj] = r;
    }
}

// the same with pixel templates instead of mixtures (PSFFluxFitModel with
// obs.psf.template / obs.template, results.py:758-767, 834-864): the model of
// a stamp is its template image, normalised to unit sum when asked
// (`template *= 1 / template.sum()`), and in the second pass
// `template * (norm * flux) / template.sum()`.
__global__ void __launch_bounds__(128)
k_template_flux_images(const int32_t *__restrict__ stamp_start,
                       const double *__restrict__ images,
                       const double *__restrict__ weights,
                       const double *__restrict__ templates, int nrow,
                       int ncol, int normalize, int ignore_zero_weight,
                       ngx_flux_result *__restrict__ res) {
    __shared__ double tot[8];
    __shared__ double scratch[32 * 4];
    __shared__ double flux_s;
    const int obj = blockIdx.x;
    const int s0 = stamp_start[obj], s1 = stamp_start[obj + 1];
    const int ntot = nrow * ncol;
    double xcorr = 0.0, msq = 0.0, chi2 = 0.0, totpix = 0.0, effpix = 0.0;
    for (int ipass = 1; ipass <= 2; ipass++) {
        for (int e = s0; e < s1; e++) {
            const double *im = images + (size_t)e * ntot;
            const double *wt = weights + (size_t)e * ntot;
            const double *tp = templates + (size_t)e * ntot;
            // template.sum(), then the sum of the normalised template
            double a1[1] = {0.0};
            for (int p = threadIdx.x; p < ntot; p += blockDim.x) a1[0] += tp[p];
            block_sum_to<1>(a1, scratch, tot);
            const double tsum = tot[0];
            const double s1n = normalize ? 1.0 / tsum : 1.0;
            const double norm = normalize ? 1.0 : tsum;
            __syncthreads();
            a1[0] = 0.0;
            for (int p = threadIdx.x; p < ntot; p += blockDim.x)
                a1[0] += tp[p] * s1n;
            block_sum_to<1>(a1, scratch, tot);
            const double msum = tot[0];
            __syncthreads();
            const double s2 = ipass == 2 ? (norm * flux_s) / msum : 1.0;
            double acc[5] = {0, 0, 0, 0, 0};
            for (int p = threadIdx.x; p < ntot; p += blockDim.x) {
                double model = tp[p] * s1n;
                if (ipass == 2) model *= s2;
                double d = im[p], w = wt[p];
                if (ipass == 1) {
                    acc[0] += model * d * w;
                    acc[1] += model * model * w;
                    if (!(ignore_zero_weight && w <= 0.0)) acc[2] += 1.0;
                    if (w > 0) acc[3] += 1.0;
                } else {
                    double diff = model - d;
                    acc[4] += diff * diff * w;
                }
            }
            block_sum_to<5>(acc, scratch, tot);
            xcorr += tot[0]; msq += tot[1];
            totpix += tot[2]; effpix += tot[3];
            chi2 += tot[4];
            __syncthreads();
        }
        if (ipass == 1) {
            if (msq == 0) break;
            if (threadIdx.x == 0) flux_s = xcorr / msq;
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        ngx_flux_result r;
        r.pad = 0;
        int flags = 0;
        double flux = -9.999e9, flux_err = 9.999e9;
        if (msq != 0) flux = xcorr / msq;
        double dof = effpix - 1;
        if (dof <= 0) dof = 1.0e-6;
        r.chi2per = chi2 / dof;
        r.dof = dof;
        if (msq == 0 || totpix == 1) {
            flags |= 1 << 14;               // DIV_ZERO
        } else {
            double arg = chi2 / msq / (totpix - 1);
            if (arg >= 0.0) flux_err = sqrt(arg);
            else flags |= 1 << 6;           // BAD_VAR
        }
        r.flags = flags;
        r.flux = flux;
        r.flux_err = flux_err;
        res[obj] = r;
    }
}

}  // namespace ngx

using namespace ngx;

extern "C" {

int ngx_template_flux_images(int nobj, const int32_t *stamp_start,
                             const double *images, const double *weights,
                             const double *templates, int nrow, int ncol,
                             int normalize, int ignore_zero_weight,
                             ngx_flux_result *res, void *stream) {
    NGX_REQUIRE(nobj >= 0 && stamp_start && images && weights && templates);
    NGX_REQUIRE(res && nrow >= 1 && ncol >= 1);
    if (nobj == 0) return NGX_OK;
    k_template_flux_images<<<nobj, 128, 0, as_stream(stream)>>>(
        stamp_start, images, weights, templates, nrow, ncol, normalize,
        ignore_zero_weight, res);
    NGX_LAUNCH_CHECK();
    return NGX_OK;
}

int ngx_template_flux_images_host(int nobj, const int32_t *stamp_start,
                                  const double *images, const double *weights,
                                  const double *templates, int nrow, int ncol,
                                  int normalize, int ignore_zero_weight,
                                  ngx_flux_result *res) {
    NGX_REQUIRE(nobj >= 0 && stamp_start && images && weights && templates);
    NGX_REQUIRE(res);
    if (nobj == 0) return NGX_OK;
    const int ns = stamp_start[nobj];
    const size_t ntot = (size_t)nrow * ncol;
    DevBuf<int32_t> dss(nobj + 1, stamp_start);
    DevBuf<double> di(ns * ntot, images), dw(ns * ntot, weights);
    DevBuf<double> dt(ns * ntot, templates);
    DevBuf<ngx_flux_result> dr(nobj);
    NGX_CUDA(dss.err); NGX_CUDA(di.err); NGX_CUDA(dw.err); NGX_CUDA(dt.err);
    NGX_CUDA(dr.err);
    int rc = ngx_template_flux_images(nobj, dss.d, di.d, dw.d, dt.d, nrow,
                                      ncol, normalize, ignore_zero_weight,
                                      dr.d, 0);
    if (rc) return rc;
    NGX_CUDA(cudaDeviceSynchronize());
    NGX_CUDA(dr.to_host(res));
    return NGX_OK;
}

int ngx_weighted_sums(int n, int ngauss, const ngx_gauss *gmix,
                      const double *images, const double *weights,
                      const ngx_jacobian *jac, int nrow, int ncol,
                      const double *maxrad, int nmom, int ignore_zero_weight,
                      ngx_mom_result *res, void *stream) {
    NGX_REQUIRE(n >= 0 && ngauss >= 1 && ngauss <= NGX_MAX_GAUSS);
    NGX_REQUIRE(nrow >= 1 && ncol >= 1 && (nmom == 6 || nmom == 17));
    NGX_REQUIRE(gmix && images && weights && jac && maxrad && res);
    if (n == 0) return NGX_OK;
    const int threads = 128;
    size_t smem = sizeof(double) * (32 * (threads / 32) + 32)
        + sizeof(ngx_gauss) * (size_t)ngauss;
    if (nmom == 6)
        k_weighted_sums<6><<<n, threads, smem, as_stream(stream)>>>(
            ngauss, gmix, images, weights, jac, nrow, ncol, maxrad,
            ignore_zero_weight, res);
    else
        k_weighted_sums<17><<<n, threads, smem, as_stream(stream)>>>(
            ngauss, gmix, images, weights, jac, nrow, ncol, maxrad,
            ignore_zero_weight, res);
    NGX_LAUNCH_CHECK();
    return NGX_OK;
}

int ngx_weighted_sums_host(int n, int ngauss, const ngx_gauss *gmix,
                           const double *images, const double *weights,
                           const ngx_jacobian *jac, int nrow, int ncol,
                           const double *maxrad, int nmom,
                           int ignore_zero_weight, ngx_mom_result *res) {
    NGX_REQUIRE(n >= 0 && gmix && images && weights && jac && maxrad && res);
    if (n == 0) return NGX_OK;
    size_t ntot = (size_t)n * nrow * ncol;
    DevBuf<ngx_gauss> dg((size_t)n * ngauss, gmix);
    DevBuf<double> di(ntot, images), dw(ntot, weights), dm(n, maxrad);
    DevBuf<ngx_jacobian> dj(n, jac);
    DevBuf<ngx_mom_result> dr(n);
    NGX_CUDA(dg.err); NGX_CUDA(di.err); NGX_CUDA(dw.err); NGX_CUDA(dm.err);
    NGX_CUDA(dj.err); NGX_CUDA(dr.err);
    int rc = ngx_weighted_sums(n, ngauss, dg.d, di.d, dw.d, dj.d, nrow, ncol,
                               dm.d, nmom, ignore_zero_weight, dr.d, 0);
    if (rc) return rc;
    NGX_CUDA(cudaDeviceSynchronize());
    NGX_CUDA(dr.to_host(res));
    return NGX_OK;
}

int ngx_template_flux(int nobj, const int32_t *stamp_start,
                      const double *images, const double *weights,
                      const ngx_jacobian *jac, int nrow, int ncol,
                      const double *gm_pars, int ngauss, int normalize,
                      int ignore_zero_weight, ngx_flux_result *res,
                      void *stream) {
    NGX_REQUIRE(nobj >= 0 && stamp_start && images && weights && jac);
    NGX_REQUIRE(gm_pars && res && ngauss >= 1 && ngauss <= 16);
    NGX_REQUIRE(nrow >= 1 && ncol >= 1);
    if (nobj == 0) return NGX_OK;
    k_template_flux<<<nobj, 128, 0, as_stream(stream)>>>(
        stamp_start, images, weights, jac, nrow, ncol, gm_pars, ngauss,
        normalize, ignore_zero_weight, res);
    NGX_LAUNCH_CHECK();
    return NGX_OK;
}

int ngx_template_flux_host(int nobj, const int32_t *stamp_start,
                           const double *images, const double *weights,
                           const ngx_jacobian *jac, int nrow, int ncol,
                           const double *gm_pars, int ngauss, int normalize,
                           int ignore_zero_weight, ngx_flux_result *res) {
    NGX_REQUIRE(nobj >= 0 && stamp_start && images && weights && jac);
    NGX_REQUIRE(gm_pars && res);
    if (nobj == 0) return NGX_OK;
    const int ns = stamp_start[nobj];
    const size_t ntot = (size_t)nrow * ncol;
    DevBuf<int32_t> dss(nobj + 1, stamp_start);
    DevBuf<double> di(ns * ntot, images), dw(ns * ntot, weights);
    DevBuf<ngx_jacobian> dj(ns, jac);
    DevBuf<double> dg((size_t)ns * ngauss * 6, gm_pars);
    DevBuf<ngx_flux_result> dr(nobj);
    NGX_CUDA(dss.err); NGX_CUDA(di.err); NGX_CUDA(dw.err); NGX_CUDA(dj.err);
    NGX_CUDA(dg.err); NGX_CUDA(dr.err);
    int rc = ngx_template_flux(nobj, dss.d, di.d, dw.d, dj.d, nrow, ncol,
                               dg.d, ngauss, normalize, ignore_zero_weight,
                               dr.d, 0);
    if (rc) return rc;
    NGX_CUDA(cudaDeviceSynchronize());
    NGX_CUDA(dr.to_host(res));
    return NGX_OK;
}

}  // extern "C"
