// ngx_noise_b / ngx_noise_cov_host: the B matrix of the noise-power sandwich
// covariance of an LM fit, calc_noise_cov (ngmix/fitting/noise_cov.py:90-137):
//
//     B_ab = sum_q conj(G_a(q)) G_b(q) |n~(q)|^2 / N^2,   G_a = fft2(w d_a)
//
// with d_a the derivative image of the model with respect to parameter a, w
// the weight map, n~ the fft of the epoch's noise image and N the number of
// pixels.  With C_a = G_a conj(n~) the summand is conj(C_a) C_b, and C_a is the
// transform of the circular cross-correlation
//
//     c_a[dx, dy] = sum_{r,c} n[r, c] k_a[(r + dx) % H, (c + dy) % W],  k_a = w d_a
//
// so by Parseval B_ab = (1/N) sum_x c_a[x] c_b[x]: no transform is needed.  The
// correlation is N^2 multiply-adds per parameter (3.2e7 for six parameters at
// 48x48), done from shared memory with a sliding register window; at stamp
// sizes a direct sum is within 1e-12 of the transform route.  The caller forms
// pars_cov0 B pars_cov0.
#include "ngx_core.cuh"
#include "ngx_host.cuh"

namespace ngx {

#define NZ_JB 4          // column shifts per thread block of shifts
#define NZ_MAXK 8        // parameters per stamp

struct NoiseMap {
    int rows[NZ_MAXK];   // which image of D is parameter a's derivative
};

// grid (n, nk): one CTA correlates one derivative image with the noise image.
// GM: the two staged planes live in a per-CTA slice of a global scratch
// buffer (stamps too large for shared memory).
template <bool GM>
__global__ void __launch_bounds__(256)
k_noise_corr(int nk, NoiseMap map, const double *__restrict__ D, int nd,
             const double *__restrict__ scale,
             const double *__restrict__ weights,
             const double *__restrict__ noise, int nrow, int ncol, int wp,
             double *__restrict__ C, double *gscratch) {
    extern __shared__ double smem_d[];
    const int a = blockIdx.y, obj = blockIdx.x;
    const int ntot = nrow * ncol;
    const int p2 = 2 * wp + 1;                 // odd pitch: rows spread over banks
    double *k2 = GM ? gscratch + ((size_t)obj * nk + a)
        * ((size_t)nrow * p2 + (size_t)nrow * wp) : smem_d;
    // k2: nrow * p2 (k, columns repeated)
    double *n2 = k2 + (size_t)nrow * p2;       // nrow * wp: noise, zero padded
    const double *d = D + ((size_t)obj * nd + map.rows[a]) * ntot;
    const double *w = weights + (size_t)obj * ntot;
    const double *nz = noise + (size_t)obj * ntot;
    const double sc = scale ? scale[(size_t)obj * nk + a] : 1.0;
    for (int i = threadIdx.x; i < nrow * p2; i += blockDim.x) {
        int r = i / p2, cc = i - r * p2;
        int c = cc % ncol;
        k2[i] = w[r * ncol + c] * (d[r * ncol + c] * sc);
    }
    for (int i = threadIdx.x; i < nrow * wp; i += blockDim.x) {
        int r = i / wp, c = i - r * wp;
        n2[i] = c < ncol ? nz[r * ncol + c] : 0.0;
    }
    __syncthreads();
    double *out = C + ((size_t)obj * nk + a) * ntot;
    const int nblk = nrow * (wp / NZ_JB);      // (row shift, block of column shifts)
    for (int blk = threadIdx.x; blk < nblk; blk += blockDim.x) {
        const int jb = blk / nrow, dx = blk - jb * nrow;
        const int dy0 = jb * NZ_JB;
        double acc[NZ_JB];
#pragma unroll
        for (int j = 0; j < NZ_JB; j++) acc[j] = 0.0;
        for (int r = 0; r < nrow; r++) {
            int rr = r + dx;
            if (rr >= nrow) rr -= nrow;
            const double *krow = k2 + (size_t)rr * p2 + dy0;
            const double *nr = n2 + (size_t)r * wp;
            for (int c0 = 0; c0 < wp; c0 += NZ_JB) {
                double kw[2 * NZ_JB - 1], nv[NZ_JB];
#pragma unroll
                for (int j = 0; j < 2 * NZ_JB - 1; j++) kw[j] = krow[c0 + j];
#pragma unroll
                for (int j = 0; j < NZ_JB; j++) nv[j] = nr[c0 + j];
#pragma unroll
                for (int jc = 0; jc < NZ_JB; jc++)
#pragma unroll
                    for (int j = 0; j < NZ_JB; j++)
                        acc[j] = fma(nv[jc], kw[jc + j], acc[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < NZ_JB; j++)
            if (dy0 + j < ncol) out[dx * ncol + dy0 + j] = acc[j];
    }
}

// grid n: B[a][b] = (1/N) sum_x c_a[x] c_b[x]
__global__ void __launch_bounds__(256)
k_noise_gram(int nk, const double *__restrict__ C, int ntot,
             double *__restrict__ B) {
    __shared__ double part[8][NZ_MAXK * (NZ_MAXK + 1) / 2];
    const int obj = blockIdx.x;
    const double *c = C + (size_t)obj * nk * ntot;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc[NZ_MAXK * (NZ_MAXK + 1) / 2];
#pragma unroll
    for (int q = 0; q < NZ_MAXK * (NZ_MAXK + 1) / 2; q++) acc[q] = 0.0;
    for (int x = threadIdx.x; x < ntot; x += blockDim.x) {
        double cv[NZ_MAXK];
#pragma unroll
        for (int a = 0; a < NZ_MAXK; a++)
            cv[a] = a < nk ? c[(size_t)a * ntot + x] : 0.0;
        int q = 0;
#pragma unroll
        for (int a = 0; a < NZ_MAXK; a++)
#pragma unroll
            for (int b = 0; b <= a; b++) {
                acc[q] = fma(cv[a], cv[b], acc[q]);
                q++;
            }
    }
#pragma unroll
    for (int q = 0; q < NZ_MAXK * (NZ_MAXK + 1) / 2; q++) {
        double s = warp_sum(acc[q]);
        if (lane == 0) part[warp][q] = s;
    }
    __syncthreads();
    const int nwarps = blockDim.x >> 5;
    for (int q = threadIdx.x; q < NZ_MAXK * (NZ_MAXK + 1) / 2;
         q += blockDim.x) {
        // q -> (a, b) of the packed lower triangle
        int a = 0;
        while ((a + 1) * (a + 2) / 2 <= q) a++;
        int b = q - a * (a + 1) / 2;
        if (a >= nk) continue;
        double s = 0.0;
        for (int wi = 0; wi < nwarps; wi++) s += part[wi][q];
        s /= (double)ntot;
        B[((size_t)obj * nk + a) * nk + b] = s;
        B[((size_t)obj * nk + b) * nk + a] = s;
    }
}

static int noise_b_launch(int n, int nk, const NoiseMap &map, const double *D,
                          int nd, const double *scale, const double *weights,
                          const double *noise, int nrow, int ncol,
                          double *C, double *B, cudaStream_t st) {
    const int wp = ((ncol + NZ_JB - 1) / NZ_JB) * NZ_JB;
    size_t smem = sizeof(double)
        * ((size_t)nrow * (2 * wp + 1) + (size_t)nrow * wp);
    if (smem > (size_t)smem_optin_bytes()) {
        // staged planes in a stream-ordered global scratch, a bounded number
        // of stamps per launch
        const size_t per = (size_t)nk * (smem / sizeof(double));
        int chunk = (int)((size_t)(1u << 28) / per);       // 2 GiB of doubles
        if (chunk < 1) chunk = 1;
        if (chunk > n) chunk = n;
        double *scr = nullptr;
        NGX_CUDA(cudaMallocAsync((void **)&scr, sizeof(double) * per * chunk,
                                 st));
        const size_t ntot = (size_t)nrow * ncol;
        for (int i0 = 0; i0 < n; i0 += chunk) {
            const int m = n - i0 < chunk ? n - i0 : chunk;
            dim3 grid(m, nk);
            k_noise_corr<true><<<grid, 256, 0, st>>>(
                nk, map, D + (size_t)i0 * nd * ntot, nd,
                scale ? scale + (size_t)i0 * nk : nullptr,
                weights + i0 * ntot, noise + i0 * ntot, nrow, ncol, wp,
                C + (size_t)i0 * nk * ntot, scr);
        }
        cudaError_t e = cudaGetLastError();
        cudaFreeAsync(scr, st);
        NGX_CUDA(e);
    } else {
        NGX_CUDA(cudaFuncSetAttribute(
            k_noise_corr<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
            (int)smem));
        dim3 grid(n, nk);
        k_noise_corr<false><<<grid, 256, smem, st>>>(
            nk, map, D, nd, scale, weights, noise, nrow, ncol, wp, C, nullptr);
        NGX_LAUNCH_CHECK();
    }
    k_noise_gram<<<n, 256, 0, st>>>(nk, C, nrow * ncol, B);
    NGX_LAUNCH_CHECK();
    return NGX_OK;
}

}  // namespace ngx

using namespace ngx;

extern "C" {

int ngx_noise_b(int n, int nk, const int32_t *rows, const double *dimages,
                int nd, const double *scale, const double *weights,
                const double *noise, int nrow, int ncol, double *work,
                double *B, void *stream) {
    NGX_REQUIRE(n >= 0 && nk >= 1 && nk <= NZ_MAXK && nd >= nk);
    NGX_REQUIRE(dimages && weights && noise && work && B);
    NGX_REQUIRE(nrow >= 1 && ncol >= 1);
    if (n == 0) return NGX_OK;
    NoiseMap map;
    for (int a = 0; a < NZ_MAXK; a++) {
        map.rows[a] = a < nk ? (rows ? rows[a] : a) : 0;
        NGX_REQUIRE(map.rows[a] >= 0 && map.rows[a] < nd);
    }
    return noise_b_launch(n, nk, map, dimages, nd, scale, weights, noise, nrow,
                          ncol, work, B, as_stream(stream));
}

// generic host entry: derivative images supplied by the caller (the
// central-difference images of the models without analytic derivatives,
// noise_cov.py:203-226)
int ngx_noise_b_host(int n, int nk, const double *dimages,
                     const double *weights, const double *noise, int nrow,
                     int ncol, double *B) {
    NGX_REQUIRE(n >= 0 && nk >= 1 && nk <= NZ_MAXK);
    NGX_REQUIRE(dimages && weights && noise && B);
    if (n == 0) return NGX_OK;
    const size_t ntot = (size_t)nrow * ncol;
    const int chunk = 1024;
    for (int i0 = 0; i0 < n; i0 += chunk) {
        const int m = n - i0 < chunk ? n - i0 : chunk;
        DevBuf<double> dD((size_t)m * nk * ntot, dimages + (size_t)i0 * nk * ntot);
        DevBuf<double> dW(m * ntot, weights + i0 * ntot);
        DevBuf<double> dN(m * ntot, noise + i0 * ntot);
        DevBuf<double> dC((size_t)m * nk * ntot), dB((size_t)m * nk * nk);
        NGX_CUDA(dD.err); NGX_CUDA(dW.err); NGX_CUDA(dN.err);
        NGX_CUDA(dC.err); NGX_CUDA(dB.err);
        int rc = ngx_noise_b(m, nk, nullptr, dD.d, nk, nullptr, dW.d, dN.d,
                             nrow, ncol, dC.d, dB.d, 0);
        if (rc) return rc;
        NGX_CUDA(cudaDeviceSynchronize());
        NGX_CUDA(dB.to_host(B + (size_t)i0 * nk * nk));
    }
    return NGX_OK;
}

// analytic models (gauss, exp, dev): the derivative images come from
// deriv_images on the device (noise_cov.py:160-201); parameter order of B is
// (cen1, cen2, g1, g2, T, flux), the flux derivative being value / flux
int ngx_noise_cov_host(int n, int ngauss, const double *gpars,
                       const double *dcov, const double *flux,
                       const ngx_jacobian *jac, const double *weights,
                       const double *noise, int nrow, int ncol, double *B) {
    NGX_REQUIRE(n >= 0 && ngauss >= 1 && gpars && dcov && flux && jac);
    NGX_REQUIRE(weights && noise && B && nrow >= 1 && ncol >= 1);
    if (n == 0) return NGX_OK;
    const size_t ntot = (size_t)nrow * ncol;
    const int chunk = 1024;
    const int32_t rows[6] = {1, 2, 3, 4, 5, 0};
    double *hs = (double *)malloc(sizeof(double) * 6 * chunk);
    if (!hs) return fail(NGX_ERR_ARG, "ngx_noise_cov_host: out of memory%s%s");
    int rc = NGX_OK;
    for (int i0 = 0; i0 < n && rc == NGX_OK; i0 += chunk) {
        const int m = n - i0 < chunk ? n - i0 : chunk;
        for (int i = 0; i < m; i++) {
            for (int a = 0; a < 5; a++) hs[6 * i + a] = 1.0;
            hs[6 * i + 5] = 1.0 / flux[i0 + i];
        }
        DevBuf<double> dgp((size_t)m * ngauss * 6,
                           gpars + (size_t)i0 * ngauss * 6);
        DevBuf<double> ddc((size_t)m * ngauss * 9,
                           dcov + (size_t)i0 * ngauss * 9);
        DevBuf<ngx_jacobian> dj(m, jac + i0);
        DevBuf<double> dS((size_t)m * 6, hs);
        DevBuf<double> dW(m * ntot, weights + i0 * ntot);
        DevBuf<double> dN(m * ntot, noise + i0 * ntot);
        DevBuf<double> dD((size_t)m * 6 * ntot), dC((size_t)m * 6 * ntot);
        DevBuf<double> dB((size_t)m * 36);
        cudaError_t e = dgp.err;
        if (e == cudaSuccess) e = ddc.err;
        if (e == cudaSuccess) e = dj.err;
        if (e == cudaSuccess) e = dS.err;
        if (e == cudaSuccess) e = dW.err;
        if (e == cudaSuccess) e = dN.err;
        if (e == cudaSuccess) e = dD.err;
        if (e == cudaSuccess) e = dC.err;
        if (e == cudaSuccess) e = dB.err;
        if (e == cudaSuccess) e = dD.zero();
        if (e != cudaSuccess) {
            rc = fail(NGX_ERR_CUDA, "ngx_noise_cov_host: %s%s",
                      cudaGetErrorString(e));
            break;
        }
        rc = ngx_deriv_images(m, ngauss, dgp.d, ddc.d, dj.d, nrow, ncol, dD.d,
                              0);
        if (rc == NGX_OK)
            rc = ngx_noise_b(m, 6, rows, dD.d, 6, dS.d, dW.d, dN.d, nrow, ncol,
                             dC.d, dB.d, 0);
        if (rc == NGX_OK && cudaDeviceSynchronize() != cudaSuccess)
            rc = fail(NGX_ERR_CUDA, "ngx_noise_cov_host: %s%s",
                      cudaGetErrorString(cudaGetLastError()));
        if (rc == NGX_OK && dB.to_host(B + (size_t)i0 * 36) != cudaSuccess)
            rc = fail(NGX_ERR_CUDA, "ngx_noise_cov_host: copy back failed%s%s");
    }
    free(hs);
    return rc;
}

}  // extern "C"
