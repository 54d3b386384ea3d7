// Mixture bookkeeping and the stand-alone pixel passes:
//   ngx_gmix_fill / _fill_cm / _set_norms / _convolve   (gmix_nb.py:176-649)
//   ngx_make_pixels / ngx_make_coords                   (pixels_nb.py:6-94)
//   ngx_render                                          (render_nb.py:9-36)
//   ngx_loglike / ngx_loglike_pixels   (gmix_nb.py:824-937, one fused pass)
//   ngx_deriv_images                   (fitting/derivs_nb.py:40-127)
// plus library management and the *_host staging variants.
#include <cuda_pipeline.h>
#include "ngx_core.cuh"
#include "ngx_host.cuh"

namespace ngx {

char *err_buf() {
    static thread_local char buf[512] = "";
    return buf;
}

static int g_sm_count = 0;
static int g_smem_optin = 0;
static int g_device = -1;

static int refresh_device_info() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (dev == g_device && g_sm_count > 0) return 0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return -1;
    g_sm_count = prop.multiProcessorCount;
    g_smem_optin = (int)prop.sharedMemPerBlockOptin;
    g_device = dev;
    return 0;
}

int sm_count() {
    refresh_device_info();
    return g_sm_count > 0 ? g_sm_count : 148;
}
int smem_optin_bytes() {
    refresh_device_info();
    return g_smem_optin > 0 ? g_smem_optin : 227 * 1024;
}

// ---------------------------------------------------------------------
// mixtures: one thread per object (negligible work, launch-latency bound)
__global__ void k_gmix_fill(int n, int model, int ngauss,
                            const double *__restrict__ pars, int npars,
                            ngx_gauss *__restrict__ gmix,
                            int32_t *__restrict__ status) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *p = pars + (size_t)i * npars;
    ngx_gauss *gm = gmix + (size_t)i * ngauss;
    int st = NGX_ST_OK;
    if (model == NGX_FULL) {                    // gmix_fill_full :408-427
        for (int k = 0; k < ngauss; k++)
            gauss_set(gm[k], p[6 * k], p[6 * k + 1], p[6 * k + 2],
                      p[6 * k + 3], p[6 * k + 4], p[6 * k + 5]);
    } else if (model == NGX_COELLIP) {          // gmix_fill_coellip :375-405
        double e1, e2;
        st = g1g2_to_e1e2(p[2], p[3], &e1, &e2);
        if (st == NGX_ST_OK)
            for (int k = 0; k < ngauss; k++) {
                double Thalf = 0.5 * p[4 + k];
                gauss_set(gm[k], p[4 + ngauss + k], p[0], p[1],
                          Thalf * (1 - e1), Thalf * e2, Thalf * (1 + e1));
            }
    } else {
        ModelHead h = model_head(model, p);
        st = h.status;
        if (st == NGX_ST_OK)
            for (int k = 0; k < ngauss; k++) {
                double pk, irr, irc, icc;
                model_component(model, h, k, &pk, &irr, &irc, &icc);
                gauss_set(gm[k], pk, h.row, h.col, irr, irc, icc);
            }
    }
    if (status) status[i] = st;
}

// gmix_fill_cm :430-466; cm = fracdev, TdByTe, Tfactor
__global__ void k_gmix_fill_cm(int n, const double *__restrict__ cm,
                               const double *__restrict__ pars,
                               ngx_gauss *__restrict__ gmix,
                               int32_t *__restrict__ status) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *p = pars + (size_t)i * 6;
    ModelHead h;
    h.row = p[0]; h.col = p[1];
    h.fracdev = cm[3 * i]; h.TdByTe = cm[3 * i + 1];
    h.T = p[4] * cm[3 * i + 2];
    h.flux = p[5];
    h.status = g1g2_to_e1e2(p[2], p[3], &h.e1, &h.e2);
    if (h.status == NGX_ST_OK)
        for (int k = 0; k < 16; k++) {
            double pk, irr, irc, icc;
            model_component(NGX_CM, h, k, &pk, &irr, &irc, &icc);
            gauss_set(gmix[(size_t)i * 16 + k], pk, h.row, h.col, irr, irc,
                      icc);
        }
    if (status) status[i] = h.status;
}

__global__ void k_set_norms(int n, int ngauss, ngx_gauss *__restrict__ gmix,
                            int32_t *__restrict__ status) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int st = NGX_ST_OK;
    // the reference raises at the first bad gaussian, leaving the earlier
    // ones normalised (gmix_nb.py:186-187)
    for (int k = 0; k < ngauss && st == NGX_ST_OK; k++)
        st = gauss_set_norm(gmix[(size_t)i * ngauss + k]);
    if (status) status[i] = st;
}

// gmix_convolve_fill :609-649 (object-major order)
__global__ void k_convolve(int n, int nobj, int npsf,
                           const ngx_gauss *__restrict__ obj,
                           const ngx_gauss *__restrict__ psf,
                           ngx_gauss *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const ngx_gauss *o = obj + (size_t)i * nobj;
    const ngx_gauss *q = psf + (size_t)i * npsf;
    double prow = 0.0, pcol = 0.0, psum = 0.0;      // gmix_get_cen :108-130
    for (int k = 0; k < npsf; k++) {
        double p = q[k].p;
        prow += p * q[k].row;
        pcol += p * q[k].col;
        psum += p;
    }
    prow /= psum;
    pcol /= psum;
    double ipsum = 1.0 / psum;
    int itot = 0;
    for (int a = 0; a < nobj; a++)
        for (int b = 0; b < npsf; b++) {
            gauss_set(out[(size_t)i * nobj * npsf + itot],
                      o[a].p * q[b].p * ipsum,
                      o[a].row + (q[b].row - prow),
                      o[a].col + (q[b].col - pcol),
                      o[a].irr + q[b].irr, o[a].irc + q[b].irc,
                      o[a].icc + q[b].icc);
            itot++;
        }
}

// ---------------------------------------------------------------------
// pixel coordinates.  v,u are formed exactly as jacobian_get_vu does
// (jacobian_nb.py:4-16): two products and one sum, no fused multiply-add.
__device__ __forceinline__ void pixel_vu(const ngx_jacobian &j, int row,
                                         int col, double *v, double *u) {
    double rowdiff = row - j.row0;
    double coldiff = col - j.col0;
    *v = j.dvdrow * rowdiff + j.dvdcol * coldiff;
    *u = j.dudrow * rowdiff + j.dudcol * coldiff;
}

// fill_pixels :6-58 with ordered compaction of positive-weight pixels.
// One CTA per object; a ballot/popc scan keeps row-major order.
__global__ void k_make_pixels(int nrow, int ncol,
                              const double *__restrict__ images,
                              const double *__restrict__ weights,
                              const ngx_jacobian *__restrict__ jac,
                              int ignore_zero_weight,
                              ngx_pixel *__restrict__ pixels,
                              int32_t *__restrict__ npix_out) {
    __shared__ int warp_counts[32];
    __shared__ int running;
    int obj = blockIdx.x;
    int ntot = nrow * ncol;
    const double *im = images + (size_t)obj * ntot;
    const double *wt = weights + (size_t)obj * ntot;
    ngx_pixel *out = pixels + (size_t)obj * ntot;
    ngx_jacobian j = jac[obj];
    double area = j.scale * j.scale;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int nwarps = blockDim.x >> 5;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    for (int base = 0; base < ntot; base += blockDim.x) {
        int p = base + threadIdx.x;
        double ivar = 0.0;
        bool valid = false;
        if (p < ntot) {
            ivar = wt[p];
            valid = !(ignore_zero_weight && ivar <= 0.0);
        }
        unsigned bal = __ballot_sync(0xffffffffu, valid);
        if (lane == 0) warp_counts[warp] = __popc(bal);
        __syncthreads();
        int off = running;
        for (int w = 0; w < warp; w++) off += warp_counts[w];
        int pos = off + __popc(bal & ((1u << lane) - 1u));
        if (valid) {
            int row = p / ncol, col = p - row * ncol;
            ngx_pixel px;
            pixel_vu(j, row, col, &px.v, &px.u);
            px.area = area;
            px.val = im[p];
            if (ivar < 0.0) ivar = 0.0;
            px.ierr = sqrt(ivar);
            px.fdiff = 0.0;
            out[pos] = px;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < nwarps; w++) tot += warp_counts[w];
            running += tot;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) npix_out[obj] = running;
}

// fill_coords :61-94
__global__ void k_make_coords(int nrow, int ncol,
                              const ngx_jacobian *__restrict__ jac,
                              ngx_coord *__restrict__ coords) {
    int obj = blockIdx.x;
    int ntot = nrow * ncol;
    int p = blockIdx.y * blockDim.x + threadIdx.x;
    if (p >= ntot) return;
    ngx_jacobian j = jac[obj];
    int row = p / ncol, col = p - row * ncol;
    ngx_coord c;
    pixel_vu(j, row, col, &c.v, &c.u);
    c.area = j.scale * j.scale;
    coords[(size_t)obj * ntot + p] = c;
}

// ---------------------------------------------------------------------
// render: grid = (pixel tiles, objects).  Gaussians are staged in shared
// memory once per CTA; every thread owns one pixel.
__global__ void k_render(int ngauss, const ngx_gauss *__restrict__ gmix,
                         const ngx_jacobian *__restrict__ jac, int nrow,
                         int ncol, int fast_exp, double *__restrict__ images,
                         int32_t *__restrict__ status) {
    extern __shared__ double smem_d[];
    double *tab = smem_d;                       // 16
    GaussV *gv = (GaussV *)(smem_d + 16);       // ngauss
    ngx_gauss *gr = (ngx_gauss *)(gv + ngauss); // ngauss (exact path)
    __shared__ int bad;
    int obj = blockIdx.x;
    int ntot = nrow * ncol;
    ngx_jacobian j = jac[obj];
    double area = j.scale * j.scale;
    if (threadIdx.x == 0) bad = 0;
    load_exp_table(tab);
    __syncthreads();
    for (int k = threadIdx.x; k < ngauss; k += blockDim.x) {
        ngx_gauss g = gmix[(size_t)obj * ngauss + k];
        // render_nb.py:27-28 sets the norms when norm_set == 0
        if (gauss_set_norm(g) != NGX_ST_OK) bad = 1;
        gv[k] = make_gaussv(g, area);
        gr[k] = g;
    }
    __syncthreads();
    if (bad) {
        if (threadIdx.x == 0 && blockIdx.y == 0 && status)
            status[obj] = NGX_ST_RANGE_ERROR;
        return;
    }
    if (threadIdx.x == 0 && blockIdx.y == 0 && status)
        status[obj] = NGX_ST_OK;
    int p = blockIdx.y * blockDim.x + threadIdx.x;
    if (p >= ntot) return;
    int row = p / ncol, col = p - row * ncol;
    double v, u;
    pixel_vu(j, row, col, &v, &u);
    double s = 0.0;
    if (fast_exp) {
        for (int k = 0; k < ngauss; k++) s += eval_fast(gv[k], v, u, tab);
    } else {
        for (int k = 0; k < ngauss; k++) s += eval_exact(gr[k], v, u, area);
    }
    images[(size_t)obj * ntot + p] += s;
}

// ---------------------------------------------------------------------
// The fused fdiff / loglike / s2n pass.  One CTA per object.  Each thread
// keeps C pixels in registers and walks the gaussians in the outer loop, so
// the per-gaussian constants are read from shared memory once per C
// evaluations.
template <int C>
__device__ __forceinline__ void value_chunk(const GaussV *gv, int ngauss,
                                            const double *tab,
                                            const double (&v)[C],
                                            const double (&u)[C],
                                            double (&acc)[C]) {
    for (int k = 0; k < ngauss; k++) {
        GaussV g = gv[k];
#pragma unroll
        for (int c = 0; c < C; c++) acc[c] += eval_fast(g, v[c], u[c], tab);
    }
}

#define LL_C 4

// Pre-built pixel records (GMix.fill_fdiff / get_loglike on obs.pixels):
// every record takes part, fdiff[i] belongs to pixel i.
__global__ void k_loglike_pixels(int ngauss,
                                 const ngx_gauss *__restrict__ gmix,
                                 const ngx_pixel *__restrict__ pixels,
                                 int maxpix,
                                 const int32_t *__restrict__ npix_in,
                                 double *__restrict__ out4,
                                 double *__restrict__ fdiff,
                                 int32_t *__restrict__ status) {
    extern __shared__ double smem_d[];
    double *tab = smem_d;                       // 16
    double *scratch = smem_d + 16;              // 32 * nwarps + 32
    GaussV *gv = (GaussV *)(scratch + 32 * 9);  // ngauss
    __shared__ int bad;
    const int obj = blockIdx.x;
    const int ntot = npix_in[obj];
    const ngx_pixel *px0 = pixels + (size_t)obj * maxpix;
    const double area = ntot > 0 ? px0[0].area : 0.0;
    if (threadIdx.x == 0) bad = 0;
    load_exp_table(tab);
    __syncthreads();
    for (int k = threadIdx.x; k < ngauss; k += blockDim.x) {
        ngx_gauss g = gmix[(size_t)obj * ngauss + k];
        if (gauss_set_norm(g) != NGX_ST_OK) bad = 1;
        gv[k] = make_gaussv(g, area);
    }
    __syncthreads();
    if (bad) {
        if (threadIdx.x == 0) {
            if (status) status[obj] = NGX_ST_RANGE_ERROR;
            for (int q = 0; q < 4; q++) out4[4 * (size_t)obj + q] = nan("");
        }
        return;
    }
    double sums[4] = {0.0, 0.0, 0.0, 0.0};   // chi2, s2n_numer, s2n_denom, n
    for (int base = 0; base < ntot; base += blockDim.x * LL_C) {
        double v[LL_C], u[LL_C], acc[LL_C], val[LL_C], ierr[LL_C];
#pragma unroll
        for (int c = 0; c < LL_C; c++) {
            int p = base + c * blockDim.x + threadIdx.x;
            acc[c] = 0.0;
            // lanes past the end sit at the first pixel and are dropped below
            ngx_pixel px = px0[p < ntot ? p : 0];
            v[c] = px.v; u[c] = px.u; val[c] = px.val; ierr[c] = px.ierr;
        }
        value_chunk<LL_C>(gv, ngauss, tab, v, u, acc);
#pragma unroll
        for (int c = 0; c < LL_C; c++) {
            int p = base + c * blockDim.x + threadIdx.x;
            if (p < ntot) {
                double ivar = ierr[c] * ierr[c];
                double diff = acc[c] - val[c];
                sums[0] += diff * diff * ivar;
                sums[1] += val[c] * acc[c] * ivar;
                sums[2] += acc[c] * acc[c] * ivar;
                sums[3] += 1.0;
                if (fdiff) fdiff[(size_t)obj * maxpix + p] = diff * ierr[c];
            }
        }
    }
    block_sum_to<4>(sums, scratch + 32, scratch);
    if (threadIdx.x == 0) {
        out4[4 * (size_t)obj + 0] = -0.5 * scratch[0];
        out4[4 * (size_t)obj + 1] = scratch[1];
        out4[4 * (size_t)obj + 2] = scratch[2];
        out4[4 * (size_t)obj + 3] = scratch[3];
        if (status) status[obj] = NGX_ST_OK;
    }
}

// ---------------------------------------------------------------------
// The same fused pass for image/weight/jacobian input, organised like the
// value pass of the LM kernel.  Pixels are taken in TILED order (4 rows x 8
// columns per warp tile, C consecutive tiles per warp and chunk).  Before the
// pixel loop every tile gets a bit mask of the gaussians whose chi^2 < 25
// bounding box meets it, so a warp only visits gaussians that can contribute;
// their constants and rotated pn*exp(i) tables sit in shared memory.
// fdiff keeps the reference's order (row-major over the positive-weight
// pixels): a ballot pre-pass counts the valid pixels of every 8-pixel row
// segment, prefix sums give each segment's offset.
#define LLT_C 3
#define LLT_REC 24      // 16-entry table, then row, col, a, b, c, -, rv, ru
__global__ void __launch_bounds__(256, 3)
k_loglike_tiled(int ngauss, const ngx_gauss *__restrict__ gmix,
                const double *__restrict__ images,
                const double *__restrict__ weights,
                const ngx_jacobian *__restrict__ jac, int nrow, int ncol,
                int tpr, int nq, int ignore_zero_weight,
                double *__restrict__ out4, double *__restrict__ fdiff,
                int32_t *__restrict__ status) {
    extern __shared__ double smem_d[];
    const int ntile = nq >> 5;
    const int nwords = (ngauss + 31) >> 5;
    double *scratch = smem_d;                   // 32 + 32 * 8 (warps)
    double *mix = scratch + 32 * 9;             // ngauss * LLT_REC
    unsigned *tmask = (unsigned *)(mix + LLT_REC * ngauss);  // ntile * nwords
    int *segoff = (int *)(tmask + (size_t)ntile * nwords);  // nrow * tpr
    int *rowoff = segoff + nrow * tpr;          // nrow
    // the stamp itself, tiled order (8-byte aligned after the int tables)
    double *val_s = (double *)(((uintptr_t)(rowoff + nrow) + 7) & ~(uintptr_t)7);
    double *wt_s = val_s + nq;
    const int T = blockDim.x;
    const int obj = blockIdx.x;
    const int ntot = nrow * ncol;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const ngx_jacobian j = jac[obj];
    const double area = j.scale * j.scale;
    const double *im = images + (size_t)obj * ntot;
    const double *wt = weights + (size_t)obj * ntot;
    // start the stamp's trip from HBM now (asynchronous 8-byte copies into
    // the tiled slots); it is first needed after the mixture set-up
    for (int q = threadIdx.x; q < nq; q += T) {
        int row, col;
        q_to_rowcol(q, tpr, &row, &col);
        if (row < nrow && col < ncol) {
            __pipeline_memcpy_async(&val_s[q], &im[row * ncol + col],
                                    sizeof(double));
            __pipeline_memcpy_async(&wt_s[q], &wt[row * ncol + col],
                                    sizeof(double));
        } else {
            wt_s[q] = 1.0;          // padding slot: never a masked pixel
        }
    }
    __pipeline_commit();
    // several threads share a tile's gaussians in the mask pass below and
    // merge their partial masks with atomicOr
    const int nsplit = max(1, T / ntile);
    if (nsplit > 1)
        for (int i = threadIdx.x; i < ntile * nwords; i += T) tmask[i] = 0u;
    int inv = 0;
    for (int k = threadIdx.x; k < ngauss; k += T) {
        ngx_gauss g = gmix[(size_t)obj * ngauss + k];
        if (gauss_set_norm(g) != NGX_ST_OK) inv = 2;
        double *c = mix + LLT_REC * k + 16;
        c[0] = g.row; c[1] = g.col;
        c[2] = -0.5 * g.dcc; c[3] = -0.5 * g.drr; c[4] = g.drc;
        c[5] = 0.0;
        c[6] = 5.0 * sqrt(g.irr); c[7] = 5.0 * sqrt(g.icc);
        fill_rot_table(mix + LLT_REC * k, g.pnorm * area);
    }
    __pipeline_wait_prior(0);
    // fdiff positions need the compaction tables only when the stamp has
    // masked pixels; every thread looks at the slots it copied itself
    if (fdiff && ignore_zero_weight)
        for (int q = threadIdx.x; q < nq; q += T) inv |= wt_s[q] <= 0.0;
    const int both = __syncthreads_or(inv);
    const int compact = both & 1, bad = both & 2;
    if (compact) {
        // valid pixels per 8-pixel row segment
        for (int base = 0; base < nq; base += T) {
            int q = base + threadIdx.x;
            int row, col;
            q_to_rowcol(q, tpr, &row, &col);
            bool valid = false;
            if (q < nq && row < nrow && col < ncol)
                valid = !(ignore_zero_weight && wt_s[q] <= 0.0);
            unsigned bal = __ballot_sync(0xffffffffu, valid);
            if ((lane & 7) == 0 && q < nq && row < nrow)
                segoff[row * tpr + (col >> 3)] =
                    __popc((bal >> (lane & 24)) & 0xffu);
        }
        __syncthreads();
    }
    if (bad) {
        if (threadIdx.x == 0) {
            if (status) status[obj] = NGX_ST_RANGE_ERROR;
            for (int q = 0; q < 4; q++) out4[4 * (size_t)obj + q] = nan("");
        }
        return;
    }
    // gaussians that can reach each tile
    const int part = threadIdx.x / ntile;      // 0 when ntile >= T
    if (part < nsplit)
        for (int t = threadIdx.x - part * ntile; t < ntile; t += T) {
            double vc, uc, hv, hu;
            tile_extent(j, t, tpr, nrow, ncol, &vc, &uc, &hv, &hu);
            for (int w = 0; w < nwords; w++) {
                unsigned m = 0;
                int kend = min(32, ngauss - 32 * w);
                for (int kk = part; kk < kend; kk += nsplit) {
                    const double *g = mix + LLT_REC * (32 * w + kk) + 16;
                    if (tile_hit(vc, uc, hv, hu, g[0], g[1], g[6], g[7]))
                        m |= 1u << kk;
                }
                if (nsplit == 1) tmask[(size_t)t * nwords + w] = m;
                else if (m) atomicOr(&tmask[(size_t)t * nwords + w], m);
            }
        }
    if (compact) {
        for (int r = threadIdx.x; r < nrow; r += T) {
            int run = 0;
            for (int t = 0; t < tpr; t++) {
                int c = segoff[r * tpr + t];
                segoff[r * tpr + t] = run;
                run += c;
            }
            rowoff[r] = run;
        }
        __syncthreads();
        if (warp == 0) {
            // exclusive scan of the row counts, 32 rows per round
            int carry = 0;
            for (int r0 = 0; r0 < nrow; r0 += 32) {
                const int r = r0 + lane;
                const int c = r < nrow ? rowoff[r] : 0;
                int s2 = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    int t = __shfl_up_sync(0xffffffffu, s2, d);
                    if (lane >= d) s2 += t;
                }
                if (r < nrow) rowoff[r] = carry + s2 - c;
                carry += __shfl_sync(0xffffffffu, s2, 31);
            }
        }
    }
    __syncthreads();
    double sums[4] = {0.0, 0.0, 0.0, 0.0};
    for (int base = 0; base < nq; base += T * LLT_C) {
        // this warp's LLT_C consecutive tiles
        const int q0 = base + (warp * LLT_C) * 32 + lane;
        double v[LLT_C], u[LLT_C], acc[LLT_C];
#pragma unroll
        for (int c = 0; c < LLT_C; c++) {
            int row, col;
            q_to_rowcol(q0 + 32 * c, tpr, &row, &col);
            acc[c] = 0.0;
            // padding lanes evaluate a real pixel and are dropped below
            pixel_vu(j, min(row, nrow - 1), min(col, ncol - 1), &v[c], &u[c]);
        }
        for (int w = 0; w < nwords; w++) {
            unsigned m = 0;
#pragma unroll
            for (int c = 0; c < LLT_C; c++) {
                int t = (q0 >> 5) + c;
                if (t < ntile) m |= tmask[(size_t)t * nwords + w];
            }
#pragma unroll 1
            while (m) {
                const int k = 32 * w + __ffs((int)m) - 1;
                m &= m - 1;
                const double *tab = mix + LLT_REC * k;
                const double2 *g2 = (const double2 *)(tab + 16);
                const double2 rc = g2[0], ab = g2[1];
                const double grow = rc.x, gcol = rc.y, ga = ab.x, gb = ab.y,
                             gcx = tab[20];
                double x[LLT_C], e[LLT_C];
                bool ap[LLT_C], anyap = false;
#pragma unroll
                for (int c = 0; c < LLT_C; c++) {
                    double dv = v[c] - grow;
                    double du = u[c] - gcol;
                    double t = fma(gcx, du, ga * dv);
                    x[c] = fma(gb * du, du, dv * t);
                    unsigned slot;
                    bool inr;
                    double f = fexp_split(x[c], &slot, &inr, &ap[c]);
                    e[c] = tab[slot] * fexp_poly_c(f);
                    anyap |= ap[c];
                }
                if (__any_sync(0xffffffffu, anyap)) {
#pragma unroll
                    // some lane of some pixel group sits in the
                    // apodization band: the window of every group, without
                    // a vote per group (7 fp64 instructions cost less than
                    // the vote's branch bookkeeping; measured 0.663 -> 0.628
                    // ms per 10 000 stamps)
                    for (int c = 0; c < LLT_C; c++) {
                        double w = apod_window_fast(x[c]);
                        if (ap[c]) e[c] *= w;
                    }
                }
#pragma unroll
                for (int c = 0; c < LLT_C; c++) acc[c] += e[c];
            }
        }
#pragma unroll
        for (int c = 0; c < LLT_C; c++) {
            int q = q0 + 32 * c;
            int row, col;
            q_to_rowcol(q, tpr, &row, &col);
            bool valid = false;
            double d = 0.0, ie = 0.0;
            if (q < nq && row < nrow && col < ncol) {
                double w = wt_s[q];
                valid = !(ignore_zero_weight && w <= 0.0);
                if (valid) {
                    d = val_s[q];
                    ie = sqrt(w < 0.0 ? 0.0 : w);
                }
            }
            unsigned bal = __ballot_sync(0xffffffffu, valid);
            if (valid) {
                double ivar = ie * ie;
                double diff = acc[c] - d;
                sums[0] += diff * diff * ivar;
                sums[1] += d * acc[c] * ivar;
                sums[2] += acc[c] * acc[c] * ivar;
                sums[3] += 1.0;
                if (fdiff) {
                    int pos = row * ncol + col;
                    if (compact) {
                        unsigned seg = (bal >> (lane & 24)) & 0xffu;
                        pos = rowoff[row] + segoff[row * tpr + (col >> 3)]
                            + __popc(seg & ((1u << (lane & 7)) - 1u));
                    }
                    fdiff[(size_t)obj * ntot + pos] = diff * ie;
                }
            }
        }
    }
    block_sum_to<4>(sums, scratch + 32, scratch);
    if (threadIdx.x == 0) {
        out4[4 * (size_t)obj + 0] = -0.5 * scratch[0];
        out4[4 * (size_t)obj + 1] = scratch[1];
        out4[4 * (size_t)obj + 2] = scratch[2];
        out4[4 * (size_t)obj + 3] = scratch[3];
        if (status) status[obj] = NGX_ST_OK;
    }
}

// ---------------------------------------------------------------------
// deriv_images, derivs_nb.py:40-127, in the reference's own operation
// order (this entry point exists for parity of the derivative math; the LM
// kernel uses the restructured form in ngx_lm.cu).
__global__ void k_deriv_images(int ngauss, const double *__restrict__ gpars,
                               const double *__restrict__ dcov,
                               const ngx_jacobian *__restrict__ jac, int nrow,
                               int ncol, double *__restrict__ out) {
    extern __shared__ double smem_d[];
    double *tab = smem_d;             // 16
    double *gc = smem_d + 16;         // ngauss * 16
    int obj = blockIdx.x;
    int ntot = nrow * ncol;
    ngx_jacobian j = jac[obj];
    double area = j.scale * j.scale;
    load_exp_table(tab);
    for (int k = threadIdx.x; k < ngauss; k += blockDim.x) {
        const double *gp = gpars + ((size_t)obj * ngauss + k) * 6;
        const double *dc = dcov + ((size_t)obj * ngauss + k) * 9;
        double p = gp[0], irr = gp[3], irc = gp[4], icc = gp[5];
        double det = irr * icc - irc * irc;
        double *o = gc + 16 * k;
        o[0] = gp[1]; o[1] = gp[2];
        o[15] = det;
        if (det > 0.0) {
            o[2] = p / (NGX_TWO_PI * sqrt(det));           // norm
            double w11 = icc / det, w12 = -irc / det, w22 = irr / det;
            o[3] = w11; o[4] = w12; o[5] = w22;
            for (int a = 0; a < 3; a++) {
                o[6 + a] = w11 * dc[3 * a] + 2.0 * w12 * dc[3 * a + 1]
                    + w22 * dc[3 * a + 2];
            }
        }
    }
    __syncthreads();
    int p = blockIdx.y * blockDim.x + threadIdx.x;
    if (p >= ntot) return;
    int row = p / ncol, col = p - row * ncol;
    double v, u;
    pixel_vu(j, row, col, &v, &u);
    double o6[6] = {0, 0, 0, 0, 0, 0};
    for (int k = 0; k < ngauss; k++) {
        const double *o = gc + 16 * k;
        if (!(o[15] > 0.0)) continue;
        const double *dc = dcov + ((size_t)obj * ngauss + k) * 9;
        double dv = v - o[0];
        double du = u - o[1];
        double w11 = o[3], w12 = o[4], w22 = o[5];
        double qv = w11 * dv + w12 * du;
        double qu = w12 * dv + w22 * du;
        double chi2 = dv * qv + du * qu;
        if (chi2 >= 25.0 || chi2 < 0.0) continue;
        double x = -0.5 * chi2;
        double val = o[2] * fexp_ref(x) * area;
        double valc;
        if (chi2 > 20.0) {
            double w = apod_window_x(x);
            valc = val * (w - 2.0 * apod_window_deriv_x(x));
            val *= w;
        } else {
            valc = val;
        }
        o6[0] += val;
        o6[1] += valc * qv;
        o6[2] += valc * qu;
        for (int a = 0; a < 3; a++) {
            double quad = qv * qv * dc[3 * a] + 2.0 * qv * qu * dc[3 * a + 1]
                + qu * qu * dc[3 * a + 2];
            o6[3 + a] += 0.5 * (valc * quad - val * o[6 + a]);
        }
    }
    for (int a = 0; a < 6; a++)
        out[((size_t)obj * 6 + a) * ntot + p] += o6[a];
}

// ---------------------------------------------------------------------
// FP64 pipe peak: 8 independent dependent-FMA chains per thread
__global__ void k_peak_dfma(int iters, double *sink) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    double m = 1.0000001, c = 1.0e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c);
        a3 = fma(a3, m, c); a4 = fma(a4, m, c); a5 = fma(a5, m, c);
        a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) sink[0] = s;
}

// FP64 tensor-core probe: mode 0 = 4 independent m8n8k4 DMMA chains per warp,
// mode 1 = the same interleaved with 8 DFMA chains (do they share a pipe?)
__device__ __forceinline__ void dmma8x8x4(double &c0, double &c1, double a,
                                          double b) {
    asm volatile(
        "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, "
        "{%0,%1};\n"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}
__global__ void k_peak_dmma(int iters, int mode, double *sink) {
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0e-3;
    double c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double f0 = a, f1 = a + 1, f2 = a + 2, f3 = a + 3, f4 = a + 4, f5 = a + 5,
           f6 = a + 6, f7 = a + 7;
    const double m = 1.0000001, k = 1.0e-9;
    for (int i = 0; i < iters; i++) {
        dmma8x8x4(c[0], c[1], a, b);
        dmma8x8x4(c[2], c[3], a, b);
        dmma8x8x4(c[4], c[5], a, b);
        dmma8x8x4(c[6], c[7], a, b);
        if (mode == 1) {
            f0 = fma(f0, m, k); f1 = fma(f1, m, k); f2 = fma(f2, m, k);
            f3 = fma(f3, m, k); f4 = fma(f4, m, k); f5 = fma(f5, m, k);
            f6 = fma(f6, m, k); f7 = fma(f7, m, k);
        }
    }
    double s = c[0] + c[1] + c[2] + c[3] + c[4] + c[5] + c[6] + c[7]
        + f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7;
    if (s == 123.456) sink[0] = s;
}

}  // namespace ngx

using namespace ngx;

// =====================================================================
// C ABI
extern "C" {

int ngx_version(void) { return NGX_VERSION; }

const char *ngx_last_error(void) { return err_buf(); }

int ngx_init(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(NGX_ERR_NODEVICE,
                    "no CUDA device (%s); this engine has no CPU fallback%s",
                    cudaGetErrorString(e));
    NGX_REQUIRE(device >= 0 && device < count);
    NGX_CUDA(cudaSetDevice(device));
    NGX_CUDA(cudaFree(0));
    if (refresh_device_info() != 0)
        return fail(NGX_ERR_CUDA, "cannot query device properties%s%s");
    return NGX_OK;
}

int ngx_device_info(int *sms, int *smem_optin, char *name, int name_len) {
    int dev = 0;
    NGX_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    NGX_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sms) *sms = prop.multiProcessorCount;
    if (smem_optin) *smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (name && name_len > 0) {
        strncpy(name, prop.name, name_len - 1);
        name[name_len - 1] = 0;
    }
    return NGX_OK;
}

// page-locked host memory for the host-buffer entry points: their chunked
// H2D copies overlap the fits only from pinned memory
int ngx_host_alloc(size_t bytes, void **out) {
    NGX_REQUIRE(out);
    *out = nullptr;
    if (bytes == 0) return NGX_OK;
    NGX_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
    return NGX_OK;
}
int ngx_host_free(void *p) {
    if (p) NGX_CUDA(cudaFreeHost(p));
    return NGX_OK;
}
int ngx_host_register(void *p, size_t bytes) {
    NGX_REQUIRE(p || bytes == 0);
    if (bytes == 0) return NGX_OK;
    NGX_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return NGX_OK;
}
int ngx_host_unregister(void *p) {
    if (p) NGX_CUDA(cudaHostUnregister(p));
    return NGX_OK;
}

int ngx_gmix_fill(int n, int model, int ngauss, const double *pars, int npars,
                  ngx_gauss *gmix, int32_t *status, void *stream) {
    NGX_REQUIRE(n >= 0 && pars && gmix);
    if (n == 0) return NGX_OK;
    if (model == NGX_FULL) {
        NGX_REQUIRE(ngauss >= 1 && npars == 6 * ngauss);
    } else if (model == NGX_COELLIP) {
        NGX_REQUIRE(ngauss >= 1 && npars == 4 + 2 * ngauss);
    } else {
        NGX_REQUIRE(model != NGX_CM);
        NGX_REQUIRE(model_ngauss(model) == ngauss);
        NGX_REQUIRE(model_npars(model) == npars);
    }
    k_gmix_fill<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(
        n, model, ngauss, pars, npars, gmix, status);
    NGX_LAUNCH_CHECK();
    return NGX_OK;
}

int ngx_gmix_fill_cm(int n, const double *cm, const double *pars,
                     ngx_gauss *gmix, int32_t *status, void *stream) {
    NGX_REQUIRE(n >= 0 && cm && pars && gmix);
    if (n == 0) return NGX_OK;
    k_gmix_fill_cm<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(
        n, cm, pars, gmix, status);
    NGX_LAUNCH_CHECK();
    return NGX_OK;
}

int ngx_gmix_set_norms(int n, int ngauss, ngx_gauss *gmix, int32_t *status,
                       void *stream) {
    NGX_REQUIRE(n >= 0 && ngauss >= 1 && gmix);
    if (n == 0) return NGX_OK;
    k_set_norms<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(n, ngauss,
                                                                gmix, status);
    NGX_LAUNCH_CHECK();
    return NGX_OK;
}

int ngx_gmix_convolve(int n, int nobj, int npsf, const ngx_gauss *obj,
                      const ngx_gauss *psf, ngx_gauss *out, void *stream) {
    NGX_REQUIRE(n >= 0 && nobj >= 1 && npsf >= 1 && obj && psf && out);
    if (n == 0) return NGX_OK;
    k_convolve<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(n, nobj, npsf,
                                                               obj, psf, out);
    NGX_LAUNCH_CHECK();
    return NGX_OK;
}

int ngx_make_pixels(int n, int nrow, int ncol, const double *images,
                    const double *weights, const ngx_jacobian *jac,
                    int ignore_zero_weight, ngx_pixel *pixels, int32_t *npix,
                    void *stream) {
    NGX_REQUIRE(n >= 0 && nrow >= 1 && ncol >= 1);
    NGX_REQUIRE(images && weights && jac && pixels && npix);
    if (n == 0) return NGX_OK;
    k_make_pixels<<<n, 256, 0, as_stream(stream)>>>(
        nrow, ncol, images, weights, jac, ignore_zero_weight, pixels, npix);
    NGX_LAUNCH_CHECK();
    return NGX_OK;
}

int ngx_make_coords(int n, int nrow, int ncol, const ngx_jacobian *jac,
                    ngx_coord *coords, void *stream) {
    NGX_REQUIRE(n >= 0 && nrow >= 1 && ncol >= 1 && jac && coords);
    if (n == 0) return NGX_OK;
    dim3 grid(n, (nrow * ncol + 255) / 256);   // objects on x: no 65535 limit
    k_make_coords<<<grid, 256, 0, as_stream(stream)>>>(nrow, ncol, jac,
                                                       coords);
    NGX_LAUNCH_CHECK();
    return NGX_OK;
}

int ngx_render(int n, int ngauss, const ngx_gauss *gmix,
               const ngx_jacobian *jac, int nrow, int ncol, int fast_exp,
               double *images, int32_t *status, void *stream) {
    NGX_REQUIRE(n >= 0 && ngauss >= 1 && ngauss <= 4096);
    NGX_REQUIRE(gmix && jac && images && nrow >= 1 && ncol >= 1);
    if (n == 0) return NGX_OK;
    size_t smem = 16 * sizeof(double)
        + (size_t)ngauss * (sizeof(GaussV) + sizeof(ngx_gauss));
    NGX_REQUIRE(smem <= 48 * 1024);
    dim3 grid(n, (nrow * ncol + 255) / 256);   // objects on x: no 65535 limit
    k_render<<<grid, 256, smem, as_stream(stream)>>>(
        ngauss, gmix, jac, nrow, ncol, fast_exp, images, status);
    NGX_LAUNCH_CHECK();
    return NGX_OK;
}

static int launch_loglike(int src, int n, int ngauss, const ngx_gauss *gmix,
                          const double *images, const double *weights,
                          const ngx_jacobian *jac, int nrow, int ncol,
                          int ignore_zero_weight, const ngx_pixel *pixels,
                          int maxpix, const int32_t *npix, double *out4,
                          double *fdiff, int32_t *status, void *stream) {
    size_t smem = (16 + 32 * 9) * sizeof(double)
        + (size_t)ngauss * sizeof(GaussV);
    NGX_REQUIRE(smem <= 48 * 1024);
    int ntot = src == 0 ? nrow * ncol : maxpix;
    int threads = 256;
    if (ntot <= 1024) threads = 128;
    if (src == 0) {
        int tpr = (ncol + 7) / 8;
        int nq = 32 * tpr * ((nrow + 3) / 4);
        size_t smt = sizeof(double)
            * (32 * 9 + LLT_REC * (size_t)ngauss
               + ((size_t)(nq / 32) * ((ngauss + 31) / 32) + 1) / 2
               + 2 * (size_t)nq + 1)
            + sizeof(int) * ((size_t)nrow * tpr + nrow);
        if (smt > (size_t)smem_optin_bytes()) {
            // the stamp does not fit shared memory: build the reference's
            // pixel records in a stream-ordered scratch buffer and run the
            // pixel-record kernel on them (same values, same fdiff order)
            cudaStream_t st = as_stream(stream);
            ngx_pixel *pix = nullptr;
            int32_t *np = nullptr;
            NGX_CUDA(cudaMallocAsync((void **)&pix,
                                     sizeof(ngx_pixel) * (size_t)n * ntot,
                                     st));
            NGX_CUDA(cudaMallocAsync((void **)&np, sizeof(int32_t) * n, st));
            k_make_pixels<<<n, 256, 0, st>>>(nrow, ncol, images, weights, jac,
                                            ignore_zero_weight, pix, np);
            k_loglike_pixels<<<n, 256, smem, st>>>(ngauss, gmix, pix, ntot, np,
                                                  out4, fdiff, status);
            cudaError_t e = cudaGetLastError();
            cudaFreeAsync(pix, st);
            cudaFreeAsync(np, st);
            NGX_CUDA(e);
            return NGX_OK;
        }
        NGX_CUDA(cudaFuncSetAttribute(
            k_loglike_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize,
            (int)smt));
        k_loglike_tiled<<<n, 256, smt, as_stream(stream)>>>(
            ngauss, gmix, images, weights, jac, nrow, ncol, tpr, nq,
            ignore_zero_weight, out4, fdiff, status);
    } else
        k_loglike_pixels<<<n, threads, smem, as_stream(stream)>>>(
            ngauss, gmix, pixels, maxpix, npix, out4, fdiff, status);
    NGX_LAUNCH_CHECK();
    return NGX_OK;
}

int ngx_loglike(int n, int ngauss, const ngx_gauss *gmix, const double *images,
                const double *weights, const ngx_jacobian *jac, int nrow,
                int ncol, int ignore_zero_weight, double *out4, double *fdiff,
                int32_t *status, void *stream) {
    NGX_REQUIRE(n >= 0 && ngauss >= 1 && nrow >= 1 && ncol >= 1);
    NGX_REQUIRE(gmix && images && weights && jac && out4);
    if (n == 0) return NGX_OK;
    return launch_loglike(0, n, ngauss, gmix, images, weights, jac, nrow, ncol,
                          ignore_zero_weight, nullptr, 0, nullptr, out4, fdiff,
                          status, stream);
}

int ngx_loglike_pixels(int n, int ngauss, const ngx_gauss *gmix,
                       const ngx_pixel *pixels, int maxpix,
                       const int32_t *npix, double *out4, double *fdiff,
                       int32_t *status, void *stream) {
    NGX_REQUIRE(n >= 0 && ngauss >= 1 && maxpix >= 1);
    NGX_REQUIRE(gmix && pixels && npix && out4);
    if (n == 0) return NGX_OK;
    return launch_loglike(1, n, ngauss, gmix, nullptr, nullptr, nullptr, 0, 0,
                          0, pixels, maxpix, npix, out4, fdiff, status,
                          stream);
}

int ngx_deriv_images(int n, int ngauss, const double *gpars,
                     const double *dcov, const ngx_jacobian *jac, int nrow,
                     int ncol, double *out, void *stream) {
    NGX_REQUIRE(n >= 0 && ngauss >= 1 && gpars && dcov && jac && out);
    NGX_REQUIRE(nrow >= 1 && ncol >= 1);
    if (n == 0) return NGX_OK;
    size_t smem = (16 + 16 * (size_t)ngauss) * sizeof(double);
    NGX_REQUIRE(smem <= 48 * 1024);
    dim3 grid(n, (nrow * ncol + 127) / 128);
    k_deriv_images<<<grid, 128, smem, as_stream(stream)>>>(
        ngauss, gpars, dcov, jac, nrow, ncol, out);
    NGX_LAUNCH_CHECK();
    return NGX_OK;
}

int ngx_peak_dfma(int iters, double *dfma_per_s, double *ms_out) {
    NGX_REQUIRE(iters > 0 && dfma_per_s);
    DevBuf<double> sink(1);
    NGX_CUDA(sink.err);
    int blocks = sm_count() * 8, threads = 256;
    cudaEvent_t e0, e1;
    NGX_CUDA(cudaEventCreate(&e0));
    NGX_CUDA(cudaEventCreate(&e1));
    k_peak_dfma<<<blocks, threads>>>(iters / 10 + 1, sink.d);   // warm-up
    NGX_CUDA(cudaEventRecord(e0));
    k_peak_dfma<<<blocks, threads>>>(iters, sink.d);
    NGX_CUDA(cudaEventRecord(e1));
    NGX_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    NGX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    double ops = (double)blocks * threads * 8.0 * iters;
    *dfma_per_s = ops / (ms * 1e-3);
    if (ms_out) *ms_out = ms;
    return NGX_OK;
}

int ngx_peak_dmma(int iters, int mode, double *dmma_per_s, double *ms_out) {
    NGX_REQUIRE(iters > 0 && dmma_per_s);
    DevBuf<double> sink(1);
    NGX_CUDA(sink.err);
    int blocks = sm_count() * 8, threads = 256;
    cudaEvent_t e0, e1;
    NGX_CUDA(cudaEventCreate(&e0));
    NGX_CUDA(cudaEventCreate(&e1));
    k_peak_dmma<<<blocks, threads>>>(iters / 10 + 1, mode, sink.d);
    NGX_CUDA(cudaEventRecord(e0));
    k_peak_dmma<<<blocks, threads>>>(iters, mode, sink.d);
    NGX_CUDA(cudaEventRecord(e1));
    NGX_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    NGX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    // warp-level DMMA instructions per second
    double ops = (double)blocks * (threads / 32) * 4.0 * iters;
    *dmma_per_s = ops / (ms * 1e-3);
    if (ms_out) *ms_out = ms;
    return NGX_OK;
}

// ------------------------------------------------------------------ host
#define NGX_BUF(b) NGX_CUDA((b).err)

int ngx_gmix_fill_host(int n, int model, int ngauss, const double *pars,
                       int npars, ngx_gauss *gmix, int32_t *status) {
    NGX_REQUIRE(n >= 0 && pars && gmix);
    if (n == 0) return NGX_OK;
    DevBuf<double> dp((size_t)n * npars, pars);
    DevBuf<ngx_gauss> dg((size_t)n * ngauss);
    DevBuf<int32_t> ds(n);
    NGX_BUF(dp); NGX_BUF(dg); NGX_BUF(ds);
    NGX_CUDA(dg.zero());
    int rc = ngx_gmix_fill(n, model, ngauss, dp.d, npars, dg.d, ds.d, 0);
    if (rc) return rc;
    NGX_CUDA(cudaDeviceSynchronize());
    NGX_CUDA(dg.to_host(gmix));
    NGX_CUDA(ds.to_host(status));
    return NGX_OK;
}

int ngx_gmix_fill_cm_host(int n, const double *cm, const double *pars,
                          ngx_gauss *gmix, int32_t *status) {
    NGX_REQUIRE(n >= 0 && cm && pars && gmix);
    if (n == 0) return NGX_OK;
    DevBuf<double> dc((size_t)n * 3, cm), dp((size_t)n * 6, pars);
    DevBuf<ngx_gauss> dg((size_t)n * 16);
    DevBuf<int32_t> ds(n);
    NGX_BUF(dc); NGX_BUF(dp); NGX_BUF(dg); NGX_BUF(ds);
    NGX_CUDA(dg.zero());
    int rc = ngx_gmix_fill_cm(n, dc.d, dp.d, dg.d, ds.d, 0);
    if (rc) return rc;
    NGX_CUDA(cudaDeviceSynchronize());
    NGX_CUDA(dg.to_host(gmix));
    NGX_CUDA(ds.to_host(status));
    return NGX_OK;
}

int ngx_gmix_set_norms_host(int n, int ngauss, ngx_gauss *gmix,
                            int32_t *status) {
    NGX_REQUIRE(n >= 0 && gmix);
    if (n == 0) return NGX_OK;
    DevBuf<ngx_gauss> dg((size_t)n * ngauss, gmix);
    DevBuf<int32_t> ds(n);
    NGX_BUF(dg); NGX_BUF(ds);
    int rc = ngx_gmix_set_norms(n, ngauss, dg.d, ds.d, 0);
    if (rc) return rc;
    NGX_CUDA(cudaDeviceSynchronize());
    NGX_CUDA(dg.to_host(gmix));
    NGX_CUDA(ds.to_host(status));
    return NGX_OK;
}

int ngx_gmix_convolve_host(int n, int nobj, int npsf, const ngx_gauss *obj,
                           const ngx_gauss *psf, ngx_gauss *out) {
    NGX_REQUIRE(n >= 0 && obj && psf && out);
    if (n == 0) return NGX_OK;
    DevBuf<ngx_gauss> dobj((size_t)n * nobj, obj), dpsf((size_t)n * npsf, psf);
    DevBuf<ngx_gauss> dout((size_t)n * nobj * npsf);
    NGX_BUF(dobj); NGX_BUF(dpsf); NGX_BUF(dout);
    int rc = ngx_gmix_convolve(n, nobj, npsf, dobj.d, dpsf.d, dout.d, 0);
    if (rc) return rc;
    NGX_CUDA(cudaDeviceSynchronize());
    NGX_CUDA(dout.to_host(out));
    return NGX_OK;
}

int ngx_make_pixels_host(int n, int nrow, int ncol, const double *images,
                         const double *weights, const ngx_jacobian *jac,
                         int ignore_zero_weight, ngx_pixel *pixels,
                         int32_t *npix) {
    NGX_REQUIRE(n >= 0 && images && weights && jac && pixels && npix);
    if (n == 0) return NGX_OK;
    size_t ntot = (size_t)n * nrow * ncol;
    DevBuf<double> di(ntot, images), dw(ntot, weights);
    DevBuf<ngx_jacobian> dj(n, jac);
    DevBuf<ngx_pixel> dp(ntot);
    DevBuf<int32_t> dn(n);
    NGX_BUF(di); NGX_BUF(dw); NGX_BUF(dj); NGX_BUF(dp); NGX_BUF(dn);
    NGX_CUDA(dp.zero());
    int rc = ngx_make_pixels(n, nrow, ncol, di.d, dw.d, dj.d,
                             ignore_zero_weight, dp.d, dn.d, 0);
    if (rc) return rc;
    NGX_CUDA(cudaDeviceSynchronize());
    NGX_CUDA(dp.to_host(pixels));
    NGX_CUDA(dn.to_host(npix));
    return NGX_OK;
}

int ngx_make_coords_host(int n, int nrow, int ncol, const ngx_jacobian *jac,
                         ngx_coord *coords) {
    NGX_REQUIRE(n >= 0 && jac && coords);
    if (n == 0) return NGX_OK;
    DevBuf<ngx_jacobian> dj(n, jac);
    DevBuf<ngx_coord> dc((size_t)n * nrow * ncol);
    NGX_BUF(dj); NGX_BUF(dc);
    int rc = ngx_make_coords(n, nrow, ncol, dj.d, dc.d, 0);
    if (rc) return rc;
    NGX_CUDA(cudaDeviceSynchronize());
    NGX_CUDA(dc.to_host(coords));
    return NGX_OK;
}

int ngx_render_host(int n, int ngauss, const ngx_gauss *gmix,
                    const ngx_jacobian *jac, int nrow, int ncol, int fast_exp,
                    double *images, int32_t *status) {
    NGX_REQUIRE(n >= 0 && gmix && jac && images);
    if (n == 0) return NGX_OK;
    DevBuf<ngx_gauss> dg((size_t)n * ngauss, gmix);
    DevBuf<ngx_jacobian> dj(n, jac);
    DevBuf<double> di((size_t)n * nrow * ncol, images);
    DevBuf<int32_t> ds(n);
    NGX_BUF(dg); NGX_BUF(dj); NGX_BUF(di); NGX_BUF(ds);
    int rc = ngx_render(n, ngauss, dg.d, dj.d, nrow, ncol, fast_exp, di.d,
                        ds.d, 0);
    if (rc) return rc;
    NGX_CUDA(cudaDeviceSynchronize());
    NGX_CUDA(di.to_host(images));
    NGX_CUDA(ds.to_host(status));
    return NGX_OK;
}

int ngx_loglike_host(int n, int ngauss, const ngx_gauss *gmix,
                     const double *images, const double *weights,
                     const ngx_jacobian *jac, int nrow, int ncol,
                     int ignore_zero_weight, double *out4, double *fdiff,
                     int32_t *status) {
    NGX_REQUIRE(n >= 0 && gmix && images && weights && jac && out4);
    if (n == 0) return NGX_OK;
    size_t ntot = (size_t)n * nrow * ncol;
    DevBuf<ngx_gauss> dg((size_t)n * ngauss, gmix);
    DevBuf<double> di(ntot, images), dw(ntot, weights);
    DevBuf<ngx_jacobian> dj(n, jac);
    DevBuf<double> dout((size_t)n * 4);
    DevBuf<double> dfd(fdiff ? ntot : 0);
    DevBuf<int32_t> ds(n);
    NGX_BUF(dg); NGX_BUF(di); NGX_BUF(dw); NGX_BUF(dj); NGX_BUF(dout);
    NGX_BUF(dfd); NGX_BUF(ds);
    NGX_CUDA(dfd.zero());
    int rc = ngx_loglike(n, ngauss, dg.d, di.d, dw.d, dj.d, nrow, ncol,
                         ignore_zero_weight, dout.d, fdiff ? dfd.d : nullptr,
                         ds.d, 0);
    if (rc) return rc;
    NGX_CUDA(cudaDeviceSynchronize());
    NGX_CUDA(dout.to_host(out4));
    NGX_CUDA(dfd.to_host(fdiff));
    NGX_CUDA(ds.to_host(status));
    return NGX_OK;
}

int ngx_loglike_pixels_host(int n, int ngauss, const ngx_gauss *gmix,
                            const ngx_pixel *pixels, int maxpix,
                            const int32_t *npix, double *out4, double *fdiff,
                            int32_t *status) {
    NGX_REQUIRE(n >= 0 && gmix && pixels && npix && out4);
    if (n == 0) return NGX_OK;
    size_t ntot = (size_t)n * maxpix;
    DevBuf<ngx_gauss> dg((size_t)n * ngauss, gmix);
    DevBuf<ngx_pixel> dp(ntot, pixels);
    DevBuf<int32_t> dn(n, npix);
    DevBuf<double> dout((size_t)n * 4);
    DevBuf<double> dfd(fdiff ? ntot : 0);
    DevBuf<int32_t> ds(n);
    NGX_BUF(dg); NGX_BUF(dp); NGX_BUF(dn); NGX_BUF(dout); NGX_BUF(dfd);
    NGX_BUF(ds);
    NGX_CUDA(dfd.zero());
    int rc = ngx_loglike_pixels(n, ngauss, dg.d, dp.d, maxpix, dn.d, dout.d,
                                fdiff ? dfd.d : nullptr, ds.d, 0);
    if (rc) return rc;
    NGX_CUDA(cudaDeviceSynchronize());
    NGX_CUDA(dout.to_host(out4));
    NGX_CUDA(dfd.to_host(fdiff));
    NGX_CUDA(ds.to_host(status));
    return NGX_OK;
}

int ngx_deriv_images_host(int n, int ngauss, const double *gpars,
                          const double *dcov, const ngx_jacobian *jac,
                          int nrow, int ncol, double *out) {
    NGX_REQUIRE(n >= 0 && gpars && dcov && jac && out);
    if (n == 0) return NGX_OK;
    DevBuf<double> dgp((size_t)n * ngauss * 6, gpars);
    DevBuf<double> ddc((size_t)n * ngauss * 9, dcov);
    DevBuf<ngx_jacobian> dj(n, jac);
    DevBuf<double> dout((size_t)n * 6 * nrow * ncol, out);
    NGX_BUF(dgp); NGX_BUF(ddc); NGX_BUF(dj); NGX_BUF(dout);
    int rc = ngx_deriv_images(n, ngauss, dgp.d, ddc.d, dj.d, nrow, ncol,
                              dout.d, 0);
    if (rc) return rc;
    NGX_CUDA(cudaDeviceSynchronize());
    NGX_CUDA(dout.to_host(out));
    return NGX_OK;
}

}  // extern "C"
