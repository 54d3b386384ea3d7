// Host-side helpers shared by the C-ABI translation units: error reporting,
// argument checks and the staging buffers of the *_host entry points.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/ngx.h"
#include "ngx_copy_pool.h"

namespace ngx {

// thread-local last error text (ngx_last_error)
char *err_buf();

inline int fail(int code, const char *fmt, const char *a = "",
                const char *b = "") {
    snprintf(err_buf(), 512, fmt, a, b);
    return code;
}

#define NGX_CUDA(call)                                                      \
    do {                                                                    \
        cudaError_t e_ = (call);                                            \
        if (e_ != cudaSuccess)                                              \
            return ngx::fail(NGX_ERR_CUDA, "%s: %s", #call,                 \
                             cudaGetErrorString(e_));                       \
    } while (0)

#define NGX_REQUIRE(cond)                                                   \
    do {                                                                    \
        if (!(cond))                                                        \
            return ngx::fail(NGX_ERR_ARG, "%s: requirement failed: %s",     \
                             __func__, #cond);                              \
    } while (0)

// check for a launch error without synchronising
#define NGX_LAUNCH_CHECK() NGX_CUDA(cudaGetLastError())

inline cudaStream_t as_stream(void *s) { return (cudaStream_t)s; }

// device buffer for the *_host variants: optional copy-in at construction,
// explicit copy-out.  Errors are latched in `err`.
template <typename T>
struct DevBuf {
    T *d = nullptr;
    size_t count = 0;
    cudaError_t err = cudaSuccess;
    DevBuf(size_t n, const T *host_in = nullptr) : count(n) {
        if (n == 0) return;
        err = cudaMalloc((void **)&d, n * sizeof(T));
        if (err == cudaSuccess && host_in)
            err = cudaMemcpy(d, host_in, n * sizeof(T),
                             cudaMemcpyHostToDevice);
    }
    cudaError_t zero() {
        if (count == 0) return cudaSuccess;
        return cudaMemset(d, 0, count * sizeof(T));
    }
    cudaError_t to_host(T *host_out) {
        if (count == 0 || !host_out) return cudaSuccess;
        return cudaMemcpy(host_out, d, count * sizeof(T),
                          cudaMemcpyDeviceToHost);
    }
    ~DevBuf() {
        if (d) cudaFree(d);
    }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
};

int sm_count();          // cached SM count of the current device
int smem_optin_bytes();  // cached max opt-in dynamic shared memory per block

// Host planes -> device in chunks of ~32 MB on two streams, each chunk's
// kernel queued behind its own copies: the copy of chunk c + 1 overlaps the
// fits of chunk c (pinned planes directly; pageable ones through the pinned
// slots of ngx_copy_pool.h, filled by a few host threads - the driver's own
// staging copy runs on one thread at ~10 GB/s, which is four times the
// kernel's time for admom).  launch(lo, cn, images, weights, stream) queues
// the work for stamps [lo, lo + cn) whose planes are at the given device
// addresses.
template <typename F>
static int planes_pipeline(int n, size_t npix, const double *images,
                           const double *weights, F launch) {
    static thread_local double *d_im[2] = {nullptr, nullptr};
    static thread_local double *d_wt[2] = {nullptr, nullptr};
    static thread_local size_t cap = 0;
    static thread_local cudaStream_t st[2] = {nullptr, nullptr};
    int chunk = (int)(((size_t)32 << 20) / (npix * sizeof(double)));
    if (chunk < 4 * sm_count()) chunk = 4 * sm_count();
    if (chunk > n) chunk = n;
    const size_t need = (size_t)chunk * npix * sizeof(double);
    if (!st[0]) {
        NGX_CUDA(cudaStreamCreate(&st[0]));
        NGX_CUDA(cudaStreamCreate(&st[1]));
    }
    if (need > cap) {
        NGX_CUDA(cudaDeviceSynchronize());
        for (int b = 0; b < 2; b++) {
            cudaFree(d_im[b]); cudaFree(d_wt[b]);
            d_im[b] = d_wt[b] = nullptr;
        }
        cap = 0;
        for (int b = 0; b < 2; b++) {
            NGX_CUDA(cudaMalloc((void **)&d_im[b], need));
            NGX_CUDA(cudaMalloc((void **)&d_wt[b], need));
        }
        cap = need;
    }
    cudaPointerAttributes pa, pb;
    const bool pinned = cudaPointerGetAttributes(&pa, images) == cudaSuccess
        && pa.type == cudaMemoryTypeHost
        && cudaPointerGetAttributes(&pb, weights) == cudaSuccess
        && pb.type == cudaMemoryTypeHost;
    cudaGetLastError();
    int rc = NGX_OK;
    for (int lo = 0, c = 0; lo < n && rc == NGX_OK; lo += chunk, c++) {
        const int b = c & 1;
        const int cn = n - lo < chunk ? n - lo : chunk;
        const size_t bytes = (size_t)cn * npix * sizeof(double);
        // buffer set b was last read by the kernels of chunk c - 2, queued
        // on this same stream
        if (pinned) {
            NGX_CUDA(cudaMemcpyAsync(d_im[b], images + (size_t)lo * npix,
                                     bytes, cudaMemcpyHostToDevice, st[b]));
            NGX_CUDA(cudaMemcpyAsync(d_wt[b], weights + (size_t)lo * npix,
                                     bytes, cudaMemcpyHostToDevice, st[b]));
        } else {
            NGX_CUDA(StageRing::get().h2d(d_im[b],
                                          images + (size_t)lo * npix, bytes,
                                          st[b]));
            NGX_CUDA(StageRing::get().h2d(d_wt[b],
                                          weights + (size_t)lo * npix, bytes,
                                          st[b]));
        }
        rc = launch(lo, cn, d_im[b], d_wt[b], st[b]);
    }
    cudaError_t e0 = cudaStreamSynchronize(st[0]);
    cudaError_t e1 = cudaStreamSynchronize(st[1]);
    if (rc) return rc;
    NGX_CUDA(e0);
    NGX_CUDA(e1);
    return NGX_OK;
}

}  // namespace ngx
