// Host-side helpers shared by the C-ABI translation units: error reporting,
// argument checks and the staging buffers of the *_host entry points.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/ngx.h"

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

}  // namespace ngx
