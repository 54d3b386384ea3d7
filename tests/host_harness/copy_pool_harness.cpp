// Host check of ngmix_b200/csrc/ngx_copy_pool.h: the pool's parallel,
// non-temporal staging copy must be a plain copy for every size, offset and
// alignment (compiled with g++, no GPU involved).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../ngmix_b200/csrc/ngx_copy_pool.h"

extern "C" int copy_pool_check(int nthreads_expected) {
    ngx::CopyPool &pool = ngx::CopyPool::get();
    if (nthreads_expected > 0 && pool.nthreads() != nthreads_expected)
        return -1;
    const size_t sizes[] = {0, 1, 63, 64, 4095, 4096, 4097, (size_t)1 << 20,
                            ((size_t)2 << 20) + 17, ((size_t)9 << 20) + 4093};
    const size_t src_off[] = {0, 1, 8, 24};
    std::vector<unsigned char> src(((size_t)10 << 20) + 64);
    uint32_t x = 12345u;
    for (auto &b : src) { x = x * 1664525u + 1013904223u; b = x >> 24; }
    void *dst_raw = nullptr;
    if (posix_memalign(&dst_raw, 4096, ((size_t)10 << 20) + 8192)) return -2;
    unsigned char *dst = (unsigned char *)dst_raw;
    int bad = 0;
    for (size_t n : sizes)
        for (size_t so : src_off)
            for (size_t dof : {(size_t)0, (size_t)4096, (size_t)8}) {
                memset(dst, 0xA5, n + dof + 64);
                pool.copy(dst + dof, src.data() + so, n);
                if (memcmp(dst + dof, src.data() + so, n) != 0) bad++;
                // nothing written past the end
                for (int k = 0; k < 16; k++)
                    if (dst[dof + n + k] != 0xA5) bad++;
            }
    free(dst_raw);
    return bad;
}
