// A small pool of host threads that copies a block of pageable memory into a
// pinned staging buffer in parallel.  cudaMemcpy from pageable memory stages
// through the driver on one thread (~10 GB/s); a few threads reach several
// times that, which is what bounds the *_host calls on ordinary numpy arrays.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include <string.h>
#include <stdlib.h>
#include <stdint.h>
#include <unistd.h>

#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#define NGX_NT_COPY 1
#endif

namespace ngx {

// memcpy into a pinned staging slot that nobody on the host will read: with
// non-temporal stores the destination lines are not read for ownership first
// and do not evict the caller's data from the caches (a third less memory
// traffic per staged byte; the copy engine reads the slot from DRAM anyway)
inline void stage_copy(char *dst, const char *src, size_t n) {
#ifdef NGX_NT_COPY
    if (n >= 4096 && ((uintptr_t)dst & 15) == 0) {
        size_t i = 0;
        for (; i + 64 <= n; i += 64) {
            __m128i a = _mm_loadu_si128((const __m128i *)(src + i));
            __m128i b = _mm_loadu_si128((const __m128i *)(src + i + 16));
            __m128i c = _mm_loadu_si128((const __m128i *)(src + i + 32));
            __m128i d = _mm_loadu_si128((const __m128i *)(src + i + 48));
            _mm_stream_si128((__m128i *)(dst + i), a);
            _mm_stream_si128((__m128i *)(dst + i + 16), b);
            _mm_stream_si128((__m128i *)(dst + i + 32), c);
            _mm_stream_si128((__m128i *)(dst + i + 48), d);
        }
        _mm_sfence();
        if (i < n) memcpy(dst + i, src + i, n - i);
        return;
    }
#endif
    memcpy(dst, src, n);
}

class CopyPool {
public:
    // the process-wide pool (recreated in a forked child, whose threads are
    // gone); `use` serialises callers
    static CopyPool &get() {
        static std::mutex mk;
        static CopyPool *pool = nullptr;
        std::lock_guard<std::mutex> g(mk);
        if (!pool || pool->pid_ != getpid()) pool = new CopyPool();
        return *pool;
    }
    int nthreads() const { return (int)workers_.size() + 1; }

    // dst[0:bytes) = src[0:bytes), split over the pool and the caller
    void copy(void *dst, const void *src, size_t bytes) {
        const int nt = nthreads();
        if (nt == 1 || bytes < (size_t)2 << 20) {
            if (getenv("NGX_STAGE_MEMCPY")) memcpy(dst, src, bytes);
            else stage_copy((char *)dst, (const char *)src, bytes);
            return;
        }
        plain_ = getenv("NGX_STAGE_MEMCPY") != nullptr;
        std::lock_guard<std::mutex> use(use_);
        size_t piece = (bytes / nt + 4095) & ~(size_t)4095;
        {
            std::lock_guard<std::mutex> g(m_);
            dst_ = (char *)dst; src_ = (const char *)src;
            bytes_ = bytes; piece_ = piece;
            pending_.store(nt - 1, std::memory_order_relaxed);
            gen_++;
        }
        cv_.notify_all();
        run_piece(0);
        while (pending_.load(std::memory_order_acquire) != 0)
            std::this_thread::yield();
    }

private:
    CopyPool() : pid_(getpid()) {
        int nt = 6;
        if (const char *e = getenv("NGX_STAGE_THREADS")) nt = atoi(e);
        unsigned hw = std::thread::hardware_concurrency();
        if (hw && nt > (int)hw) nt = (int)hw;
        if (nt < 1) nt = 1;
        if (nt > 32) nt = 32;
        for (int i = 1; i < nt; i++)
            workers_.emplace_back([this, i] { loop(i); });
        for (auto &w : workers_) w.detach();
    }
    void run_piece(int i) {
        size_t lo = piece_ * (size_t)i;
        if (lo >= bytes_) return;
        size_t n = bytes_ - lo < piece_ ? bytes_ - lo : piece_;
        if (plain_) memcpy(dst_ + lo, src_ + lo, n);
        else stage_copy(dst_ + lo, src_ + lo, n);
    }
    void loop(int i) {
        unsigned long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return gen_ != seen; });
                seen = gen_;
            }
            run_piece(i);
            pending_.fetch_sub(1, std::memory_order_release);
        }
    }
    pid_t pid_;
    std::vector<std::thread> workers_;
    std::mutex m_, use_;
    std::condition_variable cv_;
    unsigned long gen_ = 0;
    std::atomic<int> pending_{0};
    char *dst_ = nullptr;
    const char *src_ = nullptr;
    size_t bytes_ = 0, piece_ = 0;
    bool plain_ = false;
};

// Host-to-device copies of pageable memory through a ring of pinned slots: the
// pool fills a free slot while the copy engine drains the others.  One ring
// per calling thread.
class StageRing {
public:
    static constexpr int NSLOT = 4;
    static constexpr size_t SLOT_BYTES = (size_t)16 << 20;
    static StageRing &get() {
        static thread_local StageRing r;
        return r;
    }
    // dst (device) = src (pageable host), asynchronously on `st`; `between`
    // runs after every piece (the caller's chance to do other host work)
    cudaError_t h2d(void *dst, const void *src, size_t bytes, cudaStream_t st,
                    const std::function<int()> *between = nullptr,
                    int *between_rc = nullptr) {
        if (!ring_) {
            cudaError_t e = cudaHostAlloc((void **)&ring_, NSLOT * SLOT_BYTES,
                                          cudaHostAllocDefault);
            if (e != cudaSuccess) { ring_ = nullptr; return e; }
            for (int i = 0; i < NSLOT; i++) {
                e = cudaEventCreateWithFlags(&ev_[i], cudaEventDisableTiming);
                if (e != cudaSuccess) return e;
            }
        }
        CopyPool &pool = CopyPool::get();
        for (size_t off = 0; off < bytes; off += SLOT_BYTES) {
            const size_t nb = bytes - off < SLOT_BYTES ? bytes - off
                                                       : SLOT_BYTES;
            const int slot = (int)(next_++ % NSLOT);
            cudaError_t e = cudaEventSynchronize(ev_[slot]);
            if (e != cudaSuccess) return e;
            char *s = ring_ + (size_t)slot * SLOT_BYTES;
            pool.copy(s, (const char *)src + off, nb);
            e = cudaMemcpyAsync((char *)dst + off, s, nb,
                                cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess) return e;
            e = cudaEventRecord(ev_[slot], st);
            if (e != cudaSuccess) return e;
            if (between) {
                int r = (*between)();
                if (r) {
                    if (between_rc) *between_rc = r;
                    return cudaSuccess;
                }
            }
        }
        return cudaSuccess;
    }

private:
    char *ring_ = nullptr;
    cudaEvent_t ev_[NSLOT];
    unsigned next_ = 0;
};

}  // namespace ngx
