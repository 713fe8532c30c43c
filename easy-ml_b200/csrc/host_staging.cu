// host_staging.cu -- see host_staging.h
#include "host_staging.h"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <cstdlib>
#include <thread>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace eml {

bool host_pointer_is_pinned(const void* p) {
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged;
}

namespace {

// Staging copies move every byte exactly once and nobody reads the destination soon (the DMA engine reads a slot from
// memory; the caller reads C much later), so the destination is written with non-temporal stores: no read-for-ownership
// of the destination lines (a third less memory traffic than memcpy's cached stores) and the caches keep the sources.
#if defined(__x86_64__)
__attribute__((target("avx2"))) void copy_streaming_avx2(char* dst, const char* src, size_t n) {
    while (n && (reinterpret_cast<uintptr_t>(dst) & 31)) {
        *dst++ = *src++;
        n--;
    }
    size_t blocks = n / 128;
    for (size_t i = 0; i < blocks; i++) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + 64));
        const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst), a);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + 96), d);
        src += 128;
        dst += 128;
    }
    n -= blocks * 128;
    if (n) memcpy(dst, src, n);
    _mm_sfence();
}
#endif

void copy_bytes(void* dst, const void* src, size_t n) {
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2") && !(getenv("EML_COPY_NT") && getenv("EML_COPY_NT")[0] == '0');
    if (avx2 && n >= 4096) {
        copy_streaming_avx2(static_cast<char*>(dst), static_cast<const char*>(src), n);
        return;
    }
#endif
    memcpy(dst, src, n);
}

// A tiny persistent pool: run(n, fn) calls fn(0..n-1) over the workers and the calling thread.  A staging job is short
// (16-32 MB: a few hundred microseconds at the 85 GB/s sixteen cores copy here), so a condition-variable wake-up per job
// (~50-100 us for fifteen sleepers) would halve the throughput: workers SPIN for a short while after a job before they go
// to sleep, and the caller spins for their completion.
class CopyPool {
  public:
    static CopyPool& get() {
        static CopyPool* pool = new CopyPool();  // never destroyed: its detached workers outlive static destruction
        return *pool;
    }
    void run(int64_t n, const std::function<void(int64_t)>& fn) {
        if (n <= 0) return;
        if (n == 1 || workers_.empty()) {
            for (int64_t i = 0; i < n; i++) fn(i);
            return;
        }
        // one job at a time: the job state below is shared with the workers, and callers driving different devices from
        // different threads are not serialised by any per-device mutex
        std::lock_guard<std::mutex> job(run_mu_);
        fn_ = &fn;
        total_ = n;
        next_.store(0, std::memory_order_relaxed);
        done_.store(0, std::memory_order_relaxed);
        generation_.fetch_add(1, std::memory_order_release);
        if (sleepers_.load(std::memory_order_acquire) > 0) {
            std::lock_guard<std::mutex> lock(mu_);
            cv_.notify_all();
        }
        work();
        const int workers = (int)workers_.size();
        for (int spins = 0; done_.load(std::memory_order_acquire) != workers; spins++) {
            if (spins > 2000) std::this_thread::yield();
            else cpu_relax();
        }
        fn_ = nullptr;
    }

  private:
    static void cpu_relax() {
#if defined(__x86_64__)
        _mm_pause();
#endif
    }
    CopyPool() {
        // the caller's thread copies too; EML_COPY_THREADS overrides the total (1 = no workers)
        unsigned hw = std::thread::hardware_concurrency();
        int n = hw > 2 ? (int)std::min(31u, hw - 1) : 0;
        if (const char* e = getenv("EML_COPY_THREADS")) {
            const int want = atoi(e);
            if (want >= 1 && want <= 64) n = want - 1;
        }
        for (int i = 0; i < n; i++) workers_.emplace_back([this] { loop(); });
        for (auto& t : workers_) t.detach();
    }
    void work() {
        for (;;) {
            int64_t i = next_.fetch_add(1, std::memory_order_relaxed);
            if (i >= total_) break;
            (*fn_)(i);
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            // spin ~100 us for the next job of the same call, then sleep
            bool woke = false;
            for (int spins = 0; spins < 20000; spins++) {
                if (generation_.load(std::memory_order_acquire) != seen) {
                    woke = true;
                    break;
                }
                cpu_relax();
            }
            if (!woke) {
                std::unique_lock<std::mutex> lock(mu_);
                sleepers_.fetch_add(1, std::memory_order_release);
                cv_.wait(lock, [&] { return generation_.load(std::memory_order_acquire) != seen; });
                sleepers_.fetch_sub(1, std::memory_order_release);
            }
            seen = generation_.load(std::memory_order_acquire);
            work();
            done_.fetch_add(1, std::memory_order_release);
        }
    }
    std::mutex run_mu_;  // held by the caller for a whole job
    std::mutex mu_;
    std::condition_variable cv_;
    std::vector<std::thread> workers_;
    const std::function<void(int64_t)>* fn_ = nullptr;
    int64_t total_ = 0;
    std::atomic<int64_t> next_{0};
    std::atomic<int> done_{0}, sleepers_{0};
    std::atomic<uint64_t> generation_{0};
};

// Two rings per device, one per direction: an upload never waits for a download.  What matters most is how far the
// staging thread can run AHEAD of the copy engines: with 8 x 32 MB shared by both directions (round 1) the 8192^2 f64
// product took 43-45 ms from pageable memory; with >= 384 MB of slots per direction it takes 34.7-36 ms (page-locked
// caller buffers: 33.4 ms) -- measured sweep on the 16-core host of the B200 box: 6 x 32 MB 44.8, 32 x 8 MB 42.8,
// 12 x 32 MB 36.2, 8 x 64 MB 35.3, 24 x 16 MB 34.7.  Slots are page-locked on first use, so a small problem pins little.
constexpr int MAX_SLOTS = 32;
// slots per direction and their size: EML_SLOTS (2..32) / EML_SLOT_MB (1..256) override the defaults, read once
const int SLOTS = [] {
    int n = 24;
    if (const char* e = getenv("EML_SLOTS")) {
        const int want = atoi(e);
        if (want >= 2 && want <= MAX_SLOTS) n = want;
    }
    return n;
}();
const size_t SLOT_BYTES = [] {
    size_t mb = 16;
    if (const char* e = getenv("EML_SLOT_MB")) {
        const int want = atoi(e);
        if (want >= 1 && want <= 256) mb = (size_t)want;
    }
    return mb << 20;
}();

struct Ring {
    void* slot[2][MAX_SLOTS] = {};
    cudaEvent_t ev[2][MAX_SLOTS] = {};
    bool in_flight[2][MAX_SLOTS] = {};
};
constexpr int UP = 0, DOWN = 1;
std::mutex g_ring_mu;
std::vector<std::pair<int, Ring*>> g_rings;  // per device

int ring_for(int device, Ring** out) {
    std::lock_guard<std::mutex> lock(g_ring_mu);
    for (auto& r : g_rings)
        if (r.first == device) {
            *out = r.second;
            return EML_OK;
        }
    auto* r = new Ring();  // (slots and events are created when first used: acquire)
    g_rings.emplace_back(device, r);
    *out = r;
    return EML_OK;
}

}  // namespace

void parallel_copy_rows(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, int64_t rows) {
    if (rows <= 0 || width == 0) return;
    constexpr size_t BLOCK = size_t(512) << 10;
    if (width >= 2 * BLOCK) {
        // wide rows: every row is cut into ~512 KB pieces so that a single long row still spreads over the pool
        const int64_t pieces = (int64_t)((width + BLOCK - 1) / BLOCK);
        CopyPool::get().run(rows * pieces, [&](int64_t t) {
            const int64_t r = t / pieces;
            const size_t off = (size_t)(t % pieces) * BLOCK, len = std::min(BLOCK, width - off);
            copy_bytes((char*)dst + r * dpitch + off, (const char*)src + r * spitch + off, len);
        });
        return;
    }
    // blocks of whole rows, ~512 KB each
    int64_t rows_per_block = std::max<int64_t>(1, (int64_t)(BLOCK / width));
    int64_t blocks = (rows + rows_per_block - 1) / rows_per_block;
    CopyPool::get().run(blocks, [&](int64_t b) {
        int64_t r0 = b * rows_per_block, r1 = std::min(rows, r0 + rows_per_block);
        if (dpitch == width && spitch == width) {
            copy_bytes((char*)dst + r0 * width, (const char*)src + r0 * width, (size_t)(r1 - r0) * width);
        } else {
            for (int64_t r = r0; r < r1; r++) copy_bytes((char*)dst + r * dpitch, (const char*)src + r * spitch, width);
        }
    });
}

// Downloads whose data has already landed in their slot are copied out to the caller's memory now (never blocks): called
// between the chunks of an upload, so the D2H slots drain while the caller's thread is busy staging the other way.
int StagedCopier::drain_ready() {
    Ring* ring = nullptr;
    EML_TRY(ring_for(ctx_->device, &ring));
    while (!pending_.empty()) {
        const Pending p = pending_.front();  // stream order: the oldest chunk finishes first
        const cudaError_t q = cudaEventQuery(ring->ev[DOWN][p.slot]);
        if (q == cudaErrorNotReady) break;
        if (q != cudaSuccess) return cuda_fail(q, "cudaEventQuery", __FILE__, __LINE__);
        pending_.erase(pending_.begin());
        ring->in_flight[DOWN][p.slot] = false;
        parallel_copy_rows(p.dst, p.dpitch, ring->slot[DOWN][p.slot], p.width, p.width, p.rows);
    }
    return EML_OK;
}

int StagedCopier::acquire(int dir, int* slot) {
    Ring* ring = nullptr;
    EML_TRY(ring_for(ctx_->device, &ring));
    const int s = next_[dir]++ % SLOTS;
    if (!ring->slot[dir][s]) {
        EML_CUDA(cudaHostAlloc(&ring->slot[dir][s], SLOT_BYTES, cudaHostAllocDefault));
        EML_CUDA(cudaEventCreateWithFlags(&ring->ev[dir][s], cudaEventDisableTiming));
    }
    if (dir == DOWN) {
        // a pending chunk in this slot must reach the caller's memory first
        for (size_t i = 0; i < pending_.size(); i++)
            if (pending_[i].slot == s) {
                Pending p = pending_[i];
                pending_.erase(pending_.begin() + i);
                EML_TRY(complete(p));
                break;
            }
    }
    if (ring->in_flight[dir][s]) {
        EML_CUDA(cudaEventSynchronize(ring->ev[dir][s]));
        ring->in_flight[dir][s] = false;
    }
    *slot = s;
    return EML_OK;
}

int StagedCopier::complete(const Pending& p) {
    Ring* ring = nullptr;
    EML_TRY(ring_for(ctx_->device, &ring));
    EML_CUDA(cudaEventSynchronize(ring->ev[DOWN][p.slot]));
    ring->in_flight[DOWN][p.slot] = false;
    parallel_copy_rows(p.dst, p.dpitch, ring->slot[DOWN][p.slot], p.width, p.width, p.rows);
    return EML_OK;
}

int StagedCopier::h2d(void* dst_dev, size_t dpitch, const void* src_host, size_t spitch, size_t width, int64_t rows,
                      cudaStream_t stream) {
    if (rows <= 0 || width == 0) return EML_OK;
    if (host_pointer_is_pinned(src_host)) {
        EML_CUDA(cudaMemcpy2DAsync(dst_dev, dpitch, src_host, spitch, width, rows, cudaMemcpyHostToDevice, stream));
        return EML_OK;
    }
    if (width > SLOT_BYTES) {
        // a row wider than a slot (one whole batch element of a batched contraction, a very long matrix row): cut it
        // into slot-sized pieces, each travelling as a one-row copy
        for (int64_t r = 0; r < rows; r++)
            for (size_t off = 0; off < width; off += SLOT_BYTES)
                EML_TRY(h2d((char*)dst_dev + r * dpitch + off, dpitch, (const char*)src_host + r * spitch + off, spitch,
                            std::min(SLOT_BYTES, width - off), 1, stream));
        return EML_OK;
    }
    Ring* ring = nullptr;
    EML_TRY(ring_for(ctx_->device, &ring));
    const int64_t chunk = std::max<int64_t>(1, (int64_t)(SLOT_BYTES / width));
    for (int64_t r0 = 0; r0 < rows; r0 += chunk) {
        const int64_t n = std::min(chunk, rows - r0);
        int s;
        EML_TRY(acquire(UP, &s));
        parallel_copy_rows(ring->slot[UP][s], width, (const char*)src_host + r0 * spitch, spitch, width, n);
        EML_CUDA(cudaMemcpy2DAsync((char*)dst_dev + r0 * dpitch, dpitch, ring->slot[UP][s], width, width, n,
                                   cudaMemcpyHostToDevice, stream));
        EML_CUDA(cudaEventRecord(ring->ev[UP][s], stream));
        ring->in_flight[UP][s] = true;
        EML_TRY(drain_ready());
    }
    return EML_OK;
}

int StagedCopier::d2h(void* dst_host, size_t dpitch, const void* src_dev, size_t spitch, size_t width, int64_t rows,
                      cudaStream_t stream) {
    if (rows <= 0 || width == 0) return EML_OK;
    if (host_pointer_is_pinned(dst_host)) {
        EML_CUDA(cudaMemcpy2DAsync(dst_host, dpitch, src_dev, spitch, width, rows, cudaMemcpyDeviceToHost, stream));
        return EML_OK;
    }
    if (width > SLOT_BYTES) {
        for (int64_t r = 0; r < rows; r++)
            for (size_t off = 0; off < width; off += SLOT_BYTES)
                EML_TRY(d2h((char*)dst_host + r * dpitch + off, dpitch, (const char*)src_dev + r * spitch + off, spitch,
                            std::min(SLOT_BYTES, width - off), 1, stream));
        return EML_OK;
    }
    Ring* ring = nullptr;
    EML_TRY(ring_for(ctx_->device, &ring));
    const int64_t chunk = std::max<int64_t>(1, (int64_t)(SLOT_BYTES / width));
    for (int64_t r0 = 0; r0 < rows; r0 += chunk) {
        const int64_t n = std::min(chunk, rows - r0);
        EML_TRY(drain_ready());
        int s;
        EML_TRY(acquire(DOWN, &s));
        EML_CUDA(cudaMemcpy2DAsync(ring->slot[DOWN][s], width, (const char*)src_dev + r0 * spitch, spitch, width, n,
                                   cudaMemcpyDeviceToHost, stream));
        EML_CUDA(cudaEventRecord(ring->ev[DOWN][s], stream));
        ring->in_flight[DOWN][s] = true;
        pending_.push_back(Pending{s, (char*)dst_host + r0 * dpitch, dpitch, width, n});
    }
    return EML_OK;
}

int StagedCopier::finish() {
    int rc = EML_OK;
    while (!pending_.empty()) {
        Pending p = pending_.front();
        pending_.erase(pending_.begin());
        int r = complete(p);
        if (r != EML_OK) rc = r;
    }
    return rc;
}

}  // namespace eml
