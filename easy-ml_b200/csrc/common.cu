// common.cu -- errors, per-device contexts, workspace.
#include "common.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <vector>
#include <memory>
#include <string>

namespace eml {

namespace {
thread_local std::string t_error;
thread_local const char* t_kernel = "none";
std::atomic<int64_t> g_launches{0};
std::mutex g_ctx_mu;
std::map<int, std::unique_ptr<DeviceCtx>> g_ctx;
}  // namespace

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    t_error = buf;
}

const char* last_error() { return t_error.c_str(); }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d in `%s`", (int)e, cudaGetErrorString(e), file, line, what);
    if (e == cudaErrorMemoryAllocation) return EML_ERR_OOM;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return EML_ERR_NO_DEVICE;
    return EML_ERR_CUDA;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int64_t launches() { return g_launches.load(std::memory_order_relaxed); }
void set_last_kernel(const char* name) { t_kernel = name; }
const char* last_kernel() { return t_kernel; }

namespace {
thread_local bool t_pdl_suppressed = false;
}
bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("EML_PDL");
        return !(e && e[0] == '0');
    }();
    return on && !t_pdl_suppressed;
}
void pdl_suppress(bool suppress) {
    // Programmatic dependent launch stays on inside recorded steps: the capture turns the launch attribute into
    // programmatic graph edges (NN step replay 0.200 -> 0.195 ms).  EML_GRAPH_PDL=0: plain edges.
    static const bool keep_in_graphs = [] {
        const char* e = getenv("EML_GRAPH_PDL");
        return !(e && e[0] == '0');
    }();
    t_pdl_suppressed = suppress && !keep_in_graphs;
}

namespace {
struct TraceMark {
    const char* what;
    cudaStream_t stream;
    cudaEvent_t ev;
};
std::mutex g_trace_mu;
std::vector<TraceMark> g_trace;
thread_local cudaStream_t t_last_stream = nullptr;
bool trace_on() {
    static const bool on = [] {
        const char* e = getenv("EML_TRACE");
        return e && e[0] == '1';
    }();
    return on;
}
void trace_mark(const char* what) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(t_last_stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return;
    cudaEvent_t ev;
    if (cudaEventCreate(&ev) != cudaSuccess) return;
    cudaEventRecord(ev, t_last_stream);
    std::lock_guard<std::mutex> lock(g_trace_mu);
    if (g_trace.size() >= 8192) {
        for (size_t i = 0; i < 4096; i++) cudaEventDestroy(g_trace[i].ev);
        g_trace.erase(g_trace.begin(), g_trace.begin() + 4096);
    }
    g_trace.push_back(TraceMark{what, t_last_stream, ev});
}
}  // namespace
void trace_note_stream(cudaStream_t stream) { t_last_stream = stream; }
void trace_dump() {
    if (!trace_on()) return;
    std::lock_guard<std::mutex> lock(g_trace_mu);
    const size_t n = g_trace.size(), first = n > 80 ? n - 80 : 0;
    for (size_t i = first; i < n; i++) {
        float ms = 0;
        cudaEventElapsedTime(&ms, g_trace[first].ev, g_trace[i].ev);
        fprintf(stderr, "trace %9.1f us  stream %p  %s\n", ms * 1e3, (void*)g_trace[i].stream, g_trace[i].what);
    }
    for (auto& m : g_trace) cudaEventDestroy(m.ev);
    g_trace.clear();
}

int check_launch(const char* what) {
    if (trace_on()) trace_mark(what);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
        return EML_ERR_CUDA;
    }
    return EML_OK;
}

int get_ctx(DeviceCtx** out) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("no CUDA device present: the f32/f64 path has no CPU fallback (cudaGetDeviceCount: %s)",
                  e == cudaSuccess ? "0 devices" : cudaGetErrorString(e));
        return EML_ERR_NO_DEVICE;
    }
    int dev = 0;
    EML_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_ctx_mu);
    auto it = g_ctx.find(dev);
    if (it == g_ctx.end()) {
        auto ctx = std::make_unique<DeviceCtx>();
        ctx->device = dev;
        cudaDeviceProp prop;
        EML_CUDA(cudaGetDeviceProperties(&prop, dev));
        if (prop.major != 10) {
            set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", dev, prop.major,
                      prop.minor);
            return EML_ERR_NO_DEVICE;
        }
        ctx->sm_count = prop.multiProcessorCount;
        ctx->smem_optin = prop.sharedMemPerBlockOptin;
        EML_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        EML_CUDA(cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
        EML_CUDA(cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
        for (auto& ev : ctx->events) EML_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        cudaMemPool_t pool;
        EML_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
        uint64_t keep = UINT64_MAX;
        EML_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        it = g_ctx.emplace(dev, std::move(ctx)).first;
    }
    *out = it->second.get();
    return EML_OK;
}

namespace {
struct CacheKey {
    int device;
    cudaStream_t stream;
    size_t bytes;
    bool operator<(const CacheKey& o) const {
        if (device != o.device) return device < o.device;
        if (stream != o.stream) return stream < o.stream;
        return bytes < o.bytes;
    }
};
std::mutex g_cache_mu;
std::map<CacheKey, std::vector<void*>> g_cache;
size_t g_cached_bytes = 0;
constexpr size_t CACHE_LIMIT = size_t(24) << 30;  // beyond this, blocks go straight back to CUDA

inline size_t size_class(size_t bytes) { return bytes < 512 ? 512 : (bytes + 511) & ~size_t(511); }
}  // namespace

namespace {
thread_local bool t_forbid_new_blocks = false;
}
void stream_alloc_forbid_new(bool forbid) { t_forbid_new_blocks = forbid; }

int stream_alloc(void** p, size_t bytes, cudaStream_t stream) {
    const size_t sz = size_class(bytes);
    int dev = 0;
    EML_CUDA(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lock(g_cache_mu);
        auto it = g_cache.find(CacheKey{dev, stream, sz});
        if (it != g_cache.end() && !it->second.empty()) {
            *p = it->second.back();
            it->second.pop_back();
            g_cached_bytes -= sz;
            return EML_OK;
        }
    }
    if (t_forbid_new_blocks) {
        set_error("a buffer of %zu bytes is not in the stream cache while a step is being captured: run the step "
                  "eagerly once (same shapes) before eml_tape_capture_begin", sz);
        return EML_ERR_STATE;
    }
    cudaError_t e = cudaMallocAsync(p, sz, stream);
    if (e == cudaErrorMemoryAllocation) {  // give the cache back and retry once
        cudaGetLastError();
        stream_cache_release_all();
        e = cudaMallocAsync(p, sz, stream);
    }
    if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync", __FILE__, __LINE__);
    return EML_OK;
}

void stream_free(void* p, size_t bytes, cudaStream_t stream) {
    if (!p) return;
    const size_t sz = size_class(bytes);
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lock(g_cache_mu);
        if (g_cached_bytes + sz <= CACHE_LIMIT) {
            g_cache[CacheKey{dev, stream, sz}].push_back(p);
            g_cached_bytes += sz;
            return;
        }
    }
    cudaFreeAsync(p, stream);
}

void stream_cache_release_all() {
    std::lock_guard<std::mutex> lock(g_cache_mu);
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto& kv : g_cache) {
        cudaSetDevice(kv.first.device);
        for (void* q : kv.second) cudaFreeAsync(q, kv.first.stream);
        kv.second.clear();
    }
    g_cache.clear();
    g_cached_bytes = 0;
    cudaSetDevice(cur);
}

int TempBuf::alloc(size_t n, cudaStream_t stream) {
    s = stream;
    bytes = n;
    return stream_alloc(&p, n, stream);
}

int stage_buffer(DeviceCtx* ctx, int slot, size_t bytes, void** out) {
    if (bytes > ctx->stage_bytes[slot]) {
        if (ctx->stage[slot]) {
            EML_CUDA(cudaStreamSynchronize(ctx->stream));
            EML_CUDA(cudaFree(ctx->stage[slot]));
            ctx->stage[slot] = nullptr;
            ctx->stage_bytes[slot] = 0;
        }
        EML_CUDA(cudaMalloc(&ctx->stage[slot], bytes));
        ctx->stage_bytes[slot] = bytes;
    }
    *out = ctx->stage[slot];
    return EML_OK;
}

void shutdown_all() {
    pack_cache_release_all();
    stream_cache_release_all();
    std::lock_guard<std::mutex> lock(g_ctx_mu);
    for (auto& kv : g_ctx) {
        DeviceCtx* c = kv.second.get();
        cudaSetDevice(c->device);
        cudaDeviceSynchronize();
        for (int i = 0; i < 3; i++)
            if (c->stage[i]) cudaFree(c->stage[i]);
        if (c->stream) cudaStreamDestroy(c->stream);
        if (c->copy_in) cudaStreamDestroy(c->copy_in);
        if (c->copy_out) cudaStreamDestroy(c->copy_out);
        for (auto ev : c->events)
            if (ev) cudaEventDestroy(ev);
    }
    g_ctx.clear();
}

}  // namespace eml
