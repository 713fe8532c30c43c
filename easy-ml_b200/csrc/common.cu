// common.cu -- errors, per-device contexts, workspace.
#include "common.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <map>
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

int check_launch(const char* what) {
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

int TempBuf::alloc(size_t bytes, cudaStream_t stream) {
    s = stream;
    EML_CUDA(cudaMallocAsync(&p, bytes ? bytes : 16, stream));
    return EML_OK;
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
