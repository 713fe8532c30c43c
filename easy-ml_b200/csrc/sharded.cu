// sharded.cu -- eml_gemm_{f32,f64}_sharded: one call from one host thread, C = A * B over the GPUs of one box.
//
// BASELINE config 5 (f64 32768^2 on 2/4/8 B200) behind the drop-in boundary: the caller is Rust's
// `&Matrix<f64> * &Matrix<f64>` (matrix_view_multiplication, reference src/matrices/operations.rs:165-196) holding
// HOST buffers.  Rows of C are independent units (:187-194), so A and C are partitioned into contiguous row slabs, one
// per GPU; B is needed everywhere.  No cross-GPU reduction: every element is one k-ordered accumulation chain on one
// GPU, so the result carries the bits of the single-GPU call.
//
// Data movement (what the PCIe / NVLink topology wants):
//   * A: every GPU uploads its own slab over its own PCIe link, in K panels.
//   * B: K is cut into `rounds` panels; of each panel every GPU uploads 1/G over ITS link and the others pull that
//     piece over NVLink (cudaMemcpyPeerAsync on a side stream: copy engines, no SM) -- no link carries B more than
//     once and no GPU is a single source.  Round p's product  C_slab (+)= A_slab[:, kp] * B[kp, :]  starts as soon as
//     round p's pieces are in; the kernels' `accumulate` continues the k-ordered chains.
//   * C: in the last round the slab is cut into row panels; a finished panel is downloaded while the next one is
//     multiplied.
// One host thread per GPU drives its device (pageable caller memory goes through that device's StagedCopier ring, which
// needs CPU work between copies); cross-GPU ordering is CUDA events, with a host-side flag per event because waiting on
// an event that has not been recorded yet is a no-op.
#include <atomic>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "common.h"
#include "host_staging.h"

namespace eml {
namespace {

constexpr int MAX_GPUS = 16, MAX_ROUNDS = 8;

struct Shared {
    int G = 0, rounds = 0;
    int dev[MAX_GPUS];
    void* dB[MAX_GPUS];
    cudaEvent_t up[MAX_ROUNDS][MAX_GPUS];          // piece (round, gpu) of B has landed in gpu's copy of B
    std::atomic<int> recorded[MAX_ROUNDS][MAX_GPUS];
    std::atomic<int> failed{0};
};

int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// [first, last) of piece `i` when `total` is cut into `parts` pieces whose sizes are multiples of `quantum`
void piece(int64_t total, int parts, int i, int64_t quantum, int64_t* first, int64_t* last) {
    const int64_t each = round_up((total + parts - 1) / parts, quantum);
    *first = each * i < total ? each * i : total;
    *last = *first + each < total ? *first + each : total;
}

int enable_peers(const Shared& sh) {
    for (int a = 0; a < sh.G; a++) {
        EML_CUDA(cudaSetDevice(sh.dev[a]));
        for (int b = 0; b < sh.G; b++) {
            if (a == b) continue;
            int can = 0;
            EML_CUDA(cudaDeviceCanAccessPeer(&can, sh.dev[a], sh.dev[b]));
            if (!can) continue;  // cudaMemcpyPeerAsync still works (staged by the driver), only slower
            cudaError_t e = cudaDeviceEnablePeerAccess(sh.dev[b], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
            cudaGetLastError();
        }
    }
    return EML_OK;
}

template <class T>
int run_gpu(Shared& sh, int g, DeviceCtx* ctx, cudaStream_t peer, const T* A, const T* B, T* C, int64_t M, int64_t N,
            int64_t K, int64_t lda, int64_t ldb, int64_t ldc, T* dA, T* dB, T* dC) {
    const size_t es = sizeof(T);
    int64_t r0, r1;
    piece(M, sh.G, g, 128, &r0, &r1);
    const int64_t rows = r1 - r0;
    cudaStream_t in = ctx->copy_in, cs = ctx->stream, out = ctx->copy_out;
    int ev = 0;
    auto next_event = [&]() { return ctx->events[ev++ % 64]; };
    StagedCopier mover(ctx);
    GemmOptions same_bits;
    same_bits.no_split_k = true;  // the bits may not depend on how the problem was cut
    for (int p = 0; p < sh.rounds; p++) {
        int64_t kr0, kr1;
        piece(K, sh.rounds, p, 32, &kr0, &kr1);
        // my piece of this round's panel of B, over my PCIe link
        int64_t kb0, kb1;
        piece(kr1 - kr0, sh.G, g, 1, &kb0, &kb1);
        kb0 += kr0, kb1 += kr0;
        if (kb1 > kb0) EML_TRY(mover.h2d(dB + kb0 * N, N * es, B + kb0 * ldb, ldb * es, N * es, kb1 - kb0, in));
        EML_CUDA(cudaEventRecord(sh.up[p][g], in));
        sh.recorded[p][g].store(1, std::memory_order_release);
        // my slab's columns of this round
        if (rows > 0 && kr1 > kr0)
            EML_TRY(mover.h2d(dA + kr0, K * es, A + r0 * lda + kr0, lda * es, (kr1 - kr0) * es, rows, in));
        cudaEvent_t a_in = next_event();
        EML_CUDA(cudaEventRecord(a_in, in));
        // the other GPUs' pieces, over NVLink
        for (int step = 1; step < sh.G; step++) {
            const int h = (g + step) % sh.G;  // staggered: at any moment every GPU is pulled from by one peer
            int64_t hb0, hb1;
            piece(kr1 - kr0, sh.G, h, 1, &hb0, &hb1);
            hb0 += kr0, hb1 += kr0;
            if (hb1 <= hb0) continue;
            while (!sh.recorded[p][h].load(std::memory_order_acquire)) {
                if (sh.failed.load()) {
                    set_error("sharded gemm: another GPU's worker failed");
                    return EML_ERR_STATE;
                }
                std::this_thread::yield();
            }
            EML_CUDA(cudaStreamWaitEvent(peer, sh.up[p][h], 0));
            EML_CUDA(cudaMemcpyPeerAsync(dB + hb0 * N, sh.dev[g], static_cast<T*>(sh.dB[h]) + hb0 * N, sh.dev[h],
                                         (size_t)(hb1 - hb0) * N * es, peer));
        }
        cudaEvent_t b_in = next_event();
        EML_CUDA(cudaEventRecord(b_in, peer));
        EML_CUDA(cudaStreamWaitEvent(cs, a_in, 0));
        EML_CUDA(cudaStreamWaitEvent(cs, b_in, 0));
        if (rows <= 0 || kr1 <= kr0) continue;
        if (p + 1 < sh.rounds) {
            EML_TRY(contract_dev<T>(ctx, dA + kr0, dB + kr0 * N, dC, 1, rows, N, kr1 - kr0, Strides3{0, K, 1}, Strides3{0, N, 1},
                                    Strides3{0, N, 1}, p > 0, false, cs, &same_bits));
        } else {
            // last round: row panels, each downloaded while the next one is multiplied
            const int panels = rows >= 2048 ? 4 : 1;
            for (int q = 0; q < panels; q++) {
                int64_t q0, q1;
                piece(rows, panels, q, 128, &q0, &q1);
                if (q1 <= q0) continue;
                EML_TRY(contract_dev<T>(ctx, dA + q0 * K + kr0, dB + kr0 * N, dC + q0 * N, 1, q1 - q0, N, kr1 - kr0,
                                        Strides3{0, K, 1}, Strides3{0, N, 1}, Strides3{0, N, 1}, p > 0, false, cs, &same_bits));
                cudaEvent_t done = next_event();
                EML_CUDA(cudaEventRecord(done, cs));
                EML_CUDA(cudaStreamWaitEvent(out, done, 0));
                EML_TRY(mover.d2h(C + (r0 + q0) * ldc, ldc * es, dC + q0 * N, N * es, N * es, q1 - q0, out));
            }
        }
    }
    EML_TRY(mover.finish());
    EML_CUDA(cudaStreamSynchronize(out));
    EML_CUDA(cudaStreamSynchronize(cs));
    EML_CUDA(cudaStreamSynchronize(in));
    EML_CUDA(cudaStreamSynchronize(peer));
    return EML_OK;
}

template <class T>
int gemm_sharded(int n_gpus, const T* A, const T* B, T* C, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                 int64_t ldc) {
    int n_dev = 0;
    cudaError_t ce = cudaGetDeviceCount(&n_dev);
    if (ce != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        set_error("no CUDA device present: the f32/f64 path has no CPU fallback");
        return EML_ERR_NO_DEVICE;
    }
    if (n_gpus < 1 || n_gpus > MAX_GPUS || n_gpus > n_dev)
        EML_BAD_ARG("sharded gemm: n_gpus = %d, but this box has %d CUDA device(s) (at most %d are used)", n_gpus, n_dev, MAX_GPUS);
    if (M < 0 || N < 0 || K < 1) EML_BAD_ARG("gemm: bad extents M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
    if (lda < K || ldb < N || ldc < N) EML_BAD_ARG("gemm: leading dimension smaller than the row length");
    if (M == 0 || N == 0) return EML_OK;
    if (!A || !B || !C) EML_BAD_ARG("gemm: null operand");
    int home = 0;
    EML_CUDA(cudaGetDevice(&home));
    auto sh = std::make_unique<Shared>();
    sh->G = n_gpus;
    // rounds: the first product starts after 1 / rounds of the uploads; every round costs a launch per GPU and a partial
    // last wave, so small problems take fewer
    sh->rounds = K >= 8192 ? 8 : K >= 2048 ? 4 : K >= 256 ? 2 : 1;
    for (int g = 0; g < n_gpus; g++) sh->dev[g] = g;
    for (int p = 0; p < MAX_ROUNDS; p++)
        for (int g = 0; g < MAX_GPUS; g++) sh->recorded[p][g].store(0);
    EML_TRY(enable_peers(*sh));
    // per-GPU state, created from this thread (device buffers are the contexts' grow-only staging slots)
    std::vector<DeviceCtx*> ctx(n_gpus);
    std::vector<cudaStream_t> peer(n_gpus);
    std::vector<T*> dA(n_gpus), dC(n_gpus);
    std::vector<std::unique_lock<std::mutex>> locks;
    int rc = EML_OK;
    for (int g = 0; g < n_gpus && rc == EML_OK; g++) {
        rc = cudaSetDevice(sh->dev[g]) == cudaSuccess ? EML_OK : EML_ERR_CUDA;
        if (rc == EML_OK) rc = get_ctx(&ctx[g]);
        if (rc != EML_OK) break;
        locks.emplace_back(ctx[g]->mu);
        int64_t r0, r1;
        piece(M, n_gpus, g, 128, &r0, &r1);
        const int64_t rows = r1 - r0 > 0 ? r1 - r0 : 1;
        void *va, *vb, *vc;
        if ((rc = stage_buffer(ctx[g], 0, sizeof(T) * (size_t)(rows * K), &va)) != EML_OK) break;
        if ((rc = stage_buffer(ctx[g], 1, sizeof(T) * (size_t)(K * N), &vb)) != EML_OK) break;
        if ((rc = stage_buffer(ctx[g], 2, sizeof(T) * (size_t)(rows * N), &vc)) != EML_OK) break;
        dA[g] = (T*)va, sh->dB[g] = vb, dC[g] = (T*)vc;
        if (cudaStreamCreateWithFlags(&peer[g], cudaStreamNonBlocking) != cudaSuccess) rc = EML_ERR_CUDA;
        for (int p = 0; p < sh->rounds && rc == EML_OK; p++)
            if (cudaEventCreateWithFlags(&sh->up[p][g], cudaEventDisableTiming) != cudaSuccess) rc = EML_ERR_CUDA;
    }
    std::string first_error;
    if (rc == EML_OK) {
        std::vector<int> rcs(n_gpus, EML_OK);
        std::vector<std::string> msgs(n_gpus);
        std::vector<std::thread> workers;
        for (int g = 0; g < n_gpus; g++)
            workers.emplace_back([&, g] {
                cudaSetDevice(sh->dev[g]);
                rcs[g] = run_gpu<T>(*sh, g, ctx[g], peer[g], A, B, C, M, N, K, lda, ldb, ldc, dA[g], (T*)sh->dB[g], dC[g]);
                if (rcs[g] != EML_OK) {
                    msgs[g] = last_error();  // thread-local: carry it to the caller's thread
                    sh->failed.store(1);
                    // unblock the peers waiting for my pieces
                    for (int p = 0; p < sh->rounds; p++) sh->recorded[p][g].store(1, std::memory_order_release);
                }
            });
        for (auto& w : workers) w.join();
        for (int g = 0; g < n_gpus; g++)
            if (rcs[g] != EML_OK && rc == EML_OK) rc = rcs[g], first_error = msgs[g];
    } else {
        first_error = last_error();
    }
    // every worker has synchronised its own streams; a failed run may have left work in flight on the others
    for (int g = 0; g < n_gpus; g++) {
        if (!ctx[g]) continue;
        cudaSetDevice(sh->dev[g]);
        if (rc != EML_OK) cudaDeviceSynchronize();
        if (peer[g]) cudaStreamDestroy(peer[g]);
        for (int p = 0; p < sh->rounds; p++)
            if (sh->up[p][g]) cudaEventDestroy(sh->up[p][g]);
    }
    locks.clear();
    cudaSetDevice(home);
    if (rc != EML_OK) set_error("sharded gemm: %s", first_error.c_str());
    return rc;
}

}  // namespace
}  // namespace eml

using namespace eml;

extern "C" {

int eml_gemm_f64_sharded(int n_gpus, const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                         int64_t ldb, int64_t ldc) {
    return gemm_sharded<double>(n_gpus, A, B, C, M, N, K, lda, ldb, ldc);
}
int eml_gemm_f32_sharded(int n_gpus, const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                         int64_t ldb, int64_t ldc) {
    return gemm_sharded<float>(n_gpus, A, B, C, M, N, K, lda, ldb, ldc);
}

}  // extern "C"
