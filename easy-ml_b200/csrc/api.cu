// api.cu -- the C ABI of include/easyml_b200.h: argument checks, kernel dispatch, host staging.
#include <cstdlib>
#include <cstring>

#include "common.h"
#include "host_staging.h"

namespace eml {

namespace {

template <class T>
struct Threshold;
template <>
struct Threshold<double> {
    // DMMA tiles are 128x128: below this the SIMT kernel's 64x64 tiles fill the machine better
    static bool tensor_core(int64_t M, int64_t N, int64_t K) {
        return (M < N ? M : N) >= 32 && K >= 8 && M * N * K >= (int64_t(1) << 21);
    }
};
template <>
struct Threshold<float> {
    static bool tensor_core(int64_t M, int64_t N, int64_t K) {
        return (M < N ? M : N) >= 64 && K >= 32 && M * N * K >= (int64_t(1) << 24);
    }
};

template <class T>
int tensor_core_path(DeviceCtx* ctx, const T* A, const T* B, T* C, int64_t batch, int64_t M, int64_t N,
                     int64_t K, Strides3 sA, Strides3 sB, Strides3 sC, bool accumulate, cudaStream_t stream,
                     bool* handled, GemmOptions* opt);
template <>
int tensor_core_path<double>(DeviceCtx* ctx, const double* A, const double* B, double* C, int64_t batch,
                             int64_t M, int64_t N, int64_t K, Strides3 sA, Strides3 sB, Strides3 sC,
                             bool accumulate, cudaStream_t stream, bool* handled, GemmOptions* opt) {
    return launch_gemm_dmma_f64(ctx, A, B, C, batch, M, N, K, sA, sB, sC, accumulate, stream, handled, opt);
}
template <>
int tensor_core_path<float>(DeviceCtx* ctx, const float* A, const float* B, float* C, int64_t batch, int64_t M,
                            int64_t N, int64_t K, Strides3 sA, Strides3 sB, Strides3 sC, bool accumulate,
                            cudaStream_t stream, bool* handled, GemmOptions* opt) {
    return launch_gemm_tf32x3_f32(ctx, A, B, C, batch, M, N, K, sA, sB, sC, accumulate, stream, handled, opt);
}

}  // namespace

namespace {
template <class T>
struct SplitK;
// f64: the DMMA kernels have no split of their own.  f32: gemm_tf32x3.cu splits inside its own launch, so
// only the SIMT-bound (skinny) shapes are split here.
template <>
struct SplitK<double> {
    static bool wanted(int64_t, int64_t, int64_t) { return true; }
};
template <>
struct SplitK<float> {
    static bool wanted(int64_t M, int64_t N, int64_t K) { return !Threshold<float>::tensor_core(M, N, K); }
};
}  // namespace

template <class T>
int contract_dev(DeviceCtx* ctx, const T* A, const T* B, T* C, int64_t batch, int64_t M, int64_t N, int64_t K,
                 Strides3 sA, Strides3 sB, Strides3 sC, bool accumulate, bool exact, cudaStream_t stream,
                 GemmOptions* opt) {
    GemmOptions local;
    if (!opt) opt = &local;
    opt->remote_done = opt->upper_done = opt->epilogue_done = false;
    // ACC_ZERO_PLUS (see GemmOptions::zero_plus): kernels that cannot write `0 + product` get their zeros here
    const bool zp = accumulate && opt->zero_plus;
    if (zp && !(batch == 1 && sC.c == 1 && sC.r == N && C)) EML_BAD_ARG("contract: zero_plus needs a dense C");
    const int acc_mode = zp ? ACC_ZERO_PLUS : (accumulate ? ACC_ADD : ACC_WRITE);
    auto zeros_first = [&]() -> int {
        if (!zp) return EML_OK;
        const size_t bytes = sizeof(T) * (size_t)(M * N);
        if (((reinterpret_cast<uintptr_t>(C) | bytes) & 15) == 0) {
            ZeroRanges z;
            z.count = 1, z.p[0] = C, z.bytes[0] = bytes;
            return launch_zero_ranges(z, stream);
        }
        EML_CUDA(cudaMemsetAsync(C, 0, bytes, stream));
        return EML_OK;
    };
    // tall-skinny A^T * B (N <= 16, A contiguous along m, long K): its own chunked kernel
    if (!exact && batch == 1 && N >= 1 && N <= 16 && sA.r == 1 && sA.c != 1 && K >= 256 && M >= 32 && A && B && C)
        return launch_gemm_skinny_tn<T>(ctx, A, B, C, M, N, K, sA, sB, sC, acc_mode, stream);
    // thin products (M <= 4 rows against a dense row-major B, long K): one launch, ticketed fixed-order finish
    if (!exact && batch == 1 && A && B && C && K >= 256 && thin_shape(M, N, K) && sB.c == 1 && sB.r == N) {
        EML_TRY(zeros_first());
        return launch_thin_fwd<T>(A, sA.r, sA.c, B, M, N, K, C, sC.r, sC.c, accumulate, stream);
    }
    // short contraction, wide dense output (dH = dY * W2^T of an output layer): B in shared memory, no split
    // (its fma chain starts from +0 when it does not accumulate: that IS 0 + product)
    if (!exact && batch == 1 && A && B && C && sA.c == 1 && sC.c == 1 && smallk_shape(M, N, K, sizeof(T)))
        return launch_gemm_smallk<T>(A, sA.r, B, sB.r, sB.c, C, sC.r, M, N, K, accumulate && !zp, stream);
    // Deterministic split-K for outputs with too few tiles for 148 SMs and a long K (the loss reduction
    // L[1 x batch] * d[batch x 10], dW2 = H^T[512 x batch] * dY[batch x 10], ...): K is cut into S equal
    // chunks which run as S "batches" of one launch into a workspace [S][M][N]; a fixed-order reduction
    // adds them.  S depends on the shape only, so results stay bit-identical run to run.
    if (!exact && !opt->no_split_k && batch == 1 && M > 0 && N > 0 && K >= 512 && A && B && C &&
        SplitK<T>::wanted(M, N, K)) {
        const bool tc = Threshold<T>::tensor_core(M, N, K);
        const int64_t tile = tc ? 128 : 64, min_chunk = tc ? 256 : 64;
        const int64_t tiles = ((M + tile - 1) / tile) * ((N + tile - 1) / tile);
        const int64_t sms = ctx ? ctx->sm_count : 148;
        if (tiles * 2 <= sms) {
            int64_t S = (sms + tiles - 1) / tiles;
            if (S > K / min_chunk) S = K / min_chunk;
            while (S > 1 && (K % S != 0 || (K / S) % 4 != 0)) S--;  // equal, 32-byte aligned chunks
            if (S > 1) {
                const int64_t Kc = K / S;
                TempBuf ws;
                EML_TRY(ws.alloc(sizeof(T) * (size_t)(S * M * N), stream));
                EML_TRY(contract_dev<T>(ctx, A, B, ws.as<T>(), S, M, N, Kc, Strides3{Kc * sA.c, sA.r, sA.c},
                                        Strides3{Kc * sB.r, sB.r, sB.c}, Strides3{M * N, N, 1}, false, false, stream));
                return launch_splitk_reduce<T>(ws.as<T>(), (int)S, C, 1, M, N, sC, acc_mode, stream);
            }
        }
    }
    if (batch < 0 || M < 0 || N < 0) EML_BAD_ARG("contract: negative extent (batch=%lld M=%lld N=%lld)",
                                                 (long long)batch, (long long)M, (long long)N);
    if (K < 1) EML_BAD_ARG("contract: K must be >= 1 (got %lld); Easy ML containers have no zero-length dimension",
                           (long long)K);
    if (batch == 0 || M == 0 || N == 0) return EML_OK;
    if (!A || !B || !C) EML_BAD_ARG("contract: null operand");
    if (!exact && Threshold<T>::tensor_core(M, N, K)) {
        // the tensor-core kernels want C's column index contiguous; a row-contiguous C is the transposed
        // problem C^T = B^T A^T with the same kernels
        bool handled = false;
        int rc;
        // C = G * G^T (the covariance's Gram matrix): A(i, k) and B(k, j) read the same buffer through mirrored
        // strides, so C is symmetric -- the kernels compute the tiles on and above the diagonal only (and the f32
        // path packs G once), then the upper triangle is mirrored.  Worth it from 4 x 4 tiles of 128 up.
        opt->upper_only = (const void*)A == (const void*)B && batch == 1 && M == N && M >= 512 && !accumulate &&
                          sA.r == sB.c && sA.c == sB.r && sC.c == 1 && opt->n_remote == 0;
        // (the f32 kernels honour zero_plus themselves -- their split-K launches end in the reduction; f64: zeros first)
        if (sizeof(T) == 8) EML_TRY(zeros_first());
        if (sC.c != 1 && sC.r == 1) {
            GemmOptions swapped = *opt;  // the operands swap roles in the transposed problem
            swapped.a_id = opt->b_id;
            swapped.b_id = opt->a_id;
            swapped.epilogue = nullptr;  // (its outputs are laid out like C, not like C^T)
            Strides3 tA{sB.b, sB.c, sB.r}, tB{sA.b, sA.c, sA.r}, tC{sC.b, sC.c, sC.r};
            rc = tensor_core_path<T>(ctx, B, A, C, batch, N, M, K, tA, tB, tC, accumulate, stream, &handled, &swapped);
            opt->remote_done = swapped.remote_done;
        } else {
            rc = tensor_core_path<T>(ctx, A, B, C, batch, M, N, K, sA, sB, sC, accumulate, stream, &handled, opt);
        }
        EML_TRY(rc);
        if (handled) return opt->upper_done ? launch_mirror_upper<T>(C, M, sC.r, stream) : EML_OK;
        if (sizeof(T) == 8) opt->zero_plus = false;  // (already zeroed above; the kernels below accumulate)
    }
    if (opt->zero_plus) EML_TRY(zeros_first());
    if (!exact && N <= 16 && sA.c == 1 && M >= 256 && K >= 64)
        return launch_gemm_skinny<T>(A, B, C, batch, M, N, K, sA, sB, sC, accumulate, stream);
    return launch_gemm_simt<T>(A, B, C, batch, M, N, K, sA, sB, sC, accumulate, exact, stream);
}

template int contract_dev<float>(DeviceCtx*, const float*, const float*, float*, int64_t, int64_t, int64_t,
                                 int64_t, Strides3, Strides3, Strides3, bool, bool, cudaStream_t, GemmOptions*);
template int contract_dev<double>(DeviceCtx*, const double*, const double*, double*, int64_t, int64_t, int64_t,
                                  int64_t, Strides3, Strides3, Strides3, bool, bool, cudaStream_t, GemmOptions*);

namespace {

inline Strides3 s3(const int64_t s[3]) { return Strides3{s[0], s[1], s[2]}; }

// number of elements a strided (batch, r, c) view spans from its base pointer
inline int64_t span(int64_t nb, int64_t nr, int64_t nc, Strides3 s) {
    if (nb == 0 || nr == 0 || nc == 0) return 0;
    return 1 + (nb - 1) * s.b + (nr - 1) * s.r + (nc - 1) * s.c;
}

inline bool nonneg(Strides3 s) { return s.b >= 0 && s.r >= 0 && s.c >= 0; }

// Large row-major GEMM from host buffers: PCIe traffic is hidden under the compute instead of bracketing it.
//   phase 1 (while B is still arriving): B travels in K panels; the first M1 rows of C accumulate panel by
//            panel, C[0:M1] (+)= A[0:M1, kq] * B[kq, :]   (the kernels' `accumulate` continues the k-ordered
//            chain, so this equals one launch over the whole K);
//   phase 2 (B resident): the remaining rows in row panels -- A panel p+1 uploads and C panel p-1 downloads
//            while panel p is multiplied.
// Three streams (H2D, compute, D2H) ordered by events.  Page-locked caller buffers (eml_host_alloc) are copied
// directly; pageable ones (a Rust Vec<T>, a numpy array) go through the StagedCopier's ring of page-locked slots
// filled / drained by a thread pool, so the copies stay asynchronous either way.
template <class T>
int gemm_host_pipelined(DeviceCtx* ctx, const T* A, const T* B, T* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                        int64_t ldb, int64_t ldc) {
    std::lock_guard<std::mutex> lock(ctx->mu);
    void *vA, *vB, *vC;
    EML_TRY(stage_buffer(ctx, 0, sizeof(T) * (size_t)(M * K), &vA));
    EML_TRY(stage_buffer(ctx, 1, sizeof(T) * (size_t)(K * N), &vB));
    EML_TRY(stage_buffer(ctx, 2, sizeof(T) * (size_t)(M * N), &vC));
    T *dA = (T*)vA, *dB = (T*)vB, *dC = (T*)vC;  // dense on the device: ld = K, N, N
    cudaStream_t in = ctx->copy_in, cs = ctx->stream, out = ctx->copy_out;
    int ev = 0;
    auto next_event = [&]() { return ctx->events[ev++ % 64]; };
    const size_t es = sizeof(T);
    auto round_to = [](int64_t x, int64_t m) { return (x + m - 1) / m * m; };

    constexpr int Q = 8;                                      // K panels of phase 1
    const int64_t kq = round_to((K + Q - 1) / Q, 32);
    // Every launch covers whole waves of 148 CTAs as nearly as possible: a slab of r tile rows has r * tiles_n tiles,
    // and a 128 x 128 x K tile is ~1 ms of work at K = 8192, so a nearly empty last wave per launch is expensive.
    const int64_t tile = sizeof(T) == 8 ? 128 : 256, tiles_n = (N + tile - 1) / tile;
    const int64_t quantum = sizeof(T) == 8 ? ctx->sm_count : ctx->sm_count / 2;  // CTAs (f64) / CTA pairs (f32) per wave
    auto whole_waves = [&](int64_t target_rows, int64_t max_rows) {
        int64_t best = round_to(target_rows, tile);
        double best_eff = 0;
        for (int64_t r = target_rows / tile - 4; r <= target_rows / tile + 4; r++) {
            if (r < 1 || r * tile > max_rows) continue;
            const int64_t tiles = r * tiles_n;
            const double eff = (double)tiles / (double)(((tiles + quantum - 1) / quantum) * quantum);
            if (eff > best_eff + 1e-9) best_eff = eff, best = r * tile;
        }
        return best < max_rows ? best : max_rows;
    };
    // ~ compute(M1 rows) == upload(B + A[0:M1]) at PCIe 5 / FP64 rates (EML_HOST_M1_PCT overrides)
    int64_t pct = 55;
    if (const char* e = getenv("EML_HOST_M1_PCT")) pct = atoi(e) > 0 && atoi(e) <= 100 ? atoi(e) : pct;
    int64_t M1 = whole_waves(M * pct / 100, M);
    if (M1 > M) M1 = M;
    int64_t R = whole_waves(1152, M - M1 > 0 ? M - M1 : 1);  // row panel of phase 2 (the last one takes the rest)
    if (R < tile) R = tile;
    GemmOptions same_bits;  // same bits as one device launch over the whole problem
    same_bits.no_split_k = true;
    StagedCopier mover(ctx);

    // K panels of phase 1: the first one is split in two so that the GPU starts after half a panel's upload
    std::vector<int64_t> kcut{0};
    {
        const int64_t half = round_to(kq / 2, 32);
        if (half > 0 && 2 * half <= kq && K > kq) kcut.push_back(half);
        for (int64_t k = kq; k < K; k += kq) kcut.push_back(k);
        kcut.push_back(K);
    }
    // Row panels of phase 2: panels of R rows, except that the LAST one is small (its download is the only copy
    // nothing hides) -- the smallest slab that still fills its waves to 85 % -- and the rows that do not fit the
    // pattern go to the first panel.
    std::vector<int64_t> rcut{M1};
    if (M > M1) {
        const int64_t rest_rows = M - M1;
        int64_t tail = 0;
        for (int64_t r = 1; r * tile * 4 <= R; r++) {
            const int64_t tiles = r * tiles_n;
            if ((double)tiles / (double)(((tiles + quantum - 1) / quantum) * quantum) >= 0.85) {
                tail = r * tile;
                break;
            }
        }
        if (tail == 0 || rest_rows < R + tail) tail = 0;
        const int64_t body = rest_rows - tail;
        const int64_t first = body - (body > R ? ((body - R) / R) * R : 0);  // R + remainder (or everything)
        for (int64_t r = M1 + first; r < M - tail; r += R) rcut.push_back(r);
        if (tail) rcut.push_back(M - tail);
        rcut.push_back(M);
    }

    // the previous call's work on these streams is complete (every call ends with a synchronize)
    for (size_t q = 0; q + 1 < kcut.size(); q++) {
        const int64_t k0 = kcut[q], k1 = kcut[q + 1];
        EML_TRY(mover.h2d(dB + k0 * N, N * es, B + k0 * ldb, ldb * es, N * es, k1 - k0, in));
        EML_TRY(mover.h2d(dA + k0, K * es, A + k0, lda * es, (k1 - k0) * es, M1, in));
        cudaEvent_t e = next_event();
        EML_CUDA(cudaEventRecord(e, in));
        EML_CUDA(cudaStreamWaitEvent(cs, e, 0));
        EML_TRY(contract_dev<T>(ctx, dA + k0, dB + k0 * N, dC, 1, M1, N, k1 - k0, Strides3{0, K, 1}, Strides3{0, N, 1},
                                Strides3{0, N, 1}, q > 0, false, cs, &same_bits));
    }
    {
        cudaEvent_t e = next_event();
        EML_CUDA(cudaEventRecord(e, cs));
        EML_CUDA(cudaStreamWaitEvent(out, e, 0));
        EML_TRY(mover.d2h(C, ldc * es, dC, N * es, N * es, M1, out));
    }
    for (size_t p = 0; p + 1 < rcut.size(); p++) {
        const int64_t r0 = rcut[p], r1 = rcut[p + 1];
        EML_TRY(mover.h2d(dA + r0 * K, K * es, A + r0 * lda, lda * es, K * es, r1 - r0, in));
        cudaEvent_t e = next_event();
        EML_CUDA(cudaEventRecord(e, in));
        EML_CUDA(cudaStreamWaitEvent(cs, e, 0));
        EML_TRY(contract_dev<T>(ctx, dA + r0 * K, dB, dC + r0 * N, 1, r1 - r0, N, K, Strides3{0, K, 1},
                                Strides3{0, N, 1}, Strides3{0, N, 1}, false, false, cs, &same_bits));
        cudaEvent_t d = next_event();
        EML_CUDA(cudaEventRecord(d, cs));
        EML_CUDA(cudaStreamWaitEvent(out, d, 0));
        EML_TRY(mover.d2h(C + r0 * ldc, ldc * es, dC + r0 * N, N * es, N * es, r1 - r0, out));
    }
    EML_TRY(mover.finish());
    EML_CUDA(cudaStreamSynchronize(out));
    EML_CUDA(cudaStreamSynchronize(cs));
    EML_CUDA(cudaStreamSynchronize(in));
    return EML_OK;
}

// Batched contraction from host buffers whose batch index is the outermost one in memory (config 3's layout): the
// batches are independent, so the batch range is cut into chunks and chunk c+1 uploads while chunk c is multiplied
// and chunk c-1 downloads -- three streams ordered by events, pageable memory through the StagedCopier.  PCIe stays
// the bound (the product itself is ~20x faster than its transfers) but the two directions now overlap.
template <class T>
int contract_host_batch_pipelined(DeviceCtx* ctx, const T* A, const T* B, T* C, int64_t batch, int64_t M, int64_t N,
                                  int64_t K, Strides3 sA, Strides3 sB, Strides3 sC, int64_t chunk) {
    std::lock_guard<std::mutex> lock(ctx->mu);
    const int64_t ea = span(1, M, K, sA), eb = span(1, K, N, sB), ec = span(1, M, N, sC);  // elements of one batch
    void *vA, *vB, *vC;
    EML_TRY(stage_buffer(ctx, 0, sizeof(T) * (size_t)(batch * ea), &vA));
    EML_TRY(stage_buffer(ctx, 1, sizeof(T) * (size_t)(batch * eb), &vB));
    EML_TRY(stage_buffer(ctx, 2, sizeof(T) * (size_t)(batch * ec), &vC));
    T *dA = (T*)vA, *dB = (T*)vB, *dC = (T*)vC;  // batch b of each operand: its span, densely after batch b-1's
    cudaStream_t in = ctx->copy_in, cs = ctx->stream, out = ctx->copy_out;
    int ev = 0;
    auto next_event = [&]() { return ctx->events[ev++ % 64]; };
    const size_t es = sizeof(T);
    StagedCopier mover(ctx);
    GemmOptions opt;
    opt.no_split_k = true;  // a chunk holding a single batch must not be reformulated differently from the others
    for (int64_t b0 = 0; b0 < batch; b0 += chunk) {
        const int64_t nb = b0 + chunk < batch ? chunk : batch - b0;
        EML_TRY(mover.h2d(dA + b0 * ea, ea * es, A + b0 * sA.b, sA.b * es, ea * es, nb, in));
        EML_TRY(mover.h2d(dB + b0 * eb, eb * es, B + b0 * sB.b, sB.b * es, eb * es, nb, in));
        cudaEvent_t e = next_event();
        EML_CUDA(cudaEventRecord(e, in));
        EML_CUDA(cudaStreamWaitEvent(cs, e, 0));
        EML_TRY(contract_dev<T>(ctx, dA + b0 * ea, dB + b0 * eb, dC + b0 * ec, nb, M, N, K, Strides3{ea, sA.r, sA.c},
                                Strides3{eb, sB.r, sB.c}, Strides3{ec, sC.r, sC.c}, false, false, cs, &opt));
        cudaEvent_t d = next_event();
        EML_CUDA(cudaEventRecord(d, cs));
        EML_CUDA(cudaStreamWaitEvent(out, d, 0));
        EML_TRY(mover.d2h(C + b0 * sC.b, sC.b * es, dC + b0 * ec, ec * es, ec * es, nb, out));
    }
    EML_TRY(mover.finish());
    EML_CUDA(cudaStreamSynchronize(out));
    EML_CUDA(cudaStreamSynchronize(cs));
    EML_CUDA(cudaStreamSynchronize(in));
    return EML_OK;
}

// Host entry: stage operands, run, copy the result back.  Blocks until C is in the caller's buffer.
template <class T>
int contract_host(const T* A, const T* B, T* C, int64_t batch, int64_t M, int64_t N, int64_t K, Strides3 sA,
                  Strides3 sB, Strides3 sC, bool exact) {
    DeviceCtx* ctx = nullptr;
    EML_TRY(get_ctx(&ctx));
    if (batch < 0 || M < 0 || N < 0 || K < 1) EML_BAD_ARG("contract: bad extents");
    if (batch == 0 || M == 0 || N == 0) return EML_OK;
    if (!A || !B || !C) EML_BAD_ARG("contract: null operand");
    if (!nonneg(sA) || !nonneg(sB) || !nonneg(sC)) EML_BAD_ARG("contract: negative strides are not supported");
    int64_t na = span(batch, M, K, sA), nb = span(batch, K, N, sB), nc = span(batch, M, N, sC);
    // many independent batches, each a dense block of its operand (batch outermost in memory, C written densely so
    // that no foreign element sits inside a batch's span), and enough bytes for the bus to matter: overlap the
    // upload of one chunk of batches with the product of the previous one and the download of the one before
    if (!exact && batch >= 4 && sA.b >= span(1, M, K, sA) && sB.b >= span(1, K, N, sB) && sC.b >= span(1, M, N, sC) &&
        span(1, M, N, sC) == M * N && (na + nb + nc) * (int64_t)sizeof(T) >= (int64_t(64) << 20)) {
        int64_t chunk = (batch + 7) / 8;  // eight chunks: the exposed first upload / last download are 1/8 each
        return contract_host_batch_pipelined<T>(ctx, A, B, C, batch, M, N, K, sA, sB, sC, chunk);
    }
    std::lock_guard<std::mutex> lock(ctx->mu);
    void *dA, *dB, *dC;
    EML_TRY(stage_buffer(ctx, 0, sizeof(T) * na, &dA));
    EML_TRY(stage_buffer(ctx, 1, sizeof(T) * nb, &dB));
    EML_TRY(stage_buffer(ctx, 2, sizeof(T) * nc, &dC));
    cudaStream_t st = ctx->stream;
    EML_CUDA(cudaMemcpyAsync(dA, A, sizeof(T) * na, cudaMemcpyHostToDevice, st));
    EML_CUDA(cudaMemcpyAsync(dB, B, sizeof(T) * nb, cudaMemcpyHostToDevice, st));
    // a strided C view may interleave with elements this call must not touch: bring them along
    bool dense_c = (nc == batch * M * N);
    if (!dense_c) EML_CUDA(cudaMemcpyAsync(dC, C, sizeof(T) * nc, cudaMemcpyHostToDevice, st));
    EML_TRY(contract_dev<T>(ctx, (const T*)dA, (const T*)dB, (T*)dC, batch, M, N, K, sA, sB, sC, false, exact, st));
    EML_CUDA(cudaMemcpyAsync(C, dC, sizeof(T) * nc, cudaMemcpyDeviceToHost, st));
    EML_CUDA(cudaStreamSynchronize(st));
    return EML_OK;
}

template <class T>
int gemm_host(const T* A, const T* B, T* C, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc,
              bool exact) {
    if (M < 0 || N < 0 || K < 1) EML_BAD_ARG("gemm: bad extents M=%lld N=%lld K=%lld", (long long)M, (long long)N,
                                             (long long)K);
    if (lda < K || ldb < N || ldc < N) EML_BAD_ARG("gemm: leading dimension smaller than the row length");
    // events available: 8 (phase 1) + 1 + 2 per row panel of phase 2  ->  M <= ~ 27 * 1024 rows after M1
    if (!exact && A && B && C && M >= 2048 && M <= 40960 && N >= 512 && K >= 512 &&
        (M * K + K * N + M * N) >= (int64_t(1) << 24)) {
        DeviceCtx* ctx = nullptr;
        EML_TRY(get_ctx(&ctx));
        return gemm_host_pipelined<T>(ctx, A, B, C, M, N, K, lda, ldb, ldc);
    }
    return contract_host<T>(A, B, C, 1, M, N, K, Strides3{0, lda, 1}, Strides3{0, ldb, 1}, Strides3{0, ldc, 1},
                            exact);
}

template <class T>
int gemm_dev(const T* A, const T* B, T* C, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc,
             bool exact, void* stream) {
    DeviceCtx* ctx = nullptr;
    EML_TRY(get_ctx(&ctx));
    if (lda < K || ldb < N || ldc < N) EML_BAD_ARG("gemm: leading dimension smaller than the row length");
    return contract_dev<T>(ctx, A, B, C, 1, M, N, K, Strides3{0, lda, 1}, Strides3{0, ldb, 1}, Strides3{0, ldc, 1},
                           false, exact, (cudaStream_t)stream);
}

template <class T>
int einsum_host(const T* A, int64_t a_elems, const T* B, int64_t b_elems, T* out, int n_out, const int64_t* out_len,
                const int64_t* ao, const int64_t* bo, int n_sum, const int64_t* sum_len, const int64_t* as,
                const int64_t* bs) {
    DeviceCtx* ctx = nullptr;
    EML_TRY(get_ctx(&ctx));
    if (a_elems < 1 || b_elems < 1 || !A || !B || !out) EML_BAD_ARG("einsum2: empty operand");
    if (n_out < 0 || n_out > 6 || n_sum < 0 || n_sum > 12) EML_BAD_ARG("einsum2: rank out of range");
    int64_t total = 1, amax = 0, bmax = 0;
    for (int d = 0; d < n_out; d++) {
        if (out_len[d] < 0 || ao[d] < 0 || bo[d] < 0) EML_BAD_ARG("einsum2: negative length/stride");
        total *= out_len[d];
        if (out_len[d]) amax += (out_len[d] - 1) * ao[d], bmax += (out_len[d] - 1) * bo[d];
    }
    for (int d = 0; d < n_sum; d++) {
        if (sum_len[d] < 0 || as[d] < 0 || bs[d] < 0) EML_BAD_ARG("einsum2: negative length/stride");
        if (sum_len[d]) amax += (sum_len[d] - 1) * as[d], bmax += (sum_len[d] - 1) * bs[d];
    }
    if (total == 0) return EML_OK;
    if (amax >= a_elems || bmax >= b_elems) EML_BAD_ARG("einsum2: strides reach outside the operand buffers");
    std::lock_guard<std::mutex> lock(ctx->mu);
    void *dA, *dB, *dC;
    EML_TRY(stage_buffer(ctx, 0, sizeof(T) * a_elems, &dA));
    EML_TRY(stage_buffer(ctx, 1, sizeof(T) * b_elems, &dB));
    EML_TRY(stage_buffer(ctx, 2, sizeof(T) * total, &dC));
    cudaStream_t st = ctx->stream;
    EML_CUDA(cudaMemcpyAsync(dA, A, sizeof(T) * a_elems, cudaMemcpyHostToDevice, st));
    EML_CUDA(cudaMemcpyAsync(dB, B, sizeof(T) * b_elems, cudaMemcpyHostToDevice, st));
    EML_TRY(launch_einsum2<T>((const T*)dA, (const T*)dB, (T*)dC, n_out, out_len, ao, bo, n_sum, sum_len, as, bs, st));
    EML_CUDA(cudaMemcpyAsync(out, dC, sizeof(T) * total, cudaMemcpyDeviceToHost, st));
    EML_CUDA(cudaStreamSynchronize(st));
    return EML_OK;
}

template <class T>
int einsum_n_host(int n_ops, const T* const* x, const int64_t* elems, T* out, int n_out, const int64_t* out_len,
                  const int64_t* out_strides, int n_sum, const int64_t* sum_len, const int64_t* sum_strides) {
    DeviceCtx* ctx = nullptr;
    EML_TRY(get_ctx(&ctx));
    if (n_ops < 1 || n_ops > 6 || !x || !elems || !out) EML_BAD_ARG("einsum: 1..6 operands, non-null arguments");
    if (n_out < 0 || n_out > 6 || n_sum < 0 || n_sum > 36) EML_BAD_ARG("einsum: rank out of range");
    int64_t total = 1;
    for (int d = 0; d < n_out; d++) {
        if (out_len[d] < 0) EML_BAD_ARG("einsum: negative length");
        total *= out_len[d];
    }
    for (int i = 0; i < n_ops; i++) {  // every reachable index stays inside its operand
        if (!x[i] || elems[i] < 1) EML_BAD_ARG("einsum: empty operand %d", i);
        int64_t reach = 0;
        for (int d = 0; d < n_out; d++) {
            if (out_strides[i * n_out + d] < 0) EML_BAD_ARG("einsum: negative stride");
            if (out_len[d]) reach += (out_len[d] - 1) * out_strides[i * n_out + d];
        }
        for (int d = 0; d < n_sum; d++) {
            if (sum_len[d] < 0 || sum_strides[i * n_sum + d] < 0) EML_BAD_ARG("einsum: negative length/stride");
            if (sum_len[d]) reach += (sum_len[d] - 1) * sum_strides[i * n_sum + d];
        }
        if (total && reach >= elems[i]) EML_BAD_ARG("einsum: strides reach outside operand %d", i);
    }
    if (total == 0) return EML_OK;
    std::lock_guard<std::mutex> lock(ctx->mu);
    cudaStream_t st = ctx->stream;
    TempBuf dx[6], dout;
    const T* dptr[6];
    for (int i = 0; i < n_ops; i++) {
        EML_TRY(dx[i].alloc(sizeof(T) * (size_t)elems[i], st));
        EML_CUDA(cudaMemcpyAsync(dx[i].p, x[i], sizeof(T) * (size_t)elems[i], cudaMemcpyHostToDevice, st));
        dptr[i] = dx[i].template as<T>();
    }
    EML_TRY(dout.alloc(sizeof(T) * (size_t)total, st));
    EML_TRY(launch_einsum_n<T>(n_ops, dptr, dout.template as<T>(), n_out, out_len, out_strides, n_sum, sum_len, sum_strides, st));
    EML_CUDA(cudaMemcpyAsync(out, dout.p, sizeof(T) * (size_t)total, cudaMemcpyDeviceToHost, st));
    EML_CUDA(cudaStreamSynchronize(st));
    return EML_OK;
}

template <class T>
int dot_host(const T* x, int64_t incx, const T* y, int64_t incy, int64_t n, T* out) {
    DeviceCtx* ctx = nullptr;
    EML_TRY(get_ctx(&ctx));
    if (n < 1 || incx < 1 || incy < 1 || !x || !y || !out) EML_BAD_ARG("dot: n >= 1 and positive increments required");
    std::lock_guard<std::mutex> lock(ctx->mu);
    void *dx, *dy, *dout;
    int64_t nx = 1 + (n - 1) * incx, ny = 1 + (n - 1) * incy;
    EML_TRY(stage_buffer(ctx, 0, sizeof(T) * nx, &dx));
    EML_TRY(stage_buffer(ctx, 1, sizeof(T) * ny, &dy));
    EML_TRY(stage_buffer(ctx, 2, sizeof(T), &dout));
    cudaStream_t st = ctx->stream;
    EML_CUDA(cudaMemcpyAsync(dx, x, sizeof(T) * nx, cudaMemcpyHostToDevice, st));
    EML_CUDA(cudaMemcpyAsync(dy, y, sizeof(T) * ny, cudaMemcpyHostToDevice, st));
    EML_TRY(launch_dot<T>((const T*)dx, incx, (const T*)dy, incy, n, (T*)dout, st));
    EML_CUDA(cudaMemcpyAsync(out, dout, sizeof(T), cudaMemcpyDeviceToHost, st));
    EML_CUDA(cudaStreamSynchronize(st));
    return EML_OK;
}

template <class T>
int matmul_bwd_dev(const T* dC, const T* A, const T* B, T* dA, T* dB, int64_t M, int64_t N, int64_t K,
                   void* stream) {
    DeviceCtx* ctx = nullptr;
    EML_TRY(get_ctx(&ctx));
    if (!dC || !A || !B) EML_BAD_ARG("matmul_bwd: null operand");
    cudaStream_t st = (cudaStream_t)stream;
    // dA[M x K] += dC[M x N] * B^T[N x K]   (B^T read in place: element (n,k) at B[k*N + n])
    if (dA)
        EML_TRY(contract_dev<T>(ctx, dC, B, dA, 1, M, K, N, Strides3{0, N, 1}, Strides3{0, 1, N}, Strides3{0, K, 1},
                                true, false, st));
    // dB[K x N] += A^T[K x M] * dC[M x N]   (A^T read in place: element (k,m) at A[m*K + k])
    if (dB)
        EML_TRY(contract_dev<T>(ctx, A, dC, dB, 1, K, N, M, Strides3{0, 1, K}, Strides3{0, N, 1}, Strides3{0, N, 1},
                                true, false, st));
    return EML_OK;
}

// covariance on the device: sums -> centre -> GEMM (Xc^T Xc or Xc Xc^T via strides) -> / S
template <class T>
int covariance_dev(DeviceCtx* ctx, const T* X, int64_t rows, int64_t cols, int axis, T* out, cudaStream_t st) {
    if (!X || !out) EML_BAD_ARG("covariance: null operand");
    if (rows < 1 || cols < 1) EML_BAD_ARG("covariance: empty matrix");
    if (axis != 0 && axis != 1) EML_BAD_ARG("covariance: feature_axis must be 0 (rows) or 1 (columns)");
    const int64_t F = axis == 1 ? cols : rows, S = axis == 1 ? rows : cols;
    TempBuf sums, xc;
    EML_TRY(sums.alloc(sizeof(T) * (size_t)F, st));
    EML_TRY(xc.alloc(sizeof(T) * (size_t)(rows * cols), st));
    if (axis == 1)
        EML_TRY(launch_col_sums<T>(X, rows, cols, sums.as<T>(), st));
    else
        EML_TRY(launch_row_sums<T>(X, rows, cols, sums.as<T>(), st));
    EML_TRY(launch_center<T>(X, sums.as<T>(), S, rows, cols, axis, xc.as<T>(), st));
    const T* c = xc.as<T>();
    if (axis == 1)  // A(i, s) = Xc[s*F + i], B(s, j) = Xc[s*F + j]
        EML_TRY(contract_dev<T>(ctx, c, c, out, 1, F, F, S, Strides3{0, 1, F}, Strides3{0, F, 1}, Strides3{0, F, 1}, false,
                                false, st));
    else  // A(i, s) = Xc[i*S + s], B(s, j) = Xc[j*S + s]
        EML_TRY(contract_dev<T>(ctx, c, c, out, 1, F, F, S, Strides3{0, S, 1}, Strides3{0, 1, S}, Strides3{0, F, 1}, false,
                                false, st));
    return launch_unary<T>(EML_OP_DIV_C, (T)S, out, out, (T*)nullptr, F * F, st);
}

template <class T>
int covariance_host(const T* X, int64_t rows, int64_t cols, int64_t ldx, int axis, T* out) {
    DeviceCtx* ctx = nullptr;
    EML_TRY(get_ctx(&ctx));
    if (!X || !out) EML_BAD_ARG("covariance: null operand");
    if (rows < 1 || cols < 1 || ldx < cols) EML_BAD_ARG("covariance: bad shape %lldx%lld (ld %lld)", (long long)rows,
                                                        (long long)cols, (long long)ldx);
    const int64_t F = axis == 1 ? cols : rows;
    std::lock_guard<std::mutex> lock(ctx->mu);
    void *dX, *dOut;
    EML_TRY(stage_buffer(ctx, 0, sizeof(T) * (size_t)(rows * cols), &dX));
    EML_TRY(stage_buffer(ctx, 2, sizeof(T) * (size_t)(F * F), &dOut));
    cudaStream_t st = ctx->stream;
    EML_CUDA(cudaMemcpy2DAsync(dX, cols * sizeof(T), X, ldx * sizeof(T), cols * sizeof(T), rows, cudaMemcpyHostToDevice, st));
    EML_TRY(covariance_dev<T>(ctx, (const T*)dX, rows, cols, axis, (T*)dOut, st));
    EML_CUDA(cudaMemcpyAsync(out, dOut, sizeof(T) * (size_t)(F * F), cudaMemcpyDeviceToHost, st));
    EML_CUDA(cudaStreamSynchronize(st));
    return EML_OK;
}

#define EML_STREAM_OR_CTX(st)                 \
    DeviceCtx* ctx = nullptr;                 \
    EML_TRY(get_ctx(&ctx));                   \
    cudaStream_t st = (cudaStream_t)stream;

}  // namespace
}  // namespace eml

using namespace eml;

extern "C" {

int eml_init(int* n_devices) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    if (n_devices) *n_devices = n;
    if (n == 0) {
        set_error("no CUDA device present: the f32/f64 path has no CPU fallback");
        return EML_ERR_NO_DEVICE;
    }
    return EML_OK;
}
void eml_shutdown(void) { shutdown_all(); }
const char* eml_last_error(void) { return last_error(); }
int eml_version(void) { return 100; }
int64_t eml_kernel_launches(void) { return launches(); }
const char* eml_last_kernel(void) { return last_kernel(); }

int eml_gemm_f32(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                 int64_t ldb, int64_t ldc) {
    return gemm_host<float>(A, B, C, M, N, K, lda, ldb, ldc, false);
}
int eml_gemm_f64(const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                 int64_t ldb, int64_t ldc) {
    return gemm_host<double>(A, B, C, M, N, K, lda, ldb, ldc, false);
}
int eml_gemm_exact_f32(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                       int64_t ldb, int64_t ldc) {
    return gemm_host<float>(A, B, C, M, N, K, lda, ldb, ldc, true);
}
int eml_gemm_exact_f64(const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K,
                       int64_t lda, int64_t ldb, int64_t ldc) {
    return gemm_host<double>(A, B, C, M, N, K, lda, ldb, ldc, true);
}
int eml_contract_f32(const float* A, const float* B, float* C, int64_t batch, int64_t M, int64_t N, int64_t K,
                     const int64_t sA[3], const int64_t sB[3], const int64_t sC[3]) {
    if (!sA || !sB || !sC) EML_BAD_ARG("contract: null strides");
    return contract_host<float>(A, B, C, batch, M, N, K, s3(sA), s3(sB), s3(sC), false);
}
int eml_contract_f64(const double* A, const double* B, double* C, int64_t batch, int64_t M, int64_t N,
                     int64_t K, const int64_t sA[3], const int64_t sB[3], const int64_t sC[3]) {
    if (!sA || !sB || !sC) EML_BAD_ARG("contract: null strides");
    return contract_host<double>(A, B, C, batch, M, N, K, s3(sA), s3(sB), s3(sC), false);
}
int eml_einsum2_f32(const float* A, int64_t ae, const float* B, int64_t be, float* out, int n_out,
                    const int64_t* out_len, const int64_t* ao, const int64_t* bo, int n_sum, const int64_t* sum_len,
                    const int64_t* as, const int64_t* bs) {
    return einsum_host<float>(A, ae, B, be, out, n_out, out_len, ao, bo, n_sum, sum_len, as, bs);
}
int eml_einsum2_f64(const double* A, int64_t ae, const double* B, int64_t be, double* out, int n_out,
                    const int64_t* out_len, const int64_t* ao, const int64_t* bo, int n_sum, const int64_t* sum_len,
                    const int64_t* as, const int64_t* bs) {
    return einsum_host<double>(A, ae, B, be, out, n_out, out_len, ao, bo, n_sum, sum_len, as, bs);
}
int eml_dot_f32(const float* x, int64_t incx, const float* y, int64_t incy, int64_t n, float* out) {
    return dot_host<float>(x, incx, y, incy, n, out);
}
int eml_dot_f64(const double* x, int64_t incx, const double* y, int64_t incy, int64_t n, double* out) {
    return dot_host<double>(x, incx, y, incy, n, out);
}

int eml_alloc(void** dptr, size_t bytes) {
    DeviceCtx* ctx = nullptr;
    EML_TRY(get_ctx(&ctx));
    if (!dptr) EML_BAD_ARG("alloc: null out pointer");
    EML_CUDA(cudaMalloc(dptr, bytes ? bytes : 16));
    return EML_OK;
}
int eml_free(void* dptr) {
    if (dptr) EML_CUDA(cudaFree(dptr));
    return EML_OK;
}
int eml_alloc_async(void** dptr, size_t bytes, void* stream) {
    DeviceCtx* ctx = nullptr;
    EML_TRY(get_ctx(&ctx));
    if (!dptr) EML_BAD_ARG("alloc_async: null out pointer");
    return stream_alloc(dptr, bytes ? bytes : 16, (cudaStream_t)stream);
}
int eml_free_async(void* dptr, size_t bytes, void* stream) {
    if (dptr) stream_free(dptr, bytes ? bytes : 16, (cudaStream_t)stream);
    return EML_OK;
}
int eml_upload(void* dst, const void* src, size_t bytes) {
    DeviceCtx* ctx = nullptr;
    EML_TRY(get_ctx(&ctx));
    EML_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    EML_CUDA(cudaStreamSynchronize(ctx->stream));
    return EML_OK;
}
int eml_download(void* dst, const void* src, size_t bytes) {
    DeviceCtx* ctx = nullptr;
    EML_TRY(get_ctx(&ctx));
    EML_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    EML_CUDA(cudaStreamSynchronize(ctx->stream));
    return EML_OK;
}
int eml_gemm_f64_scatter_dev(const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K,
                             int64_t lda, int64_t ldb, int64_t ldc, int accumulate, double* const* c_remote,
                             int n_remote, void* stream) {
    DeviceCtx* ctx = nullptr;
    EML_TRY(get_ctx(&ctx));
    if (n_remote < 0 || n_remote > 7 || (n_remote && !c_remote)) EML_BAD_ARG("scatter: 0..7 remote destinations");
    if (lda < K || ldb < N || ldc < N) EML_BAD_ARG("gemm: leading dimension smaller than the row length");
    cudaStream_t st = (cudaStream_t)stream;
    GemmOptions opt;
    opt.no_split_k = true;  // the rows must carry the same bits as an unsharded launch
    opt.remote_c = c_remote;
    opt.n_remote = n_remote;
    EML_TRY(contract_dev<double>(ctx, A, B, C, 1, M, N, K, Strides3{0, lda, 1}, Strides3{0, ldb, 1},
                                 Strides3{0, ldc, 1}, accumulate != 0, false, st, &opt));
    const bool pending = n_remote > 0 && !opt.remote_done;
    if (pending)  // a kernel without the fused epilogue ran (unaligned / tiny shape): copy the rows instead
        for (int r = 0; r < n_remote; r++)
            EML_CUDA(cudaMemcpy2DAsync(c_remote[r], ldc * sizeof(double), C, ldc * sizeof(double), N * sizeof(double),
                                       M, cudaMemcpyDefault, st));
    return EML_OK;
}
int eml_ipc_export(void* dptr, void* handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    if (!dptr || !handle64) EML_BAD_ARG("ipc_export: null argument");
    cudaIpcMemHandle_t h;
    EML_CUDA(cudaIpcGetMemHandle(&h, dptr));
    memcpy(handle64, &h, 64);
    return EML_OK;
}
int eml_ipc_open(const void* handle64, void** dptr) {
    if (!dptr || !handle64) EML_BAD_ARG("ipc_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    EML_CUDA(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return EML_OK;
}
int eml_ipc_close(void* dptr) {
    if (dptr) EML_CUDA(cudaIpcCloseMemHandle(dptr));
    return EML_OK;
}
int eml_copy_async(void* dst, const void* src, size_t bytes, void* stream) {
    DeviceCtx* ctx = nullptr;
    EML_TRY(get_ctx(&ctx));
    if (!dst || !src) EML_BAD_ARG("copy_async: null pointer");
    EML_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return EML_OK;
}
int eml_einsum_n_f32(int n_ops, const float* const* x, const int64_t* elems, float* out, int n_out,
                     const int64_t* out_len, const int64_t* out_strides, int n_sum, const int64_t* sum_len,
                     const int64_t* sum_strides) {
    return einsum_n_host<float>(n_ops, x, elems, out, n_out, out_len, out_strides, n_sum, sum_len, sum_strides);
}
int eml_einsum_n_f64(int n_ops, const double* const* x, const int64_t* elems, double* out, int n_out,
                     const int64_t* out_len, const int64_t* out_strides, int n_sum, const int64_t* sum_len,
                     const int64_t* sum_strides) {
    return einsum_n_host<double>(n_ops, x, elems, out, n_out, out_len, out_strides, n_sum, sum_len, sum_strides);
}
int eml_covariance_f32(const float* X, int64_t rows, int64_t cols, int64_t ldx, int feature_axis, float* out) {
    return covariance_host<float>(X, rows, cols, ldx, feature_axis, out);
}
int eml_covariance_f64(const double* X, int64_t rows, int64_t cols, int64_t ldx, int feature_axis, double* out) {
    return covariance_host<double>(X, rows, cols, ldx, feature_axis, out);
}
int eml_covariance_f32_dev(const float* X, int64_t rows, int64_t cols, int feature_axis, float* out, void* stream) {
    DeviceCtx* ctx = nullptr;
    EML_TRY(get_ctx(&ctx));
    return covariance_dev<float>(ctx, X, rows, cols, feature_axis, out, (cudaStream_t)stream);
}
int eml_covariance_f64_dev(const double* X, int64_t rows, int64_t cols, int feature_axis, double* out, void* stream) {
    DeviceCtx* ctx = nullptr;
    EML_TRY(get_ctx(&ctx));
    return covariance_dev<double>(ctx, X, rows, cols, feature_axis, out, (cudaStream_t)stream);
}
int eml_measure_fp64_tensor_peak(double* tflops) {
    DeviceCtx* ctx = nullptr;
    EML_TRY(get_ctx(&ctx));
    if (!tflops) EML_BAD_ARG("null out pointer");
    return measure_dmma_peak(ctx, tflops);
}
int eml_host_alloc(void** hptr, size_t bytes) {
    DeviceCtx* ctx = nullptr;
    EML_TRY(get_ctx(&ctx));
    if (!hptr) EML_BAD_ARG("host_alloc: null out pointer");
    EML_CUDA(cudaHostAlloc(hptr, bytes ? bytes : 16, cudaHostAllocDefault));
    return EML_OK;
}
int eml_host_free(void* hptr) {
    if (hptr) EML_CUDA(cudaFreeHost(hptr));
    return EML_OK;
}
int eml_sync(void) {
    DeviceCtx* ctx = nullptr;
    EML_TRY(get_ctx(&ctx));
    EML_CUDA(cudaDeviceSynchronize());
    trace_dump();
    return EML_OK;
}

int eml_contract_f32_dev(const float* A, const float* B, float* C, int64_t batch, int64_t M, int64_t N,
                         int64_t K, const int64_t sA[3], const int64_t sB[3], const int64_t sC[3],
                         int accumulate, void* stream) {
    DeviceCtx* ctx = nullptr;
    EML_TRY(get_ctx(&ctx));
    if (!sA || !sB || !sC) EML_BAD_ARG("contract: null strides");
    return contract_dev<float>(ctx, A, B, C, batch, M, N, K, s3(sA), s3(sB), s3(sC), accumulate != 0, false,
                               (cudaStream_t)stream);
}
int eml_contract_f64_dev(const double* A, const double* B, double* C, int64_t batch, int64_t M, int64_t N,
                         int64_t K, const int64_t sA[3], const int64_t sB[3], const int64_t sC[3],
                         int accumulate, void* stream) {
    DeviceCtx* ctx = nullptr;
    EML_TRY(get_ctx(&ctx));
    if (!sA || !sB || !sC) EML_BAD_ARG("contract: null strides");
    return contract_dev<double>(ctx, A, B, C, batch, M, N, K, s3(sA), s3(sB), s3(sC), accumulate != 0, false,
                                (cudaStream_t)stream);
}
int eml_gemm_f32_dev(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                     int64_t ldb, int64_t ldc, void* stream) {
    return gemm_dev<float>(A, B, C, M, N, K, lda, ldb, ldc, false, stream);
}
int eml_gemm_f64_dev(const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                     int64_t ldb, int64_t ldc, void* stream) {
    return gemm_dev<double>(A, B, C, M, N, K, lda, ldb, ldc, false, stream);
}
int eml_gemm_exact_f32_dev(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K,
                           int64_t lda, int64_t ldb, int64_t ldc, void* stream) {
    return gemm_dev<float>(A, B, C, M, N, K, lda, ldb, ldc, true, stream);
}
int eml_gemm_exact_f64_dev(const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K,
                           int64_t lda, int64_t ldb, int64_t ldc, void* stream) {
    return gemm_dev<double>(A, B, C, M, N, K, lda, ldb, ldc, true, stream);
}
int eml_matmul_bwd_f32_dev(const float* dC, const float* A, const float* B, float* dA, float* dB, int64_t M,
                           int64_t N, int64_t K, void* stream) {
    return matmul_bwd_dev<float>(dC, A, B, dA, dB, M, N, K, stream);
}
int eml_matmul_bwd_f64_dev(const double* dC, const double* A, const double* B, double* dA, double* dB,
                           int64_t M, int64_t N, int64_t K, void* stream) {
    return matmul_bwd_dev<double>(dC, A, B, dA, dB, M, N, K, stream);
}

int eml_unary_f32_dev(int op, float c, const float* x, float* y, float* dydx, int64_t n, void* stream) {
    EML_STREAM_OR_CTX(st);
    return launch_unary<float>(op, c, x, y, dydx, n, st);
}
int eml_unary_f64_dev(int op, double c, const double* x, double* y, double* dydx, int64_t n, void* stream) {
    EML_STREAM_OR_CTX(st);
    return launch_unary<double>(op, c, x, y, dydx, n, st);
}
int eml_binary_f32_dev(int op, const float* x, const float* y, float* z, float* dzdx, float* dzdy, int64_t n,
                       void* stream) {
    EML_STREAM_OR_CTX(st);
    return launch_binary<float>(op, x, y, z, dzdx, dzdy, n, st);
}
int eml_binary_f64_dev(int op, const double* x, const double* y, double* z, double* dzdx, double* dzdy,
                       int64_t n, void* stream) {
    EML_STREAM_OR_CTX(st);
    return launch_binary<double>(op, x, y, z, dzdx, dzdy, n, st);
}
int eml_chain_f32_dev(const float* dout, const float* deriv, float* dparent, int64_t n, void* stream) {
    EML_STREAM_OR_CTX(st);
    return launch_chain<float>(dout, deriv, dparent, n, st);
}
int eml_chain_f64_dev(const double* dout, const double* deriv, double* dparent, int64_t n, void* stream) {
    EML_STREAM_OR_CTX(st);
    return launch_chain<double>(dout, deriv, dparent, n, st);
}
int eml_sgd_f32_dev(float* w, const float* d, float lr, int64_t n, void* stream) {
    EML_STREAM_OR_CTX(st);
    return launch_sgd<float>(w, d, lr, n, st);
}
int eml_sgd_f64_dev(double* w, const double* d, double lr, int64_t n, void* stream) {
    EML_STREAM_OR_CTX(st);
    return launch_sgd<double>(w, d, lr, n, st);
}
int eml_gemm_f64_sliced_dev(const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                            int64_t ldb, int64_t ldc, int slices, int* path, void* stream) {
    EML_STREAM_OR_CTX(st);
    if (!A || !B || !C) EML_BAD_ARG("sliced gemm: null operand");
    return gemm_f64_sliced_dev(ctx, A, B, C, M, N, K, lda, ldb, ldc, slices, path, st);
}
int eml_reorder_f32_dev(const float* in, float* out, int rank, const int64_t* out_lens, const int64_t* in_strides,
                        void* stream) {
    EML_STREAM_OR_CTX(st);
    return launch_reorder<float>(in, out, rank, out_lens, in_strides, st);
}
int eml_reorder_f64_dev(const double* in, double* out, int rank, const int64_t* out_lens, const int64_t* in_strides,
                        void* stream) {
    EML_STREAM_OR_CTX(st);
    return launch_reorder<double>(in, out, rank, out_lens, in_strides, st);
}

}  // extern "C"
