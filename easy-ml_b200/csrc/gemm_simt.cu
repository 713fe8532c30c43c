// gemm_simt.cu -- general strided batched GEMM on the FMA pipes.
//
// Role: the any-layout, any-shape kernel (ragged edges, odd strides, skinny outputs such as N = 10)
// and the bit-exact mode.  The large contiguous cases go to the tensor-core kernels
// (gemm_dmma_f64.cu, gemm_tf32x3.cu); this one is never the benchmark's dominant kernel.
//
// EXACT = the reference's operation order for one output element (scalar_product,
// reference src/tensors/operations.rs:441-445): first product seeds the sum, then k ascending with a
// separately rounded multiply and add -- no FMA.  k tiles are walked in ascending order, so the chain
// per element is exactly the sequential one.
#include <type_traits>

#include "common.h"

namespace eml {

namespace {

template <class T>
__device__ __forceinline__ T mul_rn(T a, T b);
template <>
__device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }
template <>
__device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <class T>
__device__ __forceinline__ T add_rn(T a, T b);
template <>
__device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }
template <>
__device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4, THREADS = 256;

template <class T, bool EXACT>
__global__ void __launch_bounds__(THREADS)
gemm_simt_kernel(const T* __restrict__ A, const T* __restrict__ B, T* C, int64_t batch, int64_t M,
                 int64_t N, int64_t K, Strides3 sA, Strides3 sB, Strides3 sC, int accumulate, int64_t tiles_n) {
    pdl_prologue();
    __shared__ T As[BK][BM + PAD];
    __shared__ T Bs[BK][BN + PAD];
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int64_t tile = blockIdx.x;
    const int64_t m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
    const bool a_kfast = (sA.c == 1), b_nfast = (sB.c == 1);

    for (int64_t b = blockIdx.y; b < batch; b += gridDim.y) {
        const T* Ab = A + b * sA.b;
        const T* Bb = B + b * sB.b;
        T* Cb = C + b * sC.b;
        // accumulate: start the chain from C (a K-split sequence of launches == one launch, bit for bit)
        T acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                int64_t gm = m0 + ty * 4 + i, gn = n0 + tx * 4 + j;
                acc[i][j] = (accumulate && gm < M && gn < N) ? Cb[gm * sC.r + gn * sC.c] : T(0);
            }

        for (int64_t k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
            for (int i = 0; i < (BM * BK) / THREADS; i++) {
                int e = tid + i * THREADS;
                int m = a_kfast ? e / BK : e % BM;
                int k = a_kfast ? e % BK : e / BM;
                int64_t gm = m0 + m, gk = k0 + k;
                As[k][m] = (gm < M && gk < K) ? Ab[gm * sA.r + gk * sA.c] : T(0);
            }
#pragma unroll
            for (int i = 0; i < (BN * BK) / THREADS; i++) {
                int e = tid + i * THREADS;
                int n = b_nfast ? e % BN : e / BK;
                int k = b_nfast ? e / BN : e % BK;
                int64_t gn = n0 + n, gk = k0 + k;
                Bs[k][n] = (gn < N && gk < K) ? Bb[gk * sB.r + gn * sB.c] : T(0);
            }
            __syncthreads();
            const int kmax = (K - k0) < BK ? int(K - k0) : BK;
            if (kmax == BK) {
#pragma unroll
                for (int k = 0; k < BK; k++) {
                    T a[4], bb[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) a[i] = As[k][ty * 4 + i];
#pragma unroll
                    for (int j = 0; j < 4; j++) bb[j] = Bs[k][tx * 4 + j];
#pragma unroll
                    for (int i = 0; i < 4; i++)
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            if (EXACT) {
                                T p = mul_rn<T>(a[i], bb[j]);
                                acc[i][j] = (k0 == 0 && k == 0) ? p : add_rn<T>(acc[i][j], p);
                            } else {
                                acc[i][j] = fma(a[i], bb[j], acc[i][j]);
                            }
                        }
                }
            } else {
                for (int k = 0; k < kmax; k++) {
                    T a[4], bb[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) a[i] = As[k][ty * 4 + i];
#pragma unroll
                    for (int j = 0; j < 4; j++) bb[j] = Bs[k][tx * 4 + j];
#pragma unroll
                    for (int i = 0; i < 4; i++)
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            if (EXACT) {
                                T p = mul_rn<T>(a[i], bb[j]);
                                acc[i][j] = (k0 == 0 && k == 0) ? p : add_rn<T>(acc[i][j], p);
                            } else {
                                acc[i][j] = fma(a[i], bb[j], acc[i][j]);
                            }
                        }
                }
            }
            __syncthreads();
        }

#pragma unroll
        for (int i = 0; i < 4; i++) {
            int64_t gm = m0 + ty * 4 + i;
            if (gm >= M) continue;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                int64_t gn = n0 + tx * 4 + j;
                if (gn >= N) continue;
                Cb[gm * sC.r + gn * sC.c] = acc[i][j];
            }
        }
    }
}

}  // namespace

template <class T>
int launch_gemm_simt(const T* A, const T* B, T* C, int64_t batch, int64_t M, int64_t N, int64_t K, Strides3 sA,
                     Strides3 sB, Strides3 sC, bool accumulate, bool exact, cudaStream_t stream) {
    if (batch == 0 || M == 0 || N == 0) return EML_OK;
    int64_t tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
    if (tiles_m * tiles_n > 0x7fffffffLL) EML_BAD_ARG("gemm_simt: too many tiles");
    dim3 grid((unsigned)(tiles_m * tiles_n), (unsigned)(batch < 65535 ? batch : 65535));
    if (exact)
        launch_k(gemm_simt_kernel<T, true>, dim3(grid), dim3(THREADS), 0, stream, A, B, C, batch, M, N, K, sA, sB, sC, 0, tiles_n);
    else
        launch_k(gemm_simt_kernel<T, false>, dim3(grid), dim3(THREADS), 0, stream, A, B, C, batch, M, N, K, sA, sB, sC, accumulate ? 1 : 0, tiles_n);
    count_launch();
    set_last_kernel(exact ? "simt_exact" : "simt");
    return check_launch("gemm_simt");
}

// ------------------------------------------------------------------------------------------------
// Skinny-N GEMM: C[M x N] (+)= A[M x K] * B[K x N] with N <= 16 and A contiguous along k (the NN's
// H[8192 x 512] * W2[512 x 10]).  HBM-bound on A: every warp streams whole rows of A with coalesced loads,
// B sits in shared memory (k chunks, [n][k] so lanes hit distinct banks), each lane keeps N partial sums
// over its k's and a fixed shuffle tree combines the 32 lanes -- deterministic, no atomics.
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int SK_NMAX = 16, SK_WARPS = 8, SK_THREADS = SK_WARPS * 32;

// Sums V values per lane over the 32 lanes of a warp in the fixed tree of the xor butterfly (partners 16, 8, 4, 2, 1 apart,
// in that order), but with a fraction of its shuffles: while the count is even, a lane KEEPS half of its values and
// sends the other half to its partner (which keeps those), so a level costs V / 2 shuffles instead of V and the count
// halves; the levels that are left run the plain butterfly on what remains.  Every total is the same sequence of adds
// over the same partial sums as  v += shfl_xor(v, 16); v += shfl_xor(v, 8); ...  (an add's two operands may be
// exchanged between the partners: IEEE addition is commutative), so the bits are those of the plain butterfly.
// The shuffle unit moves one warp instruction per clock and SM; the plain butterfly of the skinny product's 4 x N sums
// (200 shuffles per warp for N = 10) took 3 of the kernel's 7 us (clock64 timeline), this form needs 45.
// On return the lane holds `count` totals, of the values base .. base + count - 1, and `writer` says whether this lane is
// the one (of those that hold the same totals) that stores them.
template <class T, int V, int OFF>
struct WarpTree {
    static __device__ __forceinline__ void run(T (&v)[V], int lane, int& base, bool& writer, T* out, int& count) {
        if constexpr (OFF == 0) {
#pragma unroll
            for (int i = 0; i < V; i++) out[i] = v[i];
            count = V;
        } else if constexpr (V % 2 == 0) {
            const bool up = (lane & OFF) != 0;
            T w[V / 2];
#pragma unroll
            for (int i = 0; i < V / 2; i++) {
                const T keep = up ? v[i + V / 2] : v[i];
                const T send = up ? v[i] : v[i + V / 2];
                w[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
            }
            if (up) base += V / 2;
            WarpTree<T, V / 2, OFF / 2>::run(w, lane, base, writer, out, count);
        } else {
#pragma unroll
            for (int i = 0; i < V; i++) v[i] += __shfl_xor_sync(0xffffffffu, v[i], OFF);
            writer = writer && (lane & OFF) == 0;
            WarpTree<T, V, OFF / 2>::run(v, lane, base, writer, out, count);
        }
    }
};



template <class T>
__global__ void __launch_bounds__(SK_THREADS)
gemm_skinny_kernel(const T* __restrict__ A, const T* __restrict__ B, T* C, int64_t batch, int64_t M, int N, int64_t K,
                   Strides3 sA, Strides3 sB, Strides3 sC, int accumulate) {
    pdl_prologue();
    constexpr int KC = 2048 / sizeof(T);  // k chunk: 512 (f32) / 256 (f64) -> 32 KB of shared memory
    constexpr int PER_LANE = KC / 32;     // loads one lane keeps in flight per chunk
    __shared__ T sBk[SK_NMAX][KC];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row_blocks = (M + SK_WARPS - 1) / SK_WARPS;
    // K fits one chunk and there is one B (the NN's H * W2): B is staged once per CTA, not once per row group
    const bool stage_once = K <= KC && batch == 1;
    auto stage = [&](const T* Bb, int64_t k0, int kc) {
        // one k per thread and pass: its N loads are independent (all in flight together) and the stores are
        // conflict free (consecutive threads, consecutive k)
        // (the tail [kc, KC) is zero filled: the compute loop multiplies it by zero-padded A values, and stale
        //  shared memory could hold NaN / Inf patterns)
        for (int k = threadIdx.x; k < KC; k += SK_THREADS) {
            T v[SK_NMAX];
#pragma unroll
            for (int n = 0; n < SK_NMAX; n++) v[n] = (n < N && k < kc) ? Bb[(k0 + k) * sB.r + n * sB.c] : T(0);
#pragma unroll
            for (int n = 0; n < SK_NMAX; n++)
                if (n < N) sBk[n][k] = v[n];
        }
    };
    if (stage_once) {
        stage(B, 0, (int)K);
        __syncthreads();
    }
    for (int64_t blk = blockIdx.x; blk < row_blocks * batch; blk += gridDim.x) {
        const int64_t b = blk / row_blocks;
        const int64_t m = (blk % row_blocks) * SK_WARPS + warp;  // one row per warp
        const T* Ab = A + b * sA.b;
        const T* Bb = B + b * sB.b;
        T acc[SK_NMAX];
#pragma unroll
        for (int n = 0; n < SK_NMAX; n++) acc[n] = T(0);
        for (int64_t k0 = 0; k0 < K; k0 += KC) {
            const int kc = (int)((K - k0) < KC ? (K - k0) : KC);
            if (!stage_once) {
                __syncthreads();  // the previous chunk has been consumed
                stage(Bb, k0, kc);
                __syncthreads();
            }
            if (m < M) {
                // all of the lane's loads of this chunk are issued before the first use: PER_LANE coalesced 128-byte
                // warp loads in flight per warp (one dependent load per 32 k is latency-bound)
                T av[PER_LANE];
                const T* a = Ab + m * sA.r + k0;
#pragma unroll
                for (int j = 0; j < PER_LANE; j++) av[j] = (lane + 32 * j < kc) ? a[lane + 32 * j] : T(0);
#pragma unroll
                for (int n = 0; n < SK_NMAX; n++)
                    if (n < N) {
#pragma unroll
                        for (int j = 0; j < PER_LANE; j++) acc[n] = fma(av[j], sBk[n][lane + 32 * j], acc[n]);
                    }
            }
        }
        {
            // (every lane of the warp takes part, also for a row past the end: its sums are zeros and nothing is stored)
            T tot[SK_NMAX];
            int base = 0, count = 0;
            bool writer = true;
            WarpTree<T, SK_NMAX, 16>::run(acc, lane, base, writer, tot, count);
            if (writer && m < M) {
#pragma unroll
                for (int i = 0; i < SK_NMAX; i++) {
                    if (i >= count) break;
                    const int n = base + i;
                    if (n < N) {
                        T* p = C + b * sC.b + m * sC.r + n * sC.c;
                        *p = accumulate ? *p + tot[i] : tot[i];
                    }
                }
            }
        }
    }
}
}  // namespace

// ------------------------------------------------------------------------------------------------
// Skinny-N GEMM with A contiguous along m (the NN's dW2 = H^T[512 x batch] * dY[batch x 10]): each thread owns one
// output row m and N accumulators, a CTA walks one k chunk with coalesced loads of A[k][m..] (B's chunk sits in
// shared memory), and the per-chunk partials are added in chunk order by the split-K reduction: deterministic.
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int TN_THREADS = 256, TN_KC = 64;  // k rows of B staged per pass

template <class T>
__global__ void __launch_bounds__(TN_THREADS)
gemm_skinny_tn_kernel(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ ws, int64_t M, int N, int64_t K,
                      Strides3 sA, Strides3 sB, int64_t k_per_chunk) {
    pdl_prologue();
    __shared__ T sB_[TN_KC][SK_NMAX];
    const int64_t m = (int64_t)blockIdx.x * TN_THREADS + threadIdx.x;
    const int64_t kc0 = (int64_t)blockIdx.y * k_per_chunk, kc1 = kc0 + k_per_chunk < K ? kc0 + k_per_chunk : K;
    T acc[SK_NMAX];
#pragma unroll
    for (int n = 0; n < SK_NMAX; n++) acc[n] = T(0);
    for (int64_t k0 = kc0; k0 < kc1; k0 += TN_KC) {
        const int kc = (int)(kc1 - k0 < TN_KC ? kc1 - k0 : TN_KC);
        __syncthreads();
        for (int e = threadIdx.x; e < kc * N; e += TN_THREADS) {
            int k = e / N, n = e % N;
            sB_[k][n] = B[(k0 + k) * sB.r + n * sB.c];
        }
        __syncthreads();
        if (m < M) {
            const T* a = A + m * sA.r + k0 * sA.c;
            int k = 0;
            for (; k + 16 <= kc; k += 16) {  // sixteen independent (coalesced across the CTA) loads in flight
                T av[16];
#pragma unroll
                for (int j = 0; j < 16; j++) av[j] = a[(k + j) * sA.c];
#pragma unroll
                for (int n = 0; n < SK_NMAX; n++)
                    if (n < N) {
#pragma unroll
                        for (int j = 0; j < 16; j++) acc[n] = fma(av[j], sB_[k + j][n], acc[n]);
                    }
            }
            for (; k < kc; k++) {
                T a0 = a[k * sA.c];
#pragma unroll
                for (int n = 0; n < SK_NMAX; n++)
                    if (n < N) acc[n] = fma(a0, sB_[k][n], acc[n]);
            }
        }
    }
    if (m < M) {
        T* out = ws + ((int64_t)blockIdx.y * M + m) * N;
#pragma unroll
        for (int n = 0; n < SK_NMAX; n++)
            if (n < N) out[n] = acc[n];
    }
}
}  // namespace

// The same product when one CTA can own EVERY output row (M <= 256 * TNW_MT): a CTA walks one k range, each thread
// keeps TNW_MT rows x N accumulators, so a k row of A (M contiguous elements) is read once, by one CTA, with all of a
// thread's loads of eight k rows in flight, and B's chunk is staged once per CTA (no barrier inside the k loop).
namespace {
constexpr int TNW_MT = 4, TNW_KC = 128;

template <class T, int MT>
__global__ void __launch_bounds__(TN_THREADS)
gemm_tn_wide_kernel(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ ws, int64_t M, int N, int64_t K,
                    int64_t a_k_stride, Strides3 sB, int64_t k_per_block) {
    pdl_prologue();
    __shared__ T sB_[TNW_KC][SK_NMAX];
    const int64_t kb0 = (int64_t)blockIdx.x * k_per_block, kb1 = kb0 + k_per_block < K ? kb0 + k_per_block : K;
    T acc[MT][SK_NMAX];
#pragma unroll
    for (int i = 0; i < MT; i++)
#pragma unroll
        for (int n = 0; n < SK_NMAX; n++) acc[i][n] = T(0);
    for (int64_t k0 = kb0; k0 < kb1; k0 += TNW_KC) {
        const int kc = (int)(kb1 - k0 < TNW_KC ? kb1 - k0 : TNW_KC);
        __syncthreads();
        for (int e = threadIdx.x; e < kc * N; e += TN_THREADS) {
            const int k = e / N, n = e % N;
            sB_[k][n] = B[(k0 + k) * sB.r + n * sB.c];
        }
        __syncthreads();
        const T* a = A + k0 * a_k_stride + threadIdx.x;
        // eight k rows in flight per thread (rows past the chunk read as zero), and the next eight are requested before
        // the fma of the current ones: the chunk's memory latency overlaps its arithmetic
        T nx[8][MT];
#pragma unroll
        for (int j = 0; j < 8; j++)
#pragma unroll
            for (int i = 0; i < MT; i++)
                nx[j][i] = (j < kc && threadIdx.x + i * TN_THREADS < M) ? a[j * a_k_stride + i * TN_THREADS] : T(0);
        for (int k = 0; k < kc; k += 8) {
            T av[8][MT];
#pragma unroll
            for (int j = 0; j < 8; j++)
#pragma unroll
                for (int i = 0; i < MT; i++) av[j][i] = nx[j][i];
            if (k + 8 < kc) {
#pragma unroll
                for (int j = 0; j < 8; j++)
#pragma unroll
                    for (int i = 0; i < MT; i++)
                        nx[j][i] = (k + 8 + j < kc && threadIdx.x + i * TN_THREADS < M) ? a[(k + 8 + j) * a_k_stride + i * TN_THREADS] : T(0);
            }
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (k + j >= kc) break;
#pragma unroll
                for (int n = 0; n < SK_NMAX; n++)
                    if (n < N) {
                        const T b = sB_[k + j][n];
#pragma unroll
                        for (int i = 0; i < MT; i++) acc[i][n] = fma(av[j][i], b, acc[i][n]);
                    }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < MT; i++) {
        const int64_t m = threadIdx.x + i * TN_THREADS;
        if (m < M) {
            T* out = ws + ((int64_t)blockIdx.x * M + m) * N;
#pragma unroll
            for (int n = 0; n < SK_NMAX; n++)
                if (n < N) out[n] = acc[i][n];
        }
    }
}
}  // namespace

template <class T>
int launch_gemm_skinny_tn(DeviceCtx* ctx, const T* A, const T* B, T* C, int64_t M, int64_t N, int64_t K, Strides3 sA,
                          Strides3 sB, Strides3 sC, int accumulate, cudaStream_t stream) {
    if (M <= (int64_t)TN_THREADS * TNW_MT) {
        const int64_t sms = ctx ? ctx->sm_count : 148;
        int64_t blocks = 2 * sms;  // two CTAs per SM (latency hiding); the partials (blocks x M x N) stay small
        if (blocks > (K + 15) / 16) blocks = (K + 15) / 16;
        if (blocks < 1) blocks = 1;
        const int64_t k_per_block = (K + blocks - 1) / blocks;
        blocks = (K + k_per_block - 1) / k_per_block;
        TempBuf ws;
        EML_TRY(ws.alloc(sizeof(T) * (size_t)(blocks * M * N), stream));
        const int mt = (int)((M + TN_THREADS - 1) / TN_THREADS);
        auto go = [&](auto kernel) {
            launch_k(kernel, dim3((unsigned)blocks), dim3(TN_THREADS), 0, stream, A, B, ws.as<T>(), M, (int)N, K, sA.c, sB, k_per_block);
        };
        if (mt == 1) go(gemm_tn_wide_kernel<T, 1>);
        else if (mt == 2) go(gemm_tn_wide_kernel<T, 2>);
        else go(gemm_tn_wide_kernel<T, 4>);
        count_launch();
        set_last_kernel("skinny_tn");
        EML_TRY(check_launch("gemm_tn_wide"));
        return launch_splitk_reduce<T>(ws.as<T>(), (int)blocks, C, 1, M, N, sC, accumulate, stream);
    }
    const int64_t m_blocks = (M + TN_THREADS - 1) / TN_THREADS;
    const int64_t sms = ctx ? ctx->sm_count : 148;
    int64_t chunks = (2 * sms + m_blocks - 1) / m_blocks;  // ~2 CTAs per SM
    if (chunks > (K + TN_KC - 1) / TN_KC) chunks = (K + TN_KC - 1) / TN_KC;
    if (chunks < 1) chunks = 1;
    const int64_t k_per_chunk = ((K + chunks - 1) / chunks + TN_KC - 1) / TN_KC * TN_KC;
    chunks = (K + k_per_chunk - 1) / k_per_chunk;
    TempBuf ws;
    EML_TRY(ws.alloc(sizeof(T) * (size_t)(chunks * M * N), stream));
    dim3 grid((unsigned)m_blocks, (unsigned)chunks);
    launch_k(gemm_skinny_tn_kernel<T>, dim3(grid), dim3(TN_THREADS), 0, stream, A, B, ws.as<T>(), M, (int)N, K, sA, sB, k_per_chunk);
    count_launch();
    set_last_kernel("skinny_tn");
    EML_TRY(check_launch("gemm_skinny_tn"));
    return launch_splitk_reduce<T>(ws.as<T>(), (int)chunks, C, 1, M, N, sC, accumulate, stream);
}
template int launch_gemm_skinny_tn<float>(DeviceCtx*, const float*, const float*, float*, int64_t, int64_t, int64_t,
                                          Strides3, Strides3, Strides3, int, cudaStream_t);
template int launch_gemm_skinny_tn<double>(DeviceCtx*, const double*, const double*, double*, int64_t, int64_t, int64_t,
                                           Strides3, Strides3, Strides3, int, cudaStream_t);

// The common case of the skinny product -- one dense row-major B[K x N] small enough to live in L1, K a multiple of
// 32 lanes x one 16-byte vector (an output layer: H[8192 x 512] * W2[512 x 10]).  A lane owns VEC consecutive k of
// every 32 * VEC; the VEC rows of B it needs are VEC * N CONSECUTIVE numbers of the row-major B, i.e. N 16-byte loads
// (served by L1: every warp reads the same 20 KB), kept in registers while the warp's ROWS rows of A stream by -- no
// shared memory, no staging prologue, no index arithmetic in the loop.  N is a template parameter so that the
// accumulators and B's slice are registers.  One fixed shuffle tree per output; deterministic.
namespace {
constexpr int SKR_ROWS = 4;

// STAGED: B is first copied into shared memory.  A lane's slice of B (VEC rows = N 16-byte units) starts N units after its
// neighbour's: read from global memory that is 32 different lines per load instruction (N x 32 wavefronts of the L1 per
// step -- what bounded the first version: 9.7 us for H[8192 x 512] * W2[512 x 10]); in shared memory, with every slice
// padded to an odd number of units, the same loads are conflict-free.  Persistent blocks stage B once.  Same operations in
// the same order either way.
template <class T, int N, bool STAGED>
__global__ void __launch_bounds__(SK_THREADS)
gemm_skinny_rows_kernel(const T* __restrict__ A, const T* __restrict__ B, T* C, int64_t M, int K, int64_t lda, Strides3 sC,
                        int accumulate) {
    pdl_prologue();
    constexpr int VEC = 16 / sizeof(T);
    constexpr int NP = N | 1;  // units per slice in shared memory
    using V4 = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    extern __shared__ __align__(16) unsigned char skr_smem[];
    V4* sB = reinterpret_cast<V4*>(skr_smem);
    if constexpr (STAGED) {
        const int units = (K / VEC) * N;
        for (int u = threadIdx.x; u < units; u += SK_THREADS) sB[(u / N) * NP + u % N] = __ldg(reinterpret_cast<const V4*>(B) + u);
        __syncthreads();
    }
    const int steps = K / (32 * VEC);
    const int64_t groups = (M + SKR_ROWS - 1) / SKR_ROWS;
    // groups go round the BLOCKS first: when there are fewer groups than warps (8192 rows: 2048 groups, 2368 warps) every
    // block then has the same number of busy warps, instead of 256 full blocks and 40 idle ones on 148 SMs
    for (int64_t grp = (int64_t)warp * gridDim.x + blockIdx.x; grp < groups; grp += (int64_t)gridDim.x * SK_WARPS) {
        const int64_t m0 = grp * SKR_ROWS;
        T acc[SKR_ROWS][N];
#pragma unroll
        for (int r = 0; r < SKR_ROWS; r++)
#pragma unroll
            for (int n = 0; n < N; n++) acc[r][n] = T(0);
        for (int s = 0; s < steps; s++) {
            const int k0 = (s * 32 + lane) * VEC;
            V4 av[SKR_ROWS];
#pragma unroll
            for (int r = 0; r < SKR_ROWS; r++)
                if (m0 + r < M) av[r] = *reinterpret_cast<const V4*>(A + (m0 + r) * lda + k0);
            V4 bv[N];  // B[k0 .. k0 + VEC)[0 .. N): VEC * N consecutive numbers
            if constexpr (STAGED) {
                const V4* bp = sB + (s * 32 + lane) * NP;
#pragma unroll
                for (int q = 0; q < N; q++) bv[q] = bp[q];
            } else {
                const V4* bp = reinterpret_cast<const V4*>(B + (int64_t)k0 * N);
#pragma unroll
                for (int q = 0; q < N; q++) bv[q] = __ldg(bp + q);
            }
            const T* be = reinterpret_cast<const T*>(bv);
#pragma unroll
            for (int r = 0; r < SKR_ROWS; r++) {
                if (m0 + r >= M) break;
                const T* ae = reinterpret_cast<const T*>(&av[r]);
#pragma unroll
                for (int v = 0; v < VEC; v++)
#pragma unroll
                    for (int n = 0; n < N; n++) acc[r][n] = fma(ae[v], be[v * N + n], acc[r][n]);
            }
        }
        {
            constexpr int V = SKR_ROWS * N;
            T flat[V], tot[V];
#pragma unroll
            for (int r = 0; r < SKR_ROWS; r++)
#pragma unroll
                for (int n = 0; n < N; n++) flat[r * N + n] = acc[r][n];
            int base = 0, count = 0;
            bool writer = true;
            WarpTree<T, V, 16>::run(flat, lane, base, writer, tot, count);
            if (writer) {
#pragma unroll
                for (int i = 0; i < V; i++) {
                    if (i >= count) break;
                    const int e = base + i, r = e / N, n = e % N;
                    if (m0 + r < M) {
                        T* p = C + (m0 + r) * sC.r + n * sC.c;
                        *p = accumulate ? *p + tot[i] : tot[i];
                    }
                }
            }
        }
    }
}

template <class T, int N>
int launch_skinny_rows(const T* A, const T* B, T* C, int64_t M, int64_t K, int64_t lda, Strides3 sC, bool accumulate,
                       cudaStream_t stream) {
    const int64_t groups = (M + SKR_ROWS - 1) / SKR_ROWS;
    int64_t blocks = (groups + SK_WARPS - 1) / SK_WARPS;
    constexpr int VEC = 16 / sizeof(T);
    const size_t staged_bytes = (size_t)(K / VEC) * (N | 1) * 16;
    static const bool staging = [] {
        const char* e = getenv("EML_SKINNY_STAGED");  // =0: B through the L1 (comparison path, same bits)
        return !(e && e[0] == '0');
    }();
    if (staging && staged_bytes <= (size_t(48) << 10) && blocks >= 8) {
        // persistent: B is staged once per block; two blocks per SM once there is a group for every SM
        blocks = groups >= 148 * 2 ? 148 * 2 : (groups >= 148 ? 148 : groups);
        launch_k(gemm_skinny_rows_kernel<T, N, true>, dim3((unsigned)blocks), dim3(SK_THREADS), staged_bytes, stream, A, B, C, M, (int)K,
                 lda, sC, accumulate ? 1 : 0);
        count_launch();
        set_last_kernel("skinny_n");
        return check_launch("gemm_skinny_rows");
    }
    if (blocks > 148 * 4) blocks = 148 * 4;
    launch_k(gemm_skinny_rows_kernel<T, N, false>, dim3((unsigned)blocks), dim3(SK_THREADS), 0, stream, A, B, C, M, (int)K, lda, sC,
             accumulate ? 1 : 0);
    count_launch();
    set_last_kernel("skinny_n");
    return check_launch("gemm_skinny_rows");
}
}  // namespace

template <class T>
int launch_gemm_skinny(const T* A, const T* B, T* C, int64_t batch, int64_t M, int64_t N, int64_t K, Strides3 sA,
                       Strides3 sB, Strides3 sC, bool accumulate, cudaStream_t stream) {
    constexpr int64_t VEC = 16 / sizeof(T);
    if (batch == 1 && sA.c == 1 && sB.c == 1 && sB.r == N && K % (32 * VEC) == 0 && sA.r % VEC == 0 &&
        (reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0 &&
        K * N * (int64_t)sizeof(T) <= (int64_t(64) << 10)) {
        switch (N) {
#define EML_SKR(NN) \
    case NN:        \
        return launch_skinny_rows<T, NN>(A, B, C, M, K, sA.r, sC, accumulate, stream);
            EML_SKR(1) EML_SKR(2) EML_SKR(3) EML_SKR(4) EML_SKR(5) EML_SKR(6) EML_SKR(7) EML_SKR(8) EML_SKR(9) EML_SKR(10)
            EML_SKR(11) EML_SKR(12) EML_SKR(13) EML_SKR(14) EML_SKR(15) EML_SKR(16)
#undef EML_SKR
            default: break;
        }
    }
    int64_t blocks = ((M + SK_WARPS - 1) / SK_WARPS) * batch;
    // staging B costs a block as much as several rows of A: when it is staged once per block (K fits one chunk),
    // fewer, longer-lived blocks win
    const bool stage_once = K <= (int64_t)(2048 / sizeof(T)) && batch == 1;
    const int64_t cap = stage_once ? 148 * 2 : 148 * 6;
    if (blocks > cap) blocks = cap;
    launch_k(gemm_skinny_kernel<T>, dim3((unsigned)blocks), dim3(SK_THREADS), 0, stream, A, B, C, batch, M, (int)N, K, sA, sB, sC,
                                                                       accumulate ? 1 : 0);
    count_launch();
    set_last_kernel("skinny_n");
    return check_launch("gemm_skinny");
}
template int launch_gemm_skinny<float>(const float*, const float*, float*, int64_t, int64_t, int64_t, int64_t, Strides3,
                                       Strides3, Strides3, bool, cudaStream_t);
template int launch_gemm_skinny<double>(const double*, const double*, double*, int64_t, int64_t, int64_t, int64_t,
                                        Strides3, Strides3, Strides3, bool, cudaStream_t);

namespace {
template <class T>
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const T* __restrict__ ws, int splits, int64_t split_stride, T* __restrict__ C, int64_t M,
                     int64_t N, Strides3 sC, int64_t batch, int accumulate) {
    pdl_prologue();
    const int64_t total = batch * M * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t b = i / (M * N), r = (i / N) % M, c = i % N;
        T* p = C + b * sC.b + r * sC.r + c * sC.c;
        T acc = accumulate == 1 ? *p : T(0);  // (mode 2, "0 + sum", is this very chain started from +0)
        for (int s = 0; s < splits; s += 8) {  // eight loads in flight, added in split order
            T v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = s + u < splits ? ws[(int64_t)(s + u) * split_stride + i] : T(0);
#pragma unroll
            for (int u = 0; u < 8; u++)
                if (s + u < splits) acc += v[u];
        }
        *p = acc;
    }
}
}  // namespace

namespace {
// few outputs, many splits (the loss reduction: 10 outputs x 128 splits): one warp per output element, lanes take
// every 32nd split and a fixed shuffle tree combines them -- still a fixed order
template <class T>
__global__ void __launch_bounds__(256)
splitk_reduce_warp_kernel(const T* __restrict__ ws, int splits, int64_t split_stride, T* __restrict__ C, int64_t M,
                          int64_t N, Strides3 sC, int64_t batch, int accumulate) {
    pdl_prologue();
    const int lane = threadIdx.x & 31;
    const int64_t total = batch * M * N;
    for (int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); i < total; i += (int64_t)gridDim.x * 8) {
        T acc = T(0);
        for (int s = lane; s < splits; s += 32 * 8) {  // eight loads in flight per lane, added in split order
            T v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = s + 32 * u < splits ? ws[(int64_t)(s + 32 * u) * split_stride + i] : T(0);
#pragma unroll
            for (int u = 0; u < 8; u++) acc += v[u];
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) {
            int64_t b = i / (M * N), r = (i / N) % M, c = i % N;
            T* p = C + b * sC.b + r * sC.r + c * sC.c;
            *p = accumulate == 1 ? *p + acc : (accumulate == 2 ? T(0) + acc : acc);
        }
    }
}
}  // namespace

namespace {
// Few outputs, MANY splits (one partial per CTA of a k-parallel kernel: 5120 outputs x 296 partials): consecutive
// threads take consecutive outputs (coalesced; one warp per output reads one useful number per 32-byte sector), a CTA
// adds one contiguous range of splits in order, and the CTA that draws the last ticket of its output block adds the
// ranges in order.  The ranges depend on the shape only: a fixed order.
template <class T>
__global__ void __launch_bounds__(256)
splitk_reduce_2level_kernel(const T* __restrict__ ws, int splits, int per_range, int64_t total, T* part, unsigned* tickets,
                            T* __restrict__ C, int64_t M, int64_t N, Strides3 sC, int accumulate) {
    pdl_prologue();
    __shared__ bool last;
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int s0 = blockIdx.y * per_range, s1 = s0 + per_range < splits ? s0 + per_range : splits;
    if (i < total) {
        T acc = T(0);
        for (int s = s0; s < s1; s += 8) {
            T v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = s + u < s1 ? ws[(int64_t)(s + u) * total + i] : T(0);
#pragma unroll
            for (int u = 0; u < 8; u++) acc += v[u];
        }
        part[(int64_t)blockIdx.y * total + i] = acc;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicInc(tickets + blockIdx.x, gridDim.y - 1) == gridDim.y - 1;
    __syncthreads();
    if (!last || i >= total) return;
    __threadfence();
    T acc = T(0);
    for (int r = 0; r < (int)gridDim.y; r += 8) {
        T v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = r + u < (int)gridDim.y ? __ldcg(part + (int64_t)(r + u) * total + i) : T(0);
#pragma unroll
        for (int u = 0; u < 8; u++) acc += v[u];
    }
    const int64_t b = i / (M * N), rr = (i / N) % M, c = i % N;
    T* p = C + b * sC.b + rr * sC.r + c * sC.c;
    *p = accumulate == 1 ? *p + acc : (accumulate == 2 ? T(0) + acc : acc);
}
}  // namespace

bool splitk_reduce_is_sequential(int64_t total, int splits) {
    if (total <= 65536 && splits >= 64 && (total + 255) / 256 <= 4096) return false;  // two-level
    if (total <= 8192 && splits >= 16) return false;                                  // warp per element
    return true;
}

template <class T>
int launch_splitk_reduce(const T* ws, int splits, T* C, int64_t batch, int64_t M, int64_t N, Strides3 sC,
                         int accumulate, cudaStream_t stream) {
    int64_t total = batch * M * N;
    if (total == 0) return EML_OK;
    if (total <= 65536 && splits >= 64 && (total + 255) / 256 <= 4096) {
        const int ranges = 16, per_range = (splits + ranges - 1) / ranges;
        const int used = (splits + per_range - 1) / per_range;
        TempBuf part;
        EML_TRY(part.alloc(sizeof(T) * (size_t)(used * total), stream));
        unsigned* tickets = nullptr;
        EML_TRY(stream_tickets(stream, &tickets));
        launch_k(splitk_reduce_2level_kernel<T>, dim3((unsigned)((total + 255) / 256), (unsigned)used), dim3(256), 0, stream, ws,
                 splits, per_range, total, part.as<T>(), tickets, C, M, N, sC, accumulate);
        count_launch();
        return check_launch("split-K reduce (two level)");
    }
    if (total <= 8192 && splits >= 16) {
        int blocks = (int)((total + 7) / 8);
        launch_k(splitk_reduce_warp_kernel<T>, dim3(blocks), dim3(256), 0, stream, ws, splits, total, C, M, N, sC, batch, accumulate);
        count_launch();
        return check_launch("split-K reduce (warp per element)");
    }
    int64_t want = (total + 255) / 256;
    int blocks = (int)(want < 148 * 8 ? want : 148 * 8);
    launch_k(splitk_reduce_kernel<T>, dim3(blocks), dim3(256), 0, stream, ws, splits, total, C, M, N, sC, batch, accumulate);
    count_launch();
    return check_launch("split-K reduce");
}
template int launch_splitk_reduce<float>(const float*, int, float*, int64_t, int64_t, int64_t, Strides3, int,
                                         cudaStream_t);
template int launch_splitk_reduce<double>(const double*, int, double*, int64_t, int64_t, int64_t, Strides3, int,
                                          cudaStream_t);

template int launch_gemm_simt<float>(const float*, const float*, float*, int64_t, int64_t, int64_t, int64_t,
                                     Strides3, Strides3, Strides3, bool, bool, cudaStream_t);
template int launch_gemm_simt<double>(const double*, const double*, double*, int64_t, int64_t, int64_t, int64_t,
                                      Strides3, Strides3, Strides3, bool, bool, cudaStream_t);

}  // namespace eml
