// tape_kernels.cu -- single-launch kernels for the parts of a reverse sweep that are neither a tensor-core GEMM nor an
// elementwise group:
//
//   * the reference's constant-operand behaviour: a constant container carries Index 0 and record_scalar_product records
//     it as a parent (reference container_operations.rs:1291-1296, container_record/mod.rs:120-126), so the sum of its
//     would-be gradient lands in derivatives[0]:
//        A constant:  sum_ik (G B^T)[i,k] = sum_j colsum(G)[j] * colsum(B)[j]          (quirk_cols)
//        B constant:  sum_kj (A^T G)[k,j] = sum_i rowsum(A)[i] * rowsum(G)[i]          (quirk_rows)
//     one launch each (block partials, a ticket elects the last block, which adds them in index order);
//   * "thin" products, M <= 4 rows (a ones-row times a matrix is how a container is summed; a single sample through a
//     layer): forward and the whole backward of the node -- dA, dB and either constant-operand sum -- in one launch each;
//   * C (+)= A * B with a short contraction (K <= 32) and a wide output (dH = dY * W2^T of the output layer): every
//     thread owns one 16-byte vector of an output row, B sits in shared memory, no split / reduce pass.
//
// All sums run in a fixed order (no atomics on data; the atomic only elects the finishing block): bit-identical run to
// run.  Products accumulate with fma in ascending k like the other GPU GEMM kernels.
#include <type_traits>

#include "common.h"

namespace eml {

namespace {

constexpr int TK_THREADS = 256;

template <class T>
__device__ __forceinline__ T block_sum(T v, T* smem /* >= 8 */) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) smem[w] = v;
    __syncthreads();
    T total = T(0);
    if (threadIdx.x == 0)
        for (int i = 0; i < TK_THREADS / 32; i++) total += smem[i];
    return total;  // valid in thread 0
}

// true in every thread of exactly one block: the one that drew the last ticket (the counter resets itself)
__device__ __forceinline__ bool last_block(unsigned* ticket, unsigned blocks) {
    __shared__ bool last;
    if (blocks == 1) {  // nobody to wait for: no fence, no atomic round trip
        __syncthreads();
        return true;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicInc(ticket, blocks - 1) == blocks - 1;
    __syncthreads();
    if (last) __threadfence();
    return last;
}

// sum of X[r * stride + c] for r = first, first + step, ... < end, in that order.  Eight loads are issued before the
// first add (out-of-range ones are replaced by +0): a loop with one dependent load per iteration pays the full memory
// latency per element, which is what bounds these small reductions.
template <class T>
__device__ __forceinline__ T strided_sum(const T* __restrict__ X, int64_t first, int64_t end, int64_t step, int64_t stride,
                                         int64_t c) {
    T acc = T(0);
    for (int64_t r = first; r < end; r += 8 * step) {
        T v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int64_t rr = r + u * step;
            v[u] = rr < end ? __ldcg(X + rr * stride + c) : T(0);
        }
#pragma unroll
        for (int u = 0; u < 8; u++) acc += v[u];
    }
    return acc;
}

// partial[c] = sum over rows [r0, r1) of X[r * N + c] for all c < N; a block-wide, fixed-order column sum.
// N <= 256: thread t owns column t % N and every (256 / N)-th row, so a warp reads consecutive addresses.
template <class T>
__device__ __forceinline__ void block_col_sums(const T* __restrict__ X, int64_t r0, int64_t r1, int64_t N, T* partial,
                                               T* smem /* 256 */) {
    const int t = threadIdx.x;
    if (N <= TK_THREADS) {
        const int L = TK_THREADS / (int)N, rl = t / (int)N, c = t % (int)N;
        T acc = T(0);
        if (rl < L) acc = strided_sum<T>(X, r0 + rl, r1, L, N, c);
        __syncthreads();
        smem[t] = acc;
        __syncthreads();
        if (t < N) {
            T tot = T(0);
            for (int l = 0; l < L; l++) tot += smem[l * (int)N + t];
            partial[t] = tot;
        }
    } else {
        for (int64_t c = t; c < N; c += TK_THREADS) partial[c] = strided_sum<T>(X, r0, r1, 1, N, c);
    }
}

// column sums of rows [r0, r1) for the 32 columns starting at c0: 8 row lanes x 32 columns, eight independent loads in
// flight per thread (a single dependent load per thread is latency-bound), fixed order
template <class T>
__device__ __forceinline__ T tile_col_sum(const T* __restrict__ X, int64_t r0, int64_t r1, int64_t c, int64_t N, bool active,
                                          T (*smem)[33]) {
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    T acc = T(0);
    if (active) acc = strided_sum<T>(X, r0 + ty, r1, 8, N, c);
    __syncthreads();
    smem[ty][tx] = acc;
    __syncthreads();
    T t = T(0);
    if (ty == 0)
        for (int i = 0; i < 8; i++) t += smem[i][tx];
    return t;  // valid for ty == 0
}

// grid (column blocks of 32, row chunks): block (x, y) sums its chunk of G's rows and of B's rows for its 32 columns;
// the last chunk of a column block (ticket x) adds the chunks in order and forms sum_c gs[c] * bs[c] over its columns;
// the last column block (ticket MAX - 1) adds those in order and writes the sum to out[0].
template <class T>
__global__ void __launch_bounds__(TK_THREADS)
quirk_cols_kernel(const T* __restrict__ G, int64_t gm, int64_t g_rows, const T* __restrict__ B, int64_t bm, int64_t b_rows,
                  int64_t N, T* partial_g, T* partial_b, T* partial_cb, unsigned* tickets, T* out) {
    pdl_prologue();
    __shared__ T smem[8][33];
    __shared__ bool last;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t c = (int64_t)blockIdx.x * 32 + tx;
    const bool active = c < N;
    {
        const int64_t r0 = (int64_t)blockIdx.y * g_rows, r1 = r0 + g_rows < gm ? r0 + g_rows : gm;
        const T t = tile_col_sum<T>(G, r0, r1, c, N, active, smem);
        if (ty == 0 && active) partial_g[(int64_t)blockIdx.y * N + c] = t;
    }
    {
        const int64_t r0 = (int64_t)blockIdx.y * b_rows, r1 = r0 + b_rows < bm ? r0 + b_rows : bm;
        const T t = tile_col_sum<T>(B, r0, r1, c, N, active, smem);
        if (ty == 0 && active) partial_b[(int64_t)blockIdx.y * N + c] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicInc(tickets + blockIdx.x, gridDim.y - 1) == gridDim.y - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    // chunks in order (eight row lanes take every eighth chunk, then the lanes in order)
    T gs = T(0), bs = T(0);
    if (active) {
        gs = strided_sum<T>(partial_g, ty, gridDim.y, 8, N, c);
        bs = strided_sum<T>(partial_b, ty, gridDim.y, 8, N, c);
    }
    __syncthreads();
    smem[ty][tx] = gs;
    __syncthreads();
    T gtot = T(0);
    if (ty == 0)
        for (int i = 0; i < 8; i++) gtot += smem[i][tx];
    __syncthreads();
    smem[ty][tx] = bs;
    __syncthreads();
    if (ty == 0) {
        T btot = T(0);
        for (int i = 0; i < 8; i++) btot += smem[i][tx];
        T v = active ? gtot * btot : T(0);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (tx == 0) partial_cb[blockIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicInc(tickets + TICKET_SECOND_LEVEL, gridDim.x - 1) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (threadIdx.x == 0) {
        T total = T(0);
        for (int i = 0; i < (int)gridDim.x; i++) total += __ldcg(partial_cb + i);
        out[0] = total;  // (a slot of its own: the sweep folds the slots in order)
    }
}

// out[0] = sum_i rowsum(A)[i] * rowsum(G)[i]; one warp per row, rows of a warp in ascending order
template <class T>
__global__ void __launch_bounds__(TK_THREADS)
quirk_rows_kernel(const T* __restrict__ A, int64_t K, const T* __restrict__ G, int64_t N, int64_t M, T* partial,
                  unsigned* ticket, T* out) {
    pdl_prologue();
    __shared__ T smem[TK_THREADS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    T acc = T(0);
    for (int64_t i = (int64_t)blockIdx.x * 8 + w; i < M; i += (int64_t)gridDim.x * 8) {
        T ra = strided_sum<T>(A + i * K, lane, K, 32, 1, 0), rg = strided_sum<T>(G + i * N, lane, N, 32, 1, 0);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ra += __shfl_xor_sync(0xffffffffu, ra, o);
            rg += __shfl_xor_sync(0xffffffffu, rg, o);
        }
        acc += ra * rg;  // the same value in every lane
    }
    __syncthreads();
    if (lane == 0) smem[w] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        T tot = T(0);
        for (int i = 0; i < 8; i++) tot += smem[i];
        partial[blockIdx.x] = tot;
    }
    if (!last_block(ticket, gridDim.x)) return;
    T v = T(0);
    for (int i = threadIdx.x; i < (int)gridDim.x; i += TK_THREADS) v += __ldcg(partial + i);
    const T total = block_sum<T>(v, smem);
    if (threadIdx.x == 0) out[0] = total;  // (a slot of its own: the sweep folds the slots in order)
}

// ---- thin products: M <= TH_M rows --------------------------------------------------------------------------------
constexpr int TH_M = 4, TH_N = 256;

// C[i][c] (+)= sum_k A(i,k) * B[k][c]; B dense row-major K x N, A and C strided; blocks own k chunks
template <class T>
__global__ void __launch_bounds__(TK_THREADS)
thin_fwd_kernel(const T* __restrict__ A, int64_t sAr, int64_t sAc, const T* __restrict__ B, int M, int N, int64_t K,
                int64_t k_per_block, T* partial, unsigned* ticket, T* C, int64_t sCr, int64_t sCc, int accumulate) {
    pdl_prologue();
    __shared__ T smem[TH_M][TK_THREADS];
    const int t = threadIdx.x, L = TK_THREADS / N, rl = t / N, c = t % N;
    const int64_t k0 = (int64_t)blockIdx.x * k_per_block, k1 = k0 + k_per_block < K ? k0 + k_per_block : K;
    T acc[TH_M];
#pragma unroll
    for (int i = 0; i < TH_M; i++) acc[i] = T(0);
    if (rl < L) {
        for (int64_t k = k0 + rl; k < k1; k += 4 * (int64_t)L) {  // four k per batch: all loads before the first fma
            T bv[4], av[4][TH_M];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int64_t kk = k + (int64_t)u * L;
                bv[u] = kk < k1 ? B[kk * N + c] : T(0);
#pragma unroll
                for (int i = 0; i < TH_M; i++) av[u][i] = (i < M && kk < k1) ? A[i * sAr + kk * sAc] : T(0);
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int i = 0; i < TH_M; i++)
                    if (i < M) acc[i] = fma(av[u][i], bv[u], acc[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < TH_M; i++) smem[i][t] = acc[i];
    __syncthreads();
    if (t < N) {
        for (int i = 0; i < M; i++) {
            T tot = T(0);
            for (int l = 0; l < L; l++) tot += smem[i][l * N + t];
            partial[((int64_t)blockIdx.x * M + i) * N + t] = tot;
        }
    }
    if (!last_block(ticket, gridDim.x)) return;
    // one warp per output element: lanes take every 32nd block's partial, a fixed shuffle tree adds the lanes
    const int lane = t & 31;
    for (int e = t >> 5; e < M * N; e += TK_THREADS / 32) {
        const int i = e / N, cc = e % N;
        T tot = T(0);
        for (int b = lane; b < (int)gridDim.x; b += 32) tot += __ldcg(partial + ((int64_t)b * M + i) * N + cc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if (lane == 0) {
            T* p = C + i * sCr + cc * sCc;
            *p = accumulate ? *p + tot : tot;
        }
    }
}

// The whole reverse sweep of one thin node C = A * B (A: M x K, B: K x N, G = dC: M x N, all dense row-major):
//   dA[i][k] (+)= sum_j G[i][j] B[k][j]          (dA != null)        or  slot0 += sum_j colsum(G)[j] colsum(B)[j]  (quirk_a)
//   dB[k][j] (+)= sum_i A[i][k] G[i][j]          (dB != null)        or  slot0 += sum_i rowsum(A)[i] rowsum(G)[i]  (quirk_b)
// blocks own k chunks; the constant-operand sums finish in the block that draws the last ticket.
template <class T>
__global__ void __launch_bounds__(TK_THREADS)
thin_bwd_kernel(const T* __restrict__ G, const T* __restrict__ A, const T* __restrict__ B, int M, int N, int64_t K,
                int64_t k_per_block, T* dA, int over_a, T* dB, int over_b, int quirk_a, int quirk_b, T* partial,
                unsigned* ticket, T* slot0) {
    pdl_prologue();
    __shared__ T sG[TH_M * TH_N];
    __shared__ T smem[TK_THREADS];
    const int t = threadIdx.x;
    for (int e = t; e < M * N; e += TK_THREADS) sG[e] = G[e];
    __syncthreads();
    const int64_t k0 = (int64_t)blockIdx.x * k_per_block, k1 = k0 + k_per_block < K ? k0 + k_per_block : K;
    if (dA) {
        for (int64_t k = k0 + t; k < k1; k += TK_THREADS)
            for (int i = 0; i < M; i++) {
                T acc = over_a ? T(0) : dA[i * K + k];
                for (int j = 0; j < N; j++) acc = fma(sG[i * N + j], B[k * N + j], acc);
                dA[i * K + k] = acc;
            }
    }
    if (dB) {
        const int span = (int)((k1 - k0) * N);  // <= k_per_block * 256
        for (int e = t; e < span; e += TK_THREADS) {
            const int64_t k = k0 + e / N;
            const int j = e % N;
            T acc = over_b ? T(0) : dB[k0 * N + e];
            for (int i = 0; i < M; i++) acc = fma(A[i * K + k], sG[i * N + j], acc);
            dB[k0 * N + e] = acc;
        }
    }
    if (!quirk_a && !quirk_b) return;
    if (quirk_a) block_col_sums<T>(B, k0, k1, N, partial + (int64_t)blockIdx.x * N, smem);
    if (quirk_b) {
        for (int i = 0; i < M; i++) {
            T v = T(0);
            for (int64_t k = k0 + t; k < k1; k += TK_THREADS) v += A[i * K + k];
            const T tot = block_sum<T>(v, smem);
            if (t == 0) partial[(int64_t)blockIdx.x * TH_M + i] = tot;
        }
    }
    if (!last_block(ticket, gridDim.x)) return;
    // one warp per column (quirk_a) / row (quirk_b): lanes take every 32nd block's partial, fixed shuffle tree; the
    // warps' products are then added in index order
    const int lane = t & 31, w = t >> 5;
    const int items = quirk_a ? N : M, width = quirk_a ? N : TH_M;
    T v = T(0);
    for (int e = w; e < items; e += TK_THREADS / 32) {
        T ps = T(0);
        for (int b = lane; b < (int)gridDim.x; b += 32) ps += __ldcg(partial + (int64_t)b * width + e);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
        T gs = T(0);
        if (quirk_a)
            for (int i = 0; i < M; i++) gs += sG[i * N + e];
        else
            for (int j = 0; j < N; j++) gs += sG[e * N + j];
        v += gs * ps;  // the same value in every lane
    }
    __syncthreads();
    if (lane == 0) smem[w] = v;
    __syncthreads();
    if (t == 0) {
        T total = T(0);
        for (int i = 0; i < TK_THREADS / 32; i++) total += smem[i];
        slot0[0] = total;  // (a slot of its own: the sweep folds the slots in order)
    }
}

// ---- short contraction, wide output: C[m][n] (+)= sum_{k < K} A[m][k] * B(k, n), K <= 32 ---------------------------
constexpr int SMK_MAX_K = 32;

template <class T, int V>
__global__ void __launch_bounds__(TK_THREADS)
gemm_smallk_kernel(const T* __restrict__ A, int64_t lda, const T* __restrict__ B, int64_t sBr, int64_t sBc, T* __restrict__ C,
                   int64_t ldc, int64_t M, int64_t N, int K, int accumulate) {
    pdl_prologue();
    extern __shared__ __align__(16) unsigned char smk_raw[];
    T* sB = reinterpret_cast<T*>(smk_raw);  // [K][N]; staged once per (persistent) block
    if (sBc == 1) {
        for (int64_t e = threadIdx.x; e < (int64_t)K * N; e += TK_THREADS) sB[e] = B[(e / N) * sBr + e % N];
    } else {  // B(k, n) contiguous along k (a transposed operand): read along k, coalesced
        for (int64_t e = threadIdx.x; e < (int64_t)K * N; e += TK_THREADS) {
            const int64_t n = e / K, k = e % K;
            sB[k * N + n] = B[k * sBr + n * sBc];
        }
    }
    __syncthreads();
    // a block takes `rows_per_pass` whole rows per pass when a row has at most 256 vectors (no division in the loop);
    // wider rows: one row per pass, threads stride over its vectors
    const int nv = (int)(N / V);
    const int rows_per_pass = nv <= TK_THREADS ? TK_THREADS / nv : 1;
    const int my_row = nv <= TK_THREADS ? (int)threadIdx.x / nv : 0;
    const int my_first = nv <= TK_THREADS ? (int)threadIdx.x % nv : (int)threadIdx.x;
    const bool idle = nv <= TK_THREADS && my_row >= rows_per_pass;
    for (int64_t m0 = (int64_t)blockIdx.x * rows_per_pass; m0 < M && !idle; m0 += (int64_t)gridDim.x * rows_per_pass)
    for (int vi = my_first; vi < nv; vi += TK_THREADS) {
        const int64_t m = m0 + my_row, n = (int64_t)vi * V;
        if (m >= M) break;
        T acc[V];
        T* cp = C + m * ldc + n;
        if (accumulate) {
            if constexpr (V > 1) {
                const auto c4 = *reinterpret_cast<const typename std::conditional<sizeof(T) == 4, float4, double2>::type*>(cp);
                const T* ce = reinterpret_cast<const T*>(&c4);
#pragma unroll
                for (int j = 0; j < V; j++) acc[j] = ce[j];
            } else {
                acc[0] = cp[0];
            }
        } else {
#pragma unroll
            for (int j = 0; j < V; j++) acc[j] = T(0);
        }
        const T* a = A + m * lda;
        for (int k = 0; k < K; k++) {
            const T av = __ldg(a + k);
            T bv[V];
            if constexpr (V > 1) {
                const auto b4 = *reinterpret_cast<const typename std::conditional<sizeof(T) == 4, float4, double2>::type*>(sB + (int64_t)k * N + n);
                const T* be = reinterpret_cast<const T*>(&b4);
#pragma unroll
                for (int j = 0; j < V; j++) bv[j] = be[j];
            } else {
                bv[0] = sB[(int64_t)k * N + n];
            }
#pragma unroll
            for (int j = 0; j < V; j++) acc[j] = fma(av, bv[j], acc[j]);
        }
        if constexpr (V > 1) {
            typename std::conditional<sizeof(T) == 4, float4, double2>::type o4;
            T* oe = reinterpret_cast<T*>(&o4);
#pragma unroll
            for (int j = 0; j < V; j++) oe[j] = acc[j];
            *reinterpret_cast<decltype(o4)*>(cp) = o4;
        } else {
            cp[0] = acc[0];
        }
    }
}

}  // namespace

template <class T>
int launch_quirk_cols(const T* G, int64_t gm, const T* B, int64_t bm, int64_t N, T* out, cudaStream_t stream) {
    if (gm < 1 || bm < 1 || N < 1) return EML_OK;
    const int64_t col_blocks = (N + 31) / 32;
    if (col_blocks > TICKET_NONFINITE) EML_BAD_ARG("constant-operand sums: more than %d columns", TICKET_NONFINITE * 32);
    // row chunks: ~2 CTAs per SM over the whole grid, at least 64 rows of the taller matrix each
    const int64_t rows = gm > bm ? gm : bm;
    int64_t chunks = (4 * 148 + col_blocks - 1) / col_blocks;  // short chunks: a few batches of loads per thread
    if (chunks > (rows + 63) / 64) chunks = (rows + 63) / 64;
    if (chunks > 256) chunks = 256;
    if (chunks < 1) chunks = 1;
    const int64_t g_rows = (gm + chunks - 1) / chunks, b_rows = (bm + chunks - 1) / chunks;
    TempBuf partial;
    EML_TRY(partial.alloc(sizeof(T) * (size_t)(2 * chunks * N + col_blocks), stream));
    T* pg = partial.as<T>();
    unsigned* tickets = nullptr;
    EML_TRY(stream_tickets(stream, &tickets));
    launch_k(quirk_cols_kernel<T>, dim3((unsigned)col_blocks, (unsigned)chunks), dim3(TK_THREADS), 0, stream, G, gm, g_rows, B, bm,
             b_rows, N, pg, pg + chunks * N, pg + 2 * chunks * N, tickets, out);
    count_launch();
    return check_launch("constant-operand column sums");
}

template <class T>
int launch_quirk_rows(const T* A, int64_t K, const T* G, int64_t N, int64_t M, T* out, cudaStream_t stream) {
    if (M < 1) return EML_OK;
    int64_t nb = (M * (K + N) + 65535) / 65536;
    if (nb > 128) nb = 128;
    if (nb > (M + 7) / 8) nb = (M + 7) / 8;
    if (nb < 1) nb = 1;
    TempBuf partial;
    EML_TRY(partial.alloc(sizeof(T) * (size_t)nb, stream));
    unsigned* tickets = nullptr;
    EML_TRY(stream_tickets(stream, &tickets));
    launch_k(quirk_rows_kernel<T>, dim3((unsigned)nb), dim3(TK_THREADS), 0, stream, A, K, G, N, M, partial.as<T>(), tickets, out);
    count_launch();
    return check_launch("constant-operand row sums");
}

namespace {
// derivatives[index] += sum of g over the elements of a re-wrapped container that carry that index: one thread per
// DISTINCT index walks its elements (perm groups them, ascending element order) -- a fixed order, no atomics
template <class T>
__global__ void __launch_bounds__(TK_THREADS)
scatter_segments_kernel(const T* __restrict__ g, const int* __restrict__ perm, const int* __restrict__ seg_start, int n_seg,
                        T* arena, const int64_t* __restrict__ seg_off) {
    pdl_prologue();
    for (int sgm = blockIdx.x * TK_THREADS + threadIdx.x; sgm < n_seg; sgm += gridDim.x * TK_THREADS) {
        T acc = T(0);
        for (int j = seg_start[sgm]; j < seg_start[sgm + 1]; j++) acc += g[perm[j]];
        arena[seg_off[sgm]] = arena[seg_off[sgm]] + acc;
    }
}
}  // namespace

template <class T>
int launch_scatter_segments(const T* g, const int* perm, const int* seg_start, int n_seg, T* arena, const int64_t* seg_off,
                            cudaStream_t stream) {
    if (n_seg < 1) return EML_OK;
    int blocks = (n_seg + TK_THREADS - 1) / TK_THREADS;
    if (blocks > 148 * 8) blocks = 148 * 8;
    launch_k(scatter_segments_kernel<T>, dim3(blocks), dim3(TK_THREADS), 0, stream, g, perm, seg_start, n_seg, arena, seg_off);
    count_launch();
    return check_launch("derivative scatter of a re-wrapped container");
}

namespace {
// the reference's reverse loop (src/differentiation.rs:632-645) over one run of scalar entries: sequential by nature
// (an entry's parents may be earlier entries of the same run), so one thread walks it
template <class T>
__global__ void scalar_run_kernel(T* arena, int64_t self_off, int64_t n, const int64_t* __restrict__ lp, const int64_t* __restrict__ rp,
                                  const T* __restrict__ ld, const T* __restrict__ rd) {
    pdl_prologue();
    if (blockIdx.x || threadIdx.x) return;
    for (int64_t e = n - 1; e >= 0; e--) {
        const T g = arena[self_off + e];
        arena[lp[e]] = arena[lp[e]] + g * ld[e];
        arena[rp[e]] = arena[rp[e]] + g * rd[e];
    }
}
}  // namespace

template <class T>
int launch_scalar_run(T* arena, int64_t self_off, int64_t n, const int64_t* lp, const int64_t* rp, const T* ld, const T* rd,
                      cudaStream_t stream) {
    if (n < 1) return EML_OK;
    launch_k(scalar_run_kernel<T>, dim3(1), dim3(32), 0, stream, arena, self_off, n, lp, rp, ld, rd);
    count_launch();
    return check_launch("scalar entries of the sweep");
}

bool thin_shape(int64_t M, int64_t N, int64_t K) { return M >= 1 && M <= TH_M && N >= 1 && N <= TH_N && K >= 1; }

namespace {
int thin_blocks(int64_t N, int64_t K, int64_t* k_per_block) {
    // a block walks ~1K elements of B (these products are small: spread them over the machine)
    int64_t nb = (K * N + 1023) / 1024;
    if (nb > 148) nb = 148;
    if (nb < 1) nb = 1;
    *k_per_block = (K + nb - 1) / nb;
    return (int)((K + *k_per_block - 1) / *k_per_block);
}
}  // namespace

template <class T>
int launch_thin_fwd(const T* A, int64_t sAr, int64_t sAc, const T* B, int64_t M, int64_t N, int64_t K, T* C, int64_t sCr,
                    int64_t sCc, bool accumulate, cudaStream_t stream) {
    int64_t kpb;
    const int nb = thin_blocks(N, K, &kpb);
    TempBuf partial;
    EML_TRY(partial.alloc(sizeof(T) * (size_t)(nb * M * N), stream));
    unsigned* tickets = nullptr;
    EML_TRY(stream_tickets(stream, &tickets));
    launch_k(thin_fwd_kernel<T>, dim3(nb), dim3(TK_THREADS), 0, stream, A, sAr, sAc, B, (int)M, (int)N, K, kpb, partial.as<T>(),
             tickets, C, sCr, sCc, accumulate ? 1 : 0);
    count_launch();
    set_last_kernel("thin");
    return check_launch("thin product");
}

template <class T>
int launch_thin_bwd(const T* G, const T* A, const T* B, int64_t M, int64_t N, int64_t K, T* dA, bool over_a, T* dB,
                    bool over_b, bool quirk_a, bool quirk_b, T* slot0, cudaStream_t stream) {
    int64_t kpb;
    const int nb = thin_blocks(N, K, &kpb);
    TempBuf partial;
    EML_TRY(partial.alloc(sizeof(T) * (size_t)(nb * (N > TH_M ? N : TH_M)), stream));
    unsigned* tickets = nullptr;
    EML_TRY(stream_tickets(stream, &tickets));
    launch_k(thin_bwd_kernel<T>, dim3(nb), dim3(TK_THREADS), 0, stream, G, A, B, (int)M, (int)N, K, kpb, dA, over_a ? 1 : 0, dB,
             over_b ? 1 : 0, quirk_a ? 1 : 0, quirk_b ? 1 : 0, partial.as<T>(), tickets, slot0);
    count_launch();
    return check_launch("thin product backward");
}

namespace {
// The same product with this thread's slice of B (KT x V numbers) held in REGISTERS: a thread owns one 16-byte vector
// of output columns and walks down the rows, so per output vector it issues K broadcast loads of A, K * V fmas and one
// 16-byte store -- no shared-memory traffic, no index arithmetic in the loop.  K <= KT <= 16.
template <class T, int V, int KT>
__global__ void __launch_bounds__(TK_THREADS)
gemm_smallk_reg_kernel(const T* __restrict__ A, int64_t lda, const T* __restrict__ B, int64_t sBr, int64_t sBc, T* __restrict__ C,
                       int64_t ldc, int64_t M, int64_t N, int K, int accumulate) {
    pdl_prologue();
    using V4 = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
    const int nv = (int)(N / V);                       // vectors per row (<= 256)
    const int lanes = TK_THREADS / nv;                 // rows a CTA handles per pass
    const int my_row = (int)threadIdx.x / nv, my_vec = (int)threadIdx.x % nv;
    extern __shared__ __align__(16) unsigned char smk_raw[];
    T* sB = reinterpret_cast<T*>(smk_raw);
    // B(k, n) = B[k + n * K] (a row-major [N][K] matrix read transposed: dH = dY * W2^T): copied as it lies, with
    // 16-byte loads and no index arithmetic; a thread's K x V slice is then V * K consecutive numbers
    const int total_b = K * (int)N;
    for (int e0 = 0; e0 < total_b; e0 += 4 * TK_THREADS) {
        T v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int e = e0 + u * TK_THREADS + (int)threadIdx.x;
            v[u] = e < total_b ? B[e] : T(0);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int e = e0 + u * TK_THREADS + (int)threadIdx.x;
            if (e < total_b) sB[e] = v[u];
        }
    }
    __syncthreads();
    if (my_row >= lanes) return;
    const int64_t n = (int64_t)my_vec * V;
    T b[KT][V];
#pragma unroll
    for (int k = 0; k < KT; k++)
#pragma unroll
        for (int j = 0; j < V; j++) b[k][j] = k < K ? sB[((int)n + j) * K + k] : T(0);
    for (int64_t m = (int64_t)blockIdx.x * lanes + my_row; m < M; m += (int64_t)gridDim.x * lanes) {
        const T* a = A + m * lda;
        T* cp = C + m * ldc + n;
        T av[KT];
#pragma unroll
        for (int k = 0; k < KT; k++) av[k] = k < K ? __ldg(a + k) : T(0);
        T acc[V];
        if (accumulate) {
            const V4 c4 = *reinterpret_cast<const V4*>(cp);
            const T* ce = reinterpret_cast<const T*>(&c4);
#pragma unroll
            for (int j = 0; j < V; j++) acc[j] = ce[j];
        } else {
#pragma unroll
            for (int j = 0; j < V; j++) acc[j] = T(0);
        }
#pragma unroll
        for (int k = 0; k < KT; k++)
            if (k < K) {
#pragma unroll
                for (int j = 0; j < V; j++) acc[j] = fma(av[k], b[k][j], acc[j]);
            }
        V4 o4;
        T* oe = reinterpret_cast<T*>(&o4);
#pragma unroll
        for (int j = 0; j < V; j++) oe[j] = acc[j];
        *reinterpret_cast<V4*>(cp) = o4;
    }
}
}  // namespace

bool smallk_shape(int64_t M, int64_t N, int64_t K, size_t elem) {
    return K >= 1 && K <= SMK_MAX_K && N >= 64 && M >= 64 && (size_t)(K * N) * elem <= (size_t(48) << 10);
}

template <class T>
int launch_gemm_smallk(const T* A, int64_t lda, const T* B, int64_t sBr, int64_t sBc, T* C, int64_t ldc, int64_t M, int64_t N,
                       int64_t K, bool accumulate, cudaStream_t stream) {
    constexpr int VN = 16 / sizeof(T);
    const bool vec = N % VN == 0 && ldc % VN == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0;
    const int64_t total = vec ? M * (N / VN) : M * N;
    int64_t blocks = (total + TK_THREADS - 1) / TK_THREADS;
    if (blocks > 148 * 4) blocks = 148 * 4;  // persistent: B is staged once per block
    if (blocks > M) blocks = M;
    const size_t smem = sizeof(T) * (size_t)(K * N);
    if (vec && K <= 16 && N / VN <= TK_THREADS && sBr == 1 && sBc == K) {
        const int64_t lanes = TK_THREADS / (N / VN);
        int64_t rb = (M + lanes - 1) / lanes;
        if (rb > 148 * 2) rb = 148 * 2;  // persistent: B is staged once per CTA
        if (K <= 8)
            launch_k(gemm_smallk_reg_kernel<T, VN, 8>, dim3((unsigned)rb), dim3(TK_THREADS), smem, stream, A, lda, B, sBr, sBc, C, ldc, M,
                     N, (int)K, accumulate ? 1 : 0);
        else
            launch_k(gemm_smallk_reg_kernel<T, VN, 16>, dim3((unsigned)rb), dim3(TK_THREADS), smem, stream, A, lda, B, sBr, sBc, C, ldc,
                     M, N, (int)K, accumulate ? 1 : 0);
        count_launch();
        set_last_kernel("small_k");
        return check_launch("short-contraction product");
    }
    if (vec)
        launch_k(gemm_smallk_kernel<T, VN>, dim3((unsigned)blocks), dim3(TK_THREADS), smem, stream, A, lda, B, sBr, sBc, C, ldc, M, N,
                 (int)K, accumulate ? 1 : 0);
    else
        launch_k(gemm_smallk_kernel<T, 1>, dim3((unsigned)blocks), dim3(TK_THREADS), smem, stream, A, lda, B, sBr, sBc, C, ldc, M, N,
                 (int)K, accumulate ? 1 : 0);
    count_launch();
    set_last_kernel("small_k");
    return check_launch("short-contraction product");
}

#define EML_INST(T)                                                                                                       \
    template int launch_scalar_run<T>(T*, int64_t, int64_t, const int64_t*, const int64_t*, const T*, const T*, cudaStream_t); \
    template int launch_scatter_segments<T>(const T*, const int*, const int*, int, T*, const int64_t*, cudaStream_t);     \
    template int launch_quirk_cols<T>(const T*, int64_t, const T*, int64_t, int64_t, T*, cudaStream_t);                   \
    template int launch_quirk_rows<T>(const T*, int64_t, const T*, int64_t, int64_t, T*, cudaStream_t);                   \
    template int launch_thin_fwd<T>(const T*, int64_t, int64_t, const T*, int64_t, int64_t, int64_t, T*, int64_t, int64_t, \
                                    bool, cudaStream_t);                                                                  \
    template int launch_thin_bwd<T>(const T*, const T*, const T*, int64_t, int64_t, int64_t, T*, bool, T*, bool, bool,    \
                                    bool, T*, cudaStream_t);                                                              \
    template int launch_gemm_smallk<T>(const T*, int64_t, const T*, int64_t, int64_t, T*, int64_t, int64_t, int64_t,      \
                                       int64_t, bool, cudaStream_t);
EML_INST(float)
EML_INST(double)
#undef EML_INST

}  // namespace eml
