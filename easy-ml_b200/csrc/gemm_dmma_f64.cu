// gemm_dmma_f64.cu -- f64 GEMM on the FP64 tensor pipe (DMMA.8x8x4) for sm_100a.
//
// tcgen05 has no f64 kind, so FP64 tensor math on Blackwell is the warp-level
// `mma.sync.aligned.m8n8k4.f64` (every larger PTX f64 shape lowers to DMMA.8x8x4 SASS on sm_100a --
// checked with cuobjdump).  Accumulators live in registers (no TMEM for f64).
//
// Tiling: CTA 128x128x16, 8 warps as 2(M) x 4(N), warp tile 64x32 = 8x4 DMMA fragments, 64 f64
// accumulators per thread.  4-stage cp.async ring in shared memory (160 KB), one CTA per SM.
// Shared layouts are padded so every fragment load (LDS.64) is bank-conflict free:
//   K-inner  tile[mn][k]  pitch 20 doubles  (operand whose k index is contiguous in memory)
//   MN-inner tile[k][mn]  pitch 132 doubles (operand whose m / n index is contiguous in memory)
// so row-major A (K-inner) x row-major B (MN-inner) and all transposed combinations run without a
// pack pass -- the backward products dA = dC * B^T and dB = A^T * dC use the same kernel.
//
// Determinism: fixed tile -> CTA map, k tiles consumed in ascending order, no atomics, no split-K.
//
// Algorithmic work per launch: 2*M*N*K flop; (M*K + K*N + M*N) * 8 bytes.
#include <cstdlib>
#include <string>

#include "common.h"
#include "sm100.cuh"

namespace eml {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 4, THREADS = 256;
constexpr int KIN_PITCH = BK + 4;    // 20
constexpr int MNIN_PITCH = BM + 4;   // 132 (BM == BN)
constexpr int OPERAND_ELEMS = BM * KIN_PITCH > BK * MNIN_PITCH ? BM * KIN_PITCH : BK * MNIN_PITCH;  // 2560
constexpr int STAGE_ELEMS = 2 * OPERAND_ELEMS;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_ELEMS * sizeof(double);  // 163840

__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_8(void* smem, const void* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Copies one operand tile into shared memory.
//   KIN:  tile[mn][k], global element (mn, k) at g[mn*ld_mn + k]          (k contiguous)
//   !KIN: tile[k][mn], global element (mn, k) at g[k*ld_k + mn]           (mn contiguous)
// `ld` is the stride of the non-contiguous index.  Out-of-range elements are zero-filled.
template <bool KIN, int VEC>
__device__ __forceinline__ void load_tile(double* smem, const double* __restrict__ g, int64_t ld, int64_t mn0,
                                          int64_t mn_extent, int64_t k0, int64_t k_extent, int tid) {
    if (KIN) {
        constexpr int CH_PER_ROW = BK / VEC;  // chunks along k per mn row
        constexpr int ITERS = BM * CH_PER_ROW / THREADS;
#pragma unroll
        for (int i = 0; i < ITERS; i++) {
            int c = tid + i * THREADS;
            int row = c / CH_PER_ROW, kc = (c % CH_PER_ROW) * VEC;
            int64_t gm = mn0 + row, gk = k0 + kc;
            int64_t left = k_extent - gk;
            int valid = (gm < mn_extent && left > 0) ? (left >= VEC ? VEC : (int)left) : 0;
            const double* src = valid ? g + gm * ld + gk : g;
            double* dst = smem + row * KIN_PITCH + kc;
            if (VEC == 2)
                cp_async_16(dst, src, valid * 8);
            else
                cp_async_8(dst, src, valid * 8);
        }
    } else {
        constexpr int CH_PER_ROW = BM / VEC;  // chunks along mn per k row
        constexpr int ITERS = BK * CH_PER_ROW / THREADS;
#pragma unroll
        for (int i = 0; i < ITERS; i++) {
            int c = tid + i * THREADS;
            int k = c / CH_PER_ROW, mc = (c % CH_PER_ROW) * VEC;
            int64_t gk = k0 + k, gm = mn0 + mc;
            int64_t left = mn_extent - gm;
            int valid = (gk < k_extent && left > 0) ? (left >= VEC ? VEC : (int)left) : 0;
            const double* src = valid ? g + gk * ld + gm : g;
            double* dst = smem + k * MNIN_PITCH + mc;
            if (VEC == 2)
                cp_async_16(dst, src, valid * 8);
            else
                cp_async_8(dst, src, valid * 8);
        }
    }
}

template <bool A_KIN, bool B_KIN, int VEC>
__global__ void __launch_bounds__(THREADS, 1)
dgemm_dmma_kernel(const double* __restrict__ A, const double* __restrict__ B, double* C, int64_t M,
                  int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int64_t strideA, int64_t strideB,
                  int64_t strideC, int tiles_m, int tiles_n, int accumulate) {
    pdl_prologue();
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32;

    // tile raster: groups of 8 tile rows walked column by column keep a wave's A/B panels L2-resident
    constexpr int GROUP = 8;
    int tile = blockIdx.x;
    int tiles_per_group = GROUP * tiles_n;
    int group = tile / tiles_per_group;
    int first_m = group * GROUP;
    int group_rows = tiles_m - first_m < GROUP ? tiles_m - first_m : GROUP;
    int in_group = tile % tiles_per_group;
    int tm = first_m + in_group % group_rows;
    int tn = in_group / group_rows;
    const int64_t m0 = (int64_t)tm * BM, n0 = (int64_t)tn * BN;

    const int64_t batch = blockIdx.y;
    A += batch * strideA;
    B += batch * strideB;
    C += batch * strideC;

    // accumulate: the accumulators START from C, so a K-split sequence of launches continues one
    // k-ordered chain per element and produces the same bits as a single launch over the whole K
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            acc[i][j][0] = acc[i][j][1] = 0.0;
            if (accumulate) {
                int64_t row = m0 + wm + i * 8 + g, col = n0 + wn + j * 8 + 2 * t;
                if (row < M && col < N) acc[i][j][0] = C[row * ldc + col];
                if (row < M && col + 1 < N) acc[i][j][1] = C[row * ldc + col + 1];
            }
        }

    const int ktiles = (int)((K + BK - 1) / BK);

    auto issue = [&](int kt) {
        if (kt < ktiles) {
            double* sA = smem + (size_t)(kt % STAGES) * STAGE_ELEMS;
            double* sB = sA + OPERAND_ELEMS;
            load_tile<A_KIN, VEC>(sA, A, lda, m0, M, (int64_t)kt * BK, K, tid);
            load_tile<B_KIN, VEC>(sB, B, ldb, n0, N, (int64_t)kt * BK, K, tid);
        }
        cp_async_commit();
    };

#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) issue(s);

    for (int kt = 0; kt < ktiles; kt++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        issue(kt + STAGES - 1);  // refills the slot consumed in iteration kt-1

        const double* sA = smem + (size_t)(kt % STAGES) * STAGE_ELEMS;
        const double* sB = sA + OPERAND_ELEMS;
#pragma unroll
        for (int ks = 0; ks < BK / 4; ks++) {
            double a[8], b[4];
#pragma unroll
            for (int i = 0; i < 8; i++)
                a[i] = A_KIN ? sA[(wm + i * 8 + g) * KIN_PITCH + ks * 4 + t]
                             : sA[(ks * 4 + t) * MNIN_PITCH + wm + i * 8 + g];
#pragma unroll
            for (int j = 0; j < 4; j++)
                b[j] = B_KIN ? sB[(wn + j * 8 + g) * KIN_PITCH + ks * 4 + t]
                             : sB[(ks * 4 + t) * MNIN_PITCH + wn + j * 8 + g];
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: each thread owns 2 adjacent columns per fragment -> 16-byte stores when aligned
    const bool vec_ok = ((ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        int64_t row = m0 + wm + i * 8 + g;
        if (row >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int64_t col = n0 + wn + j * 8 + 2 * t;
            if (col >= N) continue;
            double* p = C + row * ldc + col;
            if (vec_ok && col + 1 < N) {
                *reinterpret_cast<double2*>(p) = make_double2(acc[i][j][0], acc[i][j][1]);
            } else {
                p[0] = acc[i][j][0];
                if (col + 1 < N) p[1] = acc[i][j][1];
            }
        }
    }
}

template <bool A_KIN, bool B_KIN, int VEC>
int launch_variant(const double* A, const double* B, double* C, int64_t batch, int64_t M, int64_t N, int64_t K,
                   int64_t lda, int64_t ldb, int64_t ldc, int64_t stA, int64_t stB, int64_t stC, bool accumulate,
                   cudaStream_t stream) {
    auto kern = dgemm_dmma_kernel<A_KIN, B_KIN, VEC>;
    static thread_local int configured_dev = -1;
    int dev = 0;
    EML_CUDA(cudaGetDevice(&dev));
    if (configured_dev != dev) {
        EML_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        configured_dev = dev;
    }
    int64_t tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
    if (tiles_m * tiles_n > 0x7fffffffLL || batch > 65535) EML_BAD_ARG("dgemm_dmma: grid too large");
    dim3 grid((unsigned)(tiles_m * tiles_n), (unsigned)batch);
    launch_k(kern, dim3(grid), dim3(THREADS), SMEM_BYTES, stream, A, B, C, M, N, K, lda, ldb, ldc, stA, stB, stC, (int)tiles_m,
                                                (int)tiles_n, accumulate ? 1 : 0);
    count_launch();
    set_last_kernel("dmma_f64");
    return check_launch("dgemm_dmma");
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// =================================================================================================
// TMA-staged, warp-specialised variant (the fast path).
//
// One producer lane issues cp.async.bulk.tensor (TMA) loads into a 3-stage ring; eight consumer warps
// wait on the stage's `full` mbarrier, run the DMMA fragments and arrive on its `empty` mbarrier.  No
// CTA-wide barrier in the main loop: the consumer warps never wait for each other.
//
// Shared-memory tiles are laid out by the TMA box itself so that fragment loads stay conflict free
// without padding (TMA writes dense boxes): the tensor map views each operand with a 4-element
// (32-byte) innermost dimension, so the 4 (k or mn) x 4 (mn or k) elements a half-warp reads are 16
// consecutive doubles:
//   K-inner  operand: dims (4, MN, K/4, batch),  box (4, 128, BK/4, 1) -> smem [k/4][mn][k%4]
//   MN-inner operand: dims (4, K, MN/4, batch),  box (4, BK, 32, 1)    -> smem [mn/4][k][mn%4]
// Out-of-range box rows are zero-filled by TMA (ragged M/N/K tails need no special code).
// =================================================================================================
namespace tma {

// Extra destinations of the epilogue: the same C tile is also stored, at the same offsets, into up to 7
// peer GPUs' buffers (peer-mapped pointers, st.global over NVLink) -- the all-gather of a row-sharded GEMM
// fused into the kernel that produces the rows.
struct Remote {
    double* p[7];
    int n;
};

// 8 consumer warps (two warpgroups) + one producer warpgroup (one lane of it works).  The register file is
// re-split with setmaxnreg: a 384-thread CTA compiles to <= 168 registers per thread, the consumers need
// 128 for the accumulators alone, so the producer warpgroup hands its share over (40 vs 232).
constexpr int TBK = 32, TSTAGES = 3, CONSUMER_WARPS = 8, TTHREADS = (CONSUMER_WARPS + 4) * 32;
constexpr int PRODUCER_REGS = 40, CONSUMER_REGS = 232;  // 4*32*40 + 8*32*232 = 64512 <= 65536
constexpr int T_OPERAND_ELEMS = BM * TBK;                      // 4096 doubles = 32 KB
constexpr int T_STAGE_ELEMS = 2 * T_OPERAND_ELEMS;             // 64 KB
constexpr uint32_t T_STAGE_BYTES = T_STAGE_ELEMS * sizeof(double);
constexpr size_t T_SMEM_BYTES = (size_t)TSTAGES * T_STAGE_BYTES + 128 /*align*/ + 64 /*barriers*/;

template <bool A_KIN, bool B_KIN>
__global__ void __launch_bounds__(TTHREADS, 1)
dgemm_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, double* C,
                 int64_t M, int64_t N, int64_t K, int64_t ldc, int64_t strideC, int tiles_m, int tiles_n,
                 int accumulate, const Remote remote) {
    pdl_prologue();
    using namespace sm100;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    // offset (not a cast through an integer) so the compiler keeps the shared address space: LDS, not LD
    uint8_t* base = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    double* tiles = reinterpret_cast<double*>(base);
    uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)TSTAGES * T_STAGE_BYTES);
    uint64_t* empty = full + TSTAGES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    int tm, tn;
    if (tiles_n == 0) {
        // symmetric output (C = G * G^T): only the tiles on and above the diagonal, row by row
        int t = blockIdx.x;
        tm = 0;
        while (t >= tiles_m - tm) {
            t -= tiles_m - tm;
            tm++;
        }
        tn = tm + t;
    } else {
        constexpr int GROUP = 8;
        int tile = blockIdx.x;
        int tiles_per_group = GROUP * tiles_n;
        int group = tile / tiles_per_group;
        int first_m = group * GROUP;
        int group_rows = tiles_m - first_m < GROUP ? tiles_m - first_m : GROUP;
        int in_group = tile % tiles_per_group;
        tm = first_m + in_group % group_rows;
        tn = in_group / group_rows;
    }
    const int batch = blockIdx.y;
    const int ktiles = (int)((K + TBK - 1) / TBK);

    if (tid == 0) {
        prefetch_tensormap(&mapA);
        prefetch_tensormap(&mapB);
        for (int s = 0; s < TSTAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CONSUMER_WARPS);
        }
        fence_barrier_init();
    }
    __syncthreads();

    if (warp >= CONSUMER_WARPS) {
        // ---------------- producer warpgroup ----------------
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PRODUCER_REGS));
        if (warp == CONSUMER_WARPS && lane == 0) {
            for (int kt = 0; kt < ktiles; kt++) {
                const int s = kt % TSTAGES;
                if (kt >= TSTAGES) mbar_wait(&empty[s], ((kt / TSTAGES) - 1) & 1);
                double* sA = tiles + (size_t)s * T_STAGE_ELEMS;
                double* sB = sA + T_OPERAND_ELEMS;
                mbar_expect_tx(&full[s], T_STAGE_BYTES);
                if (A_KIN)
                    tma_load_4d(sA, &mapA, &full[s], 0, tm * BM, kt * (TBK / 4), batch);
                else
                    tma_load_4d(sA, &mapA, &full[s], 0, kt * TBK, tm * (BM / 4), batch);
                if (B_KIN)
                    tma_load_4d(sB, &mapB, &full[s], 0, tn * BN, kt * (TBK / 4), batch);
                else
                    tma_load_4d(sB, &mapB, &full[s], 0, kt * TBK, tn * (BN / 4), batch);
            }
        }
        return;
    }

    // ---------------- consumers ----------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CONSUMER_REGS));
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32;
    const int64_t m0 = (int64_t)tm * BM, n0 = (int64_t)tn * BN;
    C += (int64_t)batch * strideC;

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            acc[i][j][0] = acc[i][j][1] = 0.0;
            if (accumulate) {  // continue the k-ordered chain from C (see the cp.async kernel)
                int64_t row = m0 + wm + i * 8 + g, col = n0 + wn + j * 8 + 2 * t;
                if (row < M && col < N) acc[i][j][0] = C[row * ldc + col];
                if (row < M && col + 1 < N) acc[i][j][1] = C[row * ldc + col + 1];
            }
        }

    // per-thread fragment offset inside an operand tile (fragment 0, ks = 0); fragment i is 8 rows further
    // on, which is a compile-time stride in both layouts, so every LDS uses base + immediate
    const int ra = wm + g, cb = wn + g;
    const int a_off = A_KIN ? ra * 4 + t : (ra >> 2) * (TBK * 4) + t * 4 + (ra & 3);
    const int b_off = B_KIN ? cb * 4 + t : (cb >> 2) * (TBK * 4) + t * 4 + (cb & 3);
    constexpr int A_FR = A_KIN ? 32 : 2 * TBK * 4;  // +8 rows
    constexpr int B_FR = B_KIN ? 32 : 2 * TBK * 4;
    constexpr int A_KS = A_KIN ? BM * 4 : 16;  // offset of the next group of 4 k inside the tile
    constexpr int B_KS = B_KIN ? BN * 4 : 16;

    for (int kt = 0; kt < ktiles; kt++) {
        const int s = kt % TSTAGES;
        mbar_wait(&full[s], (kt / TSTAGES) & 1);
        const double* sA = tiles + (size_t)s * T_STAGE_ELEMS;
        const double* sB = sA + T_OPERAND_ELEMS;
#pragma unroll
        for (int ks = 0; ks < TBK / 4; ks++) {
            double a[8], b[4];
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = sA[a_off + i * A_FR + ks * A_KS];
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = sB[b_off + j * B_FR + ks * B_KS];
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }

    const bool vec_ok = ((ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        int64_t row = m0 + wm + i * 8 + g;
        if (row >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int64_t col = n0 + wn + j * 8 + 2 * t;
            if (col >= N) continue;
            double* p = C + row * ldc + col;
            if (vec_ok && col + 1 < N) {
                *reinterpret_cast<double2*>(p) = make_double2(acc[i][j][0], acc[i][j][1]);
            } else {
                p[0] = acc[i][j][0];
                if (col + 1 < N) p[1] = acc[i][j][1];
            }
        }
    }
    // fused all-gather: the finished tile goes to every peer's copy of C as well (same layout, same ldc)
    for (int r = 0; r < remote.n; r++) {
        double* R = remote.p[r] + (int64_t)batch * strideC;
        const bool rvec = ((ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(R) & 15) == 0);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            int64_t row = m0 + wm + i * 8 + g;
            if (row >= M) continue;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                int64_t col = n0 + wn + j * 8 + 2 * t;
                if (col >= N) continue;
                double* p = R + row * ldc + col;
                if (rvec && col + 1 < N) {
                    *reinterpret_cast<double2*>(p) = make_double2(acc[i][j][0], acc[i][j][1]);
                } else {
                    p[0] = acc[i][j][0];
                    if (col + 1 < N) p[1] = acc[i][j][1];
                }
            }
        }
    }
}

// Tensor map of one operand.  KIN: element (mn, k) at base[mn*ld + k]; else at base[k*ld + mn].
int operand_map(CUtensorMap* map, const double* base, bool kin, int64_t mn, int64_t k, int64_t ld, int64_t batch,
                int64_t batch_stride) {
    uint64_t dims[4], strides[3];
    uint32_t box[4];
    uint64_t bs = (uint64_t)(batch > 1 ? batch_stride : 2) * 8;
    if (kin) {
        dims[0] = 4; dims[1] = (uint64_t)mn; dims[2] = (uint64_t)(k / 4); dims[3] = (uint64_t)batch;
        box[0] = 4; box[1] = BM; box[2] = TBK / 4; box[3] = 1;
    } else {
        dims[0] = 4; dims[1] = (uint64_t)k; dims[2] = (uint64_t)(mn / 4); dims[3] = (uint64_t)batch;
        box[0] = 4; box[1] = TBK; box[2] = BM / 4; box[3] = 1;
    }
    strides[0] = (uint64_t)ld * 8;
    strides[1] = 32;
    strides[2] = bs;
    return sm100::make_tensor_map(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, base, dims, strides, box,
                                  CU_TENSOR_MAP_SWIZZLE_NONE);
}

template <bool A_KIN, bool B_KIN>
int launch(const double* A, const double* B, double* C, int64_t batch, int64_t M, int64_t N, int64_t K, int64_t lda,
           int64_t ldb, int64_t ldc, int64_t stA, int64_t stB, int64_t stC, bool accumulate, cudaStream_t stream,
           bool upper_only, GemmOptions* opt) {
    CUtensorMap mapA, mapB;
    EML_TRY(operand_map(&mapA, A, A_KIN, M, K, lda, batch, stA));
    EML_TRY(operand_map(&mapB, B, B_KIN, N, K, ldb, batch, stB));
    auto kern = dgemm_tma_kernel<A_KIN, B_KIN>;
    static thread_local int configured_dev = -1;
    int dev = 0;
    EML_CUDA(cudaGetDevice(&dev));
    if (configured_dev != dev) {
        EML_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T_SMEM_BYTES));
        configured_dev = dev;
    }
    int64_t tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
    dim3 grid((unsigned)(tiles_m * tiles_n), (unsigned)batch);
    if (upper_only) {  // the kernel reads tiles_n == 0 as "tiles on and above the diagonal only"
        grid.x = (unsigned)(tiles_m * (tiles_m + 1) / 2);
        tiles_n = 0;
        opt->upper_done = true;
    }
    Remote remote{};
    if (opt && opt->n_remote > 0) {  // fused all-gather: tell the caller this kernel delivers the tiles itself
        remote.n = opt->n_remote < 7 ? opt->n_remote : 7;
        for (int i = 0; i < remote.n; i++) remote.p[i] = opt->remote_c[i];
        opt->remote_done = true;
    }
    launch_k(kern, dim3(grid), dim3(TTHREADS), T_SMEM_BYTES, stream, mapA, mapB, C, M, N, K, ldc, stC, (int)tiles_m, (int)tiles_n,
                                                   accumulate ? 1 : 0, remote);
    count_launch();
    set_last_kernel(remote.n ? "dmma_f64_tma+peer_allgather" : "dmma_f64_tma");
    return check_launch("dgemm_tma");
}

}  // namespace tma

}  // namespace

// ------------------------------------------------------------------------------------------------
// FP64 tensor-pipe peak of THIS GPU at its current clocks: register-resident DMMA chains, no memory traffic.
// bench.py uses it as the measured roofline denominator of the f64 GEMM (MEASURED_PEAKS.json has no FP64 row).
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256, 1) dmma_peak_kernel(double* out, int iters) {
    pdl_prologue();
    double acc[32][2];
#pragma unroll
    for (int i = 0; i < 32; i++) acc[i][0] = acc[i][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 32; i++) dmma884(acc[i][0], acc[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) s += acc[i][0] + acc[i][1];
    if (s == 12345.678) out[0] = s;  // keeps the chains alive without a store in the common case
}
}  // namespace

int measure_dmma_peak(DeviceCtx* ctx, double* tflops) {
    const int sms = ctx ? ctx->sm_count : 148, iters = 20000;
    double* d = nullptr;
    EML_CUDA(cudaMalloc(&d, 8));
    cudaEvent_t e0, e1;
    EML_CUDA(cudaEventCreate(&e0));
    EML_CUDA(cudaEventCreate(&e1));
    launch_k(dmma_peak_kernel, dim3(sms), dim3(256), 0, nullptr, d, iters / 10);  // warm-up
    EML_CUDA(cudaEventRecord(e0));
    launch_k(dmma_peak_kernel, dim3(sms), dim3(256), 0, nullptr, d, iters);
    EML_CUDA(cudaEventRecord(e1));
    EML_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    EML_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    count_launch(2);
    // per warp and iteration: 32 DMMA.8x8x4 of 8*8*4 multiply-adds
    *tflops = 2.0 * 256.0 * 32.0 * 8.0 * (double)iters * sms / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return check_launch("dmma_peak");
}

int launch_gemm_dmma_f64(DeviceCtx*, const double* A, const double* B, double* C, int64_t batch, int64_t M,
                         int64_t N, int64_t K, Strides3 sA, Strides3 sB, Strides3 sC, bool accumulate,
                         cudaStream_t stream, bool* handled, GemmOptions* opt) {
    *handled = false;
    if (batch == 0 || M == 0 || N == 0) {
        *handled = true;
        return EML_OK;
    }
    // needs unit stride on one index of each operand and on C's column index; batch fits gridDim.y
    bool a_kin = sA.c == 1, a_min = sA.r == 1;
    bool b_nin = sB.c == 1, b_kin = sB.r == 1;
    if (!(a_kin || a_min) || !(b_nin || b_kin) || sC.c != 1 || batch > 65535) return EML_OK;
    // degenerate extents make both strides "1"-compatible; prefer the natural row-major reading
    bool A_KIN = a_kin;
    bool B_KIN = !b_nin;
    int64_t lda = A_KIN ? sA.r : sA.c;
    int64_t ldb = B_KIN ? sB.c : sB.r;
    int64_t ldc = sC.r;
    bool vec = al16(A) && al16(B) && (lda % 2 == 0) && (ldb % 2 == 0) && (sA.b % 2 == 0) && (sB.b % 2 == 0);
    *handled = true;
    // TMA path: 16-byte aligned rows and the 4-element inner box dimension must tile the contiguous index
    // EML_DGEMM_KERNEL=cpasync selects the cp.async kernel (A/B comparisons in tests and bench); read per call
    const char* forced = getenv("EML_DGEMM_KERNEL");
    const bool force_cpasync = forced && std::string(forced) == "cpasync";
    bool tma_ok = vec && !force_cpasync && al16(C) && ((A_KIN ? K : M) % 4 == 0) && ((B_KIN ? K : N) % 4 == 0);
    // a broadcast operand (batch stride 0) cannot be described to TMA; the cp.async kernel just adds the stride
    if (batch > 1 && (sA.b <= 0 || sB.b <= 0)) tma_ok = false;
    // likewise rows that overlap (leading dimension shorter than the row: a broadcast along m / n / k)
    if (lda < (A_KIN ? K : M) || ldb < (B_KIN ? K : N)) tma_ok = false;
    // C = G * G^T requested by contract_dev: the TMA kernel computes the tiles on and above the diagonal only
    const bool upper = opt && opt->upper_only && tma_ok && M == N && !accumulate && opt->n_remote == 0;
    if (tma_ok) {
        if (A_KIN && !B_KIN)
            return tma::launch<true, false>(A, B, C, batch, M, N, K, lda, ldb, ldc, sA.b, sB.b, sC.b, accumulate, stream, upper, opt);
        if (A_KIN && B_KIN)
            return tma::launch<true, true>(A, B, C, batch, M, N, K, lda, ldb, ldc, sA.b, sB.b, sC.b, accumulate, stream, upper, opt);
        if (!A_KIN && !B_KIN)
            return tma::launch<false, false>(A, B, C, batch, M, N, K, lda, ldb, ldc, sA.b, sB.b, sC.b, accumulate, stream, upper, opt);
        return tma::launch<false, true>(A, B, C, batch, M, N, K, lda, ldb, ldc, sA.b, sB.b, sC.b, accumulate, stream, upper, opt);
    }
#define EML_D(AK, BKK)                                                                                          \
    return vec ? launch_variant<AK, BKK, 2>(A, B, C, batch, M, N, K, lda, ldb, ldc, sA.b, sB.b, sC.b, accumulate, \
                                            stream)                                                              \
               : launch_variant<AK, BKK, 1>(A, B, C, batch, M, N, K, lda, ldb, ldc, sA.b, sB.b, sC.b, accumulate, \
                                            stream);
    if (A_KIN && !B_KIN) { EML_D(true, false) }
    if (A_KIN && B_KIN) { EML_D(true, true) }
    if (!A_KIN && !B_KIN) { EML_D(false, false) }
    EML_D(false, true)
#undef EML_D
}

}  // namespace eml
