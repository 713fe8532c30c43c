// gemm_tf32x3.cu -- f32 GEMM at f32-level accuracy on the 5th-generation tensor cores (tcgen05) for sm_100a.
//
// "3xTF32": every f32 operand x is split into big = tf32(x) and small = tf32(x - big); the product is
//      A*B ~= Asmall*Bbig + Abig*Bsmall + Abig*Bbig        (the dropped small*small term is ~2^-22 relative)
// with f32 accumulation in tensor memory (TMEM).  Products of two TF32 numbers are exact in f32, so the
// result carries f32-level error (checked against the oracle within 1e-5 * sum_k |a_k b_k|).
//
// Pipeline of one call
//   1. split_pack_kernel (HBM-bound): reads an operand once through ANY (batch, mn, k) strides -- so
//      row-major A, row-major B and the transposed reads of the backward products need no separate
//      transpose -- and writes big/small planes in the canonical K-major layout [batch][MN_pad][K_pad],
//      zero padded to whole tiles (ragged edges vanish from the GEMM main loop).
//   2. tf32x3_kernel, one 128 x BN output tile per CTA (BN = 256 or 128), warp-specialised:
//        warp 0      TMA producer: cp.async.bulk.tensor 128B-swizzled boxes of Abig, Asmall, Bbig, Bsmall
//                    into a ring of shared-memory stages (mbarrier full/empty)
//        warp 1      allocates TMEM; one elected lane issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8):
//                    12 MMAs per 32-wide k block (3 products x 4 k steps), k blocks ascending,
//                    tcgen05.commit releases the stage / signals the epilogue
//        warps 2..5  epilogue: tcgen05.ld 32 lanes x 32 columns -> registers -> 128-byte row stores
//                    (optionally C += ...), bounds-checked against the unpadded M, N
//      Optional deterministic split-K (grid.z): each split writes its partial tile to a workspace and
//      launch_splitk_reduce adds the partials in ascending split order (fixed order, no atomics).
//
// Determinism: fixed tile -> CTA map, fixed k order inside a CTA, fixed split order in the reduction.
// Algorithmic work per launch: 2*M*N*K flop (the 3x tensor work is an implementation cost, not counted);
// bytes (M*K + K*N + M*N) * 4.
#include <atomic>
#include <cstdlib>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <tuple>

#include "common.h"
#include "sm100.cuh"
#include "tf32_split.cuh"
#include "ew_rules.cuh"

namespace eml {

namespace {

using namespace sm100;

constexpr int BM = 128, BK = 32;  // BK f32 = 128 bytes = one SWIZZLE_128B row
constexpr int GEMM_THREADS = 192; // 6 warps: producer, mma, 4 epilogue

template <int BN>
struct Cfg {
    static constexpr int A_TILE_BYTES = BM * BK * 4;          // 16 KB
    static constexpr int B_TILE_BYTES = BN * BK * 4;          // 32 KB (BN=256) / 16 KB (BN=128)
    static constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
    static constexpr int STAGES = (BN == 256) ? 2 : 3;        // 192 KB either way
    static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*align*/ + 128 /*barriers*/;
    static constexpr uint32_t TMEM_COLS = BN;                 // f32 accumulator: one column per n
};

// ------------------------------------------------------------------------------------------------
// split + pack: src element (b, mn, k) at src[b*sb + mn*smn + k*sk]  ->  big/small[b][mn][k], padded.
// 32 x 32 tiles through shared memory so both the read (along whichever source index is contiguous)
// and the write (along k) are coalesced.
// ------------------------------------------------------------------------------------------------
// the split itself lives in tf32_split.cuh (shared with the in-kernel converters of gemm_tf32x3_fused.cu)
__device__ __forceinline__ void split_tf32(float x, float& big, float& small) { tf32::split(x, big, small); }

__global__ void __launch_bounds__(256)
split_pack_kernel(const float* __restrict__ src, float* __restrict__ big, float* __restrict__ small, int64_t MN,
                  int64_t K, int64_t MNp, int64_t Kp, int64_t sb, int64_t smn, int64_t sk, int k_fast, unsigned* nonfinite) {
    pdl_prologue();
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const int64_t k0 = (int64_t)blockIdx.x * 32, mn0 = (int64_t)blockIdx.y * 32, b = blockIdx.z;
    const float* s = src + b * sb;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        int row = ty + r * 8;
        // k_fast: tx walks k (source contiguous along k); else tx walks mn
        int64_t mn = mn0 + (k_fast ? row : tx), k = k0 + (k_fast ? tx : row);
        float v = (mn < MN && k < K) ? s[mn * smn + k * sk] : 0.0f;
        if (k_fast)
            tile[row][tx] = v;  // [mn][k]
        else
            tile[tx][row] = v;  // [mn][k]
    }
    __syncthreads();
    const int64_t plane = MNp * Kp;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        int row = ty + r * 8;
        int64_t mn = mn0 + row, k = k0 + tx;  // Kp, MNp are multiples of 32: always in range
        float v = tile[row][tx];
        float hi, lo;
        split_tf32(v, hi, lo);
        if (!tf32::ordinary(v)) *nonfinite = 1u;  // (rare; every writer stores the same value)
        int64_t o = b * plane + mn * Kp + k;
        big[o] = hi;
        small[o] = lo;
    }
}


// Vectorised split + pack: 64 (mn) x 64 (k) tiles, 128-bit global loads along the source's contiguous index
// and 128-bit stores along k.  Requires 16-byte aligned rows in the source (the dispatcher checks and falls
// back to split_pack_kernel otherwise).  K_FAST: the source is contiguous along k (no transpose needed, no
// shared memory); otherwise along mn (transpose through shared memory).
__device__ __forceinline__ float4 load4_guarded(const float* p, int64_t first, int64_t extent) {
    // 4 consecutive elements starting at index `first` of a contiguous run of `extent` (zero past the end)
    if (first + 3 < extent) return *reinterpret_cast<const float4*>(p);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (first < extent) v.x = p[0];
    if (first + 1 < extent) v.y = p[1];
    if (first + 2 < extent) v.z = p[2];
    return v;
}
__device__ __forceinline__ void store_split4(float* big, float* small, float4 v, unsigned* nonfinite) {
    float4 hi, lo;
    if (tf32::ordinary4(v)) {
        tf32::split4_ordinary(v, hi, lo);
    } else {
        tf32::split4_any(v, hi, lo);
        *nonfinite = 1u;  // (rare; every writer stores the same value)
    }
    *reinterpret_cast<float4*>(big) = hi;
    *reinterpret_cast<float4*>(small) = lo;
}

template <bool K_FAST>
__global__ void __launch_bounds__(256)
split_pack_vec_kernel(const float* __restrict__ src, float* __restrict__ big, float* __restrict__ small, int64_t MN,
                      int64_t K, int64_t MNp, int64_t Kp, int64_t sb, int64_t s_slow, unsigned* nonfinite) {
    pdl_prologue();
    const int tid = threadIdx.x;
    const int64_t k0 = (int64_t)blockIdx.x * 64, mn0 = (int64_t)blockIdx.y * 64, b = blockIdx.z;
    const float* s = src + b * sb;
    float* ob = big + b * MNp * Kp;
    float* os = small + b * MNp * Kp;
    const int c4 = (tid & 15) * 4, r = tid >> 4;  // 16 float4 columns x 16 rows per pass
    if (K_FAST) {
        // element (mn, k) at s[mn * s_slow + k]
#pragma unroll
        for (int pass = 0; pass < 4; pass++) {
            const int64_t mn = mn0 + r + 16 * pass, k = k0 + c4;
            if (mn >= MNp || k >= Kp) continue;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (mn < MN) v = load4_guarded(s + mn * s_slow + k, k, K);
            store_split4(ob + mn * Kp + k, os + mn * Kp + k, v, nonfinite);
        }
    } else {
        // element (mn, k) at s[k * s_slow + mn]: read rows of k, write rows of mn
        __shared__ float tile[64][65];
#pragma unroll
        for (int pass = 0; pass < 4; pass++) {
            const int kk = r + 16 * pass;
            const int64_t k = k0 + kk, mn = mn0 + c4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < K) v = load4_guarded(s + k * s_slow + mn, mn, MN);
            tile[kk][c4] = v.x;
            tile[kk][c4 + 1] = v.y;
            tile[kk][c4 + 2] = v.z;
            tile[kk][c4 + 3] = v.w;
        }
        __syncthreads();
#pragma unroll
        for (int pass = 0; pass < 4; pass++) {
            const int m = r + 16 * pass;
            const int64_t mn = mn0 + m, k = k0 + c4;
            if (mn >= MNp || k >= Kp) continue;
            float4 v = make_float4(tile[c4][m], tile[c4 + 1][m], tile[c4 + 2][m], tile[c4 + 3][m]);
            store_split4(ob + mn * Kp + k, os + mn * Kp + k, v, nonfinite);
        }
    }
}

// one operand: vector kernel when the source allows 128-bit loads, the scalar tile kernel otherwise
int launch_split_pack(const float* src, float* big, float* small, int64_t batch, int64_t MN, int64_t K, int64_t MNp,
                      int64_t Kp, int64_t sb, int64_t smn, int64_t sk, unsigned* nonfinite, cudaStream_t stream) {
    if (MNp / 32 > 65535 || batch > 65535) EML_BAD_ARG("tf32x3: operand too large for the pack grid");
    const bool al = (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (sb & 3) == 0;
    dim3 vgrid((unsigned)((Kp + 63) / 64), (unsigned)((MNp + 63) / 64), (unsigned)batch);
    if (al && sk == 1 && (smn & 3) == 0) {
        launch_k(split_pack_vec_kernel<true>, dim3(vgrid), dim3(256), 0, stream, src, big, small, MN, K, MNp, Kp, sb, smn, nonfinite);
    } else if (al && smn == 1 && (sk & 3) == 0) {
        launch_k(split_pack_vec_kernel<false>, dim3(vgrid), dim3(256), 0, stream, src, big, small, MN, K, MNp, Kp, sb, sk, nonfinite);
    } else {
        dim3 grid((unsigned)(Kp / 32), (unsigned)(MNp / 32), (unsigned)batch);
        launch_k(split_pack_kernel, dim3(grid), dim3(256), 0, stream, src, big, small, MN, K, MNp, Kp, sb, smn, sk, sk == 1, nonfinite);
    }
    count_launch();
    return check_launch("tf32x3 split/pack");
}

// ------------------------------------------------------------------------------------------------
// the GEMM
// ------------------------------------------------------------------------------------------------
struct Maps {
    CUtensorMap a_big, a_small, b_big, b_small;
};

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
tf32x3_kernel(const __grid_constant__ Maps maps, float* __restrict__ C, int64_t M, int64_t N, int64_t ldc,
              int64_t strideC, int kblocks_total, int kblocks_per_split, int accumulate, float* __restrict__ ws,
              int64_t ws_split_stride) {
    pdl_prologue();
    using cfg = Cfg<BN>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)cfg::STAGES * cfg::STAGE_BYTES);
    uint64_t* empty = full + cfg::STAGES;
    uint64_t* acc_ready = empty + cfg::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_ready + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // tile raster: groups of 16 tile rows walked column by column, so one wave of 148 CTAs touches a
    // near-square set of A and B panels (DRAM traffic per wave ~ its perimeter)
    constexpr int GROUP = 16;
    const int tiles_m = (int)((M + BM - 1) / BM), tiles_n = (int)((N + BN - 1) / BN);
    const int tiles_per_group = GROUP * tiles_n;
    const int first_m = ((int)blockIdx.x / tiles_per_group) * GROUP;
    const int group_rows = tiles_m - first_m < GROUP ? tiles_m - first_m : GROUP;
    const int in_group = (int)blockIdx.x % tiles_per_group;
    const int tm = first_m + in_group % group_rows;
    const int tn = in_group / group_rows;
    const int batch = blockIdx.y, split = blockIdx.z;
    const int kb0 = split * kblocks_per_split;
    const int kb1 = min(kblocks_total, kb0 + kblocks_per_split);
    const int nkb = kb1 - kb0;  // >= 1 by construction of the grid

    if (threadIdx.x == 0) {
        prefetch_tensormap(&maps.a_big);
        prefetch_tensormap(&maps.a_small);
        prefetch_tensormap(&maps.b_big);
        prefetch_tensormap(&maps.b_small);
        for (int s = 0; s < cfg::STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(acc_ready, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        if (lane == 0) {
            for (int i = 0; i < nkb; i++) {
                const int s = i % cfg::STAGES;
                if (i >= cfg::STAGES) mbar_wait(&empty[s], ((i / cfg::STAGES) - 1) & 1);
                uint8_t* st = base + (size_t)s * cfg::STAGE_BYTES;
                mbar_expect_tx(&full[s], cfg::STAGE_BYTES);
                const int kc = (kb0 + i) * BK;
                tma_load_3d(st, &maps.a_big, &full[s], kc, tm * BM, batch);
                tma_load_3d(st + cfg::A_TILE_BYTES, &maps.a_small, &full[s], kc, tm * BM, batch);
                tma_load_3d(st + 2 * cfg::A_TILE_BYTES, &maps.b_big, &full[s], kc, tn * BN, batch);
                tma_load_3d(st + 2 * cfg::A_TILE_BYTES + cfg::B_TILE_BYTES, &maps.b_small, &full[s], kc, tn * BN, batch);
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(BM, BN);
            for (int i = 0; i < nkb; i++) {
                const int s = i % cfg::STAGES;
                mbar_wait(&full[s], (i / cfg::STAGES) & 1);
                tc_fence_after();
                const uint32_t st = smem_u32(base + (size_t)s * cfg::STAGE_BYTES);
                const uint64_t a_big = umma_desc_k_sw128(st);
                const uint64_t a_small = umma_desc_k_sw128(st + cfg::A_TILE_BYTES);
                const uint64_t b_big = umma_desc_k_sw128(st + 2 * cfg::A_TILE_BYTES);
                const uint64_t b_small = umma_desc_k_sw128(st + 2 * cfg::A_TILE_BYTES + cfg::B_TILE_BYTES);
                // small terms first, then the dominant one; 8 tf32 = 32 bytes per k step (>>4 = 2)
#pragma unroll
                for (int k = 0; k < BK / 8; k++) umma_tf32(tmem_acc, a_small + 2 * k, b_big + 2 * k, idesc, (i | k) != 0);
#pragma unroll
                for (int k = 0; k < BK / 8; k++) umma_tf32(tmem_acc, a_big + 2 * k, b_small + 2 * k, idesc, 1);
#pragma unroll
                for (int k = 0; k < BK / 8; k++) umma_tf32(tmem_acc, a_big + 2 * k, b_big + 2 * k, idesc, 1);
                umma_commit(&empty[s]);  // the stage is free once these MMAs have read it
            }
            umma_commit(acc_ready);
        }
    } else {
        // ---------------- epilogue: TMEM -> registers -> global ----------------
        const int q = warp & 3;  // TMEM lane quarter this warp may read
        mbar_wait(acc_ready, 0);
        tc_fence_after();
        const int64_t row = (int64_t)tm * BM + q * 32 + lane;
        const int64_t n0 = (int64_t)tn * BN;
        float* out;
        int64_t ld;
        if (ws) {  // split-K partial: dense [split][batch][M][N]
            out = ws + (int64_t)split * ws_split_stride + (int64_t)batch * M * N;
            ld = N;
        } else {
            out = C + (int64_t)batch * strideC;
            ld = ldc;
        }
        const bool add = accumulate && !ws;
        const bool vec_ok = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
#pragma unroll 1
        for (int c = 0; c < BN / 32; c++) {
            uint32_t v[32];
            const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);
            tmem_ld_32x32b_x32(taddr, v);
            tmem_ld_wait();
            const int64_t col0 = n0 + c * 32;
            if (row < M && col0 < N) {
                float* p = out + row * ld + col0;
                if (vec_ok && col0 + 32 <= N) {
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        float4 o = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                               __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                        if (add) {
                            float4 old = *reinterpret_cast<const float4*>(p + 4 * j);
                            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                        }
                        *reinterpret_cast<float4*>(p + 4 * j) = o;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j++)
                        if (col0 + j < N) p[j] = add ? p[j] + __uint_as_float(v[j]) : __uint_as_float(v[j]);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_acc, cfg::TMEM_COLS);
}


// =================================================================================================
// CTA-pair, persistent variant (the fast path for M > 128 and N > 128).
//
// The single-CTA kernel above is shared-memory-bandwidth bound: each 128x256x8 MMA reads 12 KB of operands
// from shared memory in 128 cycles (96 B/clk) while TMA writes the next stage (62 B/clk) -- more than the
// 128 B/clk an SM has, which caps the tensor pipe near 80 % (ncu: 79 %).  Here two CTAs of a cluster form a
// pair (tcgen05 cta_group::2) and compute a 256 x 256 tile together: each CTA stages only ITS 128 rows of A
// and ITS 128-column half of B (64 KB per k block instead of 96 KB), the pair's tensor cores read both halves
// of B across the pair, so per SM the MMA reads 8 KB per 128 cycles and the ring holds 3 stages.
//
// Persistent: one cluster per SM pair walks a static list of work items (tile x split), the f32 accumulator
// is double buffered in TMEM (2 x 256 columns), so the epilogue of item i overlaps the MMAs of item i+1.
//   warp 0 (both CTAs)   TMA producer: own halves into own shared memory, completion on the LEADER's barrier
//   warp 1 (leader CTA)  MMA issuer: tcgen05.mma.cta_group::2 (M=256, N=256, K=8), commits multicast to both CTAs
//   warp 1 (both)        TMEM allocation / release (cta_group::2)
//   warps 2..5 (both)    epilogue: tcgen05.ld -> registers -> global; release the accumulator stage to the leader
// =================================================================================================
namespace pair {

constexpr int PM = 256, PN = 256;           // tile of one CTA pair
constexpr int HALF = 128;                   // rows of A / columns of B staged by one CTA
constexpr int TILE_BYTES = HALF * BK * 4;   // 16 KB
constexpr int STAGE_BYTES = 4 * TILE_BYTES; // A big, A small, B big, B small: 64 KB
constexpr int GROUP_BYTES = 32 * BK * 4;    // MN-major tile: one group of 32 rows x 32 k
constexpr int STAGES = 3;
constexpr int ACC_STAGES = 2;
constexpr uint32_t TMEM_COLS = 512;         // 2 accumulator stages x 256 columns
constexpr int EPI_WARPS = 8;                 // warp = (TMEM lane quarter, column half)
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int THREADS = (4 + EPI_WARPS) * 32;  // warpgroup 0: producer, MMA issuer, 2 idle; warpgroups 1-2: epilogue
// 384 threads compile to <= 168 registers; the epilogue keeps 128 sums + 32 loaded values per thread, so the
// register file is re-split with setmaxnreg: 4*32*40 + 8*32*232 = 64512 <= 65536
constexpr int CONTROL_REGS = 40, EPILOGUE_REGS = 232;
// The tensor core adds into its f32 accumulator with truncation (round toward zero): over a long chain of
// same-sign products the bias grows with the number of MMAs (measured: 5e-4 relative at K = 65536).  So an
// accumulator is only trusted for FLUSH_KB k blocks; then the MMA warp moves to the other TMEM stage and the epilogue
// warps add the drained partial into a running f32 sum held in REGISTERS (128 columns of one row per thread) with
// IEEE round-to-nearest adds, under the next chunk's MMAs; C is written once per tile.
// Measured bias on all-positive data (scripts/err_probe.py), relative to the exact result: with the truncating big /
// small split of tf32_split.cuh -- every one of the three product terms has the sign of a b, so nothing cancels --
// 2.2e-6 per 128 k of chain; 8 k blocks = 256 k keep it at 4.4e-6 (max 5.8e-6 over a 2048^3 product), inside the
// 1e-5 sum|a b| contract with room for the split's own 1.9e-6.  (Round 1 split by rounding, which halves the bias
// per k, and flushed every 16 k blocks: the same figures.)
constexpr int FLUSH_KB = 8;

struct Work {
    int tm, tn, batch, split;
};

// The non-finite repair (see tf32_repair_kernel below) folded into this kernel's epilogue when the launch writes C itself
// (no split-K workspace, no accumulation): every CTA reads the operands' "non-finite input seen" words at its start and
// draws a ticket (the last one clears the stream's shared word, as the separate kernel does); when a word was set, each
// epilogue thread re-reads the part of the row it has just stored and recomputes the NaNs as f32 fma chains from the
// unsplit operands.  Saves the separate launch that otherwise follows every product.
struct RepairArgs {
    const float* A = nullptr;  // null: not folded
    const float* B = nullptr;
    int64_t K = 0;
    int64_t sAb = 0, sAr = 0, sAc = 0, sBb = 0, sBr = 0, sBc = 0;
    const unsigned* flag_a = nullptr;
    const unsigned* flag_b = nullptr;
    unsigned* stream_flag = nullptr;
    unsigned* ticket = nullptr;
};

// tiles_n == 0: symmetric output -- only the tiles on and above the diagonal of a tiles_m x tiles_m grid, row by row
__device__ __forceinline__ Work decode(int w, int tiles_m, int tiles_n, int splits) {
    Work k;
    k.split = w % splits;
    int t = w / splits;
    if (tiles_n == 0) {
        const int tri = tiles_m * (tiles_m + 1) / 2;
        k.batch = t / tri;
        t %= tri;
        int tm = 0;
        while (t >= tiles_m - tm) {
            t -= tiles_m - tm;
            tm++;
        }
        k.tm = tm;
        k.tn = tm + t;
        return k;
    }
    const int tiles = tiles_m * tiles_n;
    k.batch = t / tiles;
    t %= tiles;
    // groups of 8 pair rows (2048 rows of C) walked column by column
    constexpr int GROUP = 8;
    const int per_group = GROUP * tiles_n;
    const int first_m = (t / per_group) * GROUP;
    const int rows = tiles_m - first_m < GROUP ? tiles_m - first_m : GROUP;
    const int in_group = t % per_group;
    k.tm = first_m + in_group % rows;
    k.tn = in_group / rows;
    return k;
}

// EPI: 0 = plain; 1 = the sigmoid chain (neg, exp, + c, c / x) applied to the finished tile in the epilogue (EpilogueChain)
// A_KIN / B_KIN: the operand's tiles are K-major ([128 rows][32 k], SWIZZLE_128B: packed planes, or a k-contiguous
// matrix read in place) or MN-major ([4 groups][32 k][32 rows], 128-byte swizzle with 32-byte atoms: an m / n-contiguous
// matrix read in place -- see sm100.cuh).  "big" may be the RAW f32 matrix: kind::tf32 truncates the low 13 bits itself.
template <int EPI, bool A_KIN, bool B_KIN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
tf32x3_pair_kernel(const __grid_constant__ Maps maps, float* __restrict__ C, int64_t M, int64_t N, int64_t ldc,
                   int64_t strideC, int kblocks_total, int kblocks_per_split, int splits, int tiles_m, int tiles_n,
                   int work_items, int accumulate, float* __restrict__ ws, int64_t ws_split_stride, const RepairArgs rep,
                   const EpilogueChain epi, int tile_w) {
    // programmatic dependent launch: barriers, TMEM and the descriptor prefetch are set up while the predecessor is still
    // running; nothing it wrote is read before pdl_wait() below
    pdl_launch_dependents();
    __shared__ unsigned repair_seen;
#ifdef EML_PAIR_TIMING
    __shared__ long long ts[16];
    const long long t_entry = clock64();
#define TS(i) do { if (blockIdx.x == EML_PAIR_TIMING) ts[i] = clock64() - t_entry; } while (0)
#else
#define TS(i) do { } while (0)
#endif
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* acc_full = empty + STAGES;
    uint64_t* acc_empty = acc_full + ACC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + ACC_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();  // 0 = leader
    const int cluster = blockIdx.x >> 1, clusters = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        prefetch_tensormap(&maps.a_big);
        prefetch_tensormap(&maps.a_small);
        prefetch_tensormap(&maps.b_big);
        prefetch_tensormap(&maps.b_small);
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full[s], 1);   // leader's copy is the one in use: its producer arrives, both CTAs' bytes land
            mbar_init(&empty[s], 1);  // one multicast commit per use
        }
        for (int a = 0; a < ACC_STAGES; a++) {
            mbar_init(&acc_full[a], 1);               // one multicast commit per work item
            mbar_init(&acc_empty[a], 2 * EPI_WARPS);  // leader's copy: every epilogue warp of both CTAs
        }
        fence_barrier_init();
    }
    __syncwarp();
    if (warp == 1) tmem_alloc_pair(tmem_slot, TMEM_COLS);
    if (threadIdx.x == 0) TS(0);
    pdl_wait();  // the predecessor has completed and flushed: global memory may be read from here on
    if (threadIdx.x == 0) TS(1);
    tc_fence_before();
    cluster_sync_all();  // barriers of BOTH CTAs are initialised before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;
    if (threadIdx.x == 0) TS(2);

    if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(CONTROL_REGS));
    if (warp == 0) {
        // ---------------- TMA producer (both CTAs) ----------------
        if (lane == 0) {
            int n = 0;
            for (int w = cluster; w < work_items; w += clusters) {
                const Work wk = decode(w, tiles_m, tiles_n, splits);
                const int kb0 = wk.split * kblocks_per_split;
                const int kb1 = min(kblocks_total, kb0 + kblocks_per_split);
                // (a tile of tile_w <= 256 columns: this CTA's half of B starts tile_w / 2 rows into it; the box still brings
                //  128 rows, of which the MMA reads the first tile_w / 2)
                const int row_a = wk.tm * PM + (int)rank * HALF, row_b = wk.tn * tile_w + (int)rank * (tile_w >> 1);
                for (int kb = kb0; kb < kb1; kb++, n++) {
                    const int s = n % STAGES;
                    if (n >= STAGES) mbar_wait(&empty[s], ((n / STAGES) - 1) & 1);
                    uint8_t* st = base + (size_t)s * STAGE_BYTES;
                    if (rank == 0) mbar_expect_tx(&full[s], 2 * STAGE_BYTES);
                    const int kc = kb * BK;
                    if (A_KIN) {
                        tma_load_3d_pair(st, &maps.a_big, &full[s], kc, row_a, wk.batch);
                        tma_load_3d_pair(st + TILE_BYTES, &maps.a_small, &full[s], kc, row_a, wk.batch);
                    } else {
#pragma unroll
                        for (int g = 0; g < 4; g++) {
                            tma_load_3d_pair(st + g * GROUP_BYTES, &maps.a_big, &full[s], row_a + 32 * g, kc, wk.batch);
                            tma_load_3d_pair(st + TILE_BYTES + g * GROUP_BYTES, &maps.a_small, &full[s], row_a + 32 * g, kc, wk.batch);
                        }
                    }
                    if (B_KIN) {
                        tma_load_3d_pair(st + 2 * TILE_BYTES, &maps.b_big, &full[s], kc, row_b, wk.batch);
                        tma_load_3d_pair(st + 3 * TILE_BYTES, &maps.b_small, &full[s], kc, row_b, wk.batch);
                    } else {
#pragma unroll
                        for (int g = 0; g < 4; g++) {
                            tma_load_3d_pair(st + 2 * TILE_BYTES + g * GROUP_BYTES, &maps.b_big, &full[s], row_b + 32 * g, kc, wk.batch);
                            tma_load_3d_pair(st + 3 * TILE_BYTES + g * GROUP_BYTES, &maps.b_small, &full[s], row_b + 32 * g, kc, wk.batch);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (leader CTA only) ----------------
        if (rank == 0 && lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(PM, tile_w) | (A_KIN ? 0u : UMMA_IDESC_A_MN) | (B_KIN ? 0u : UMMA_IDESC_B_MN);
            int n = 0, it = 0;  // it counts accumulator uses (one per k chunk)
            for (int w = cluster; w < work_items; w += clusters) {
                const Work wk = decode(w, tiles_m, tiles_n, splits);
                const int kb0 = wk.split * kblocks_per_split;
                const int kb1 = min(kblocks_total, kb0 + kblocks_per_split);
                for (int c0 = kb0; c0 < kb1; c0 += FLUSH_KB, it++) {
                    const int c1 = min(kb1, c0 + FLUSH_KB);
                    const int as = it & 1;
                    if (it >= ACC_STAGES) mbar_wait(&acc_empty[as], ((it >> 1) - 1) & 1);  // epilogue drained this stage
                    tc_fence_after();
                    const uint32_t d = tmem_acc + (uint32_t)as * PN;
                    for (int kb = c0; kb < c1; kb++, n++) {
                        const int s = n % STAGES;
                        mbar_wait(&full[s], (n / STAGES) & 1);
                        if (n == 0) TS(3);
                        if (n == 3) TS(4);
                        tc_fence_after();
                        const uint32_t st = smem_u32(base + (size_t)s * STAGE_BYTES);
                        const uint32_t a_big = st, a_small = st + TILE_BYTES, b_big = st + 2 * TILE_BYTES, b_small = st + 3 * TILE_BYTES;
                        const uint32_t first = kb != c0;
#pragma unroll
                        for (int k = 0; k < BK / 8; k++)
                            umma_tf32_pair(d, umma_tile_desc_f32<A_KIN>(a_small, k), umma_tile_desc_f32<B_KIN>(b_big, k), idesc,
                                           first | (uint32_t)k);
#pragma unroll
                        for (int k = 0; k < BK / 8; k++)
                            umma_tf32_pair(d, umma_tile_desc_f32<A_KIN>(a_big, k), umma_tile_desc_f32<B_KIN>(b_small, k), idesc, 1);
#pragma unroll
                        for (int k = 0; k < BK / 8; k++)
                            umma_tf32_pair(d, umma_tile_desc_f32<A_KIN>(a_big, k), umma_tile_desc_f32<B_KIN>(b_big, k), idesc, 1);
                        umma_commit_pair(&empty[s], 0b11);  // frees the stage in both CTAs
                    }
                    umma_commit_pair(&acc_full[as], 0b11);  // this chunk's partial is complete: wake both epilogues
                    TS(5);
                }
            }
        }
    } else if (warp >= 4) {
        // ---------------- epilogue (both CTAs) ----------------
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(EPILOGUE_REGS));
        const int q = warp & 3, half = (warp - 4) >> 2;  // TMEM lane quarter, 128-column half
        if (rep.A) {
            // the operands' "non-finite input seen" words and this CTA's ticket, read while the first MMAs run (three
            // dependent global loads and an atomic: ~0.8 us that used to sit in front of the cluster's first barrier);
            // the epilogue warps meet on a barrier of their own before anybody looks at the result
            if (threadIdx.x == 128) {
                repair_seen = ((rep.flag_a && *rep.flag_a) || (rep.flag_b && *rep.flag_b) || *rep.stream_flag) ? 1u : 0u;
                __threadfence();
                if (atomicInc(rep.ticket, gridDim.x - 1) == gridDim.x - 1) *rep.stream_flag = 0u;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
        }
        int it = 0;
        for (int w = cluster; w < work_items; w += clusters) {
            const Work wk = decode(w, tiles_m, tiles_n, splits);
            const int kb0 = wk.split * kblocks_per_split;
            const int kb1 = min(kblocks_total, kb0 + kblocks_per_split);
            float sum[128];  // this thread's row, 128 columns: the IEEE-rounded running sum over k chunks
            for (int c0 = kb0; c0 < kb1; c0 += FLUSH_KB, it++) {
                const int as = it & 1;
                mbar_wait(&acc_full[as], (it >> 1) & 1);
                if (threadIdx.x == 128) TS(6);
                tc_fence_after();
                const uint32_t t0 = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * PN + half * 128);
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(t0 + (uint32_t)(c * 32), v);
                    tmem_ld_wait();
                    if (c0 == kb0) {
#pragma unroll
                        for (int j = 0; j < 32; j++) sum[c * 32 + j] = __uint_as_float(v[j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j++) sum[c * 32 + j] += __uint_as_float(v[j]);
                    }
                }
                tc_fence_before();  // the chunk is in registers: hand the accumulator stage back to the leader
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(&acc_empty[as], 0);
                if (threadIdx.x == 128) TS(7);
            }
            // one store of the tile
            const int64_t row = (int64_t)wk.tm * PM + (int64_t)rank * HALF + q * 32 + lane;
            const int64_t n0 = (int64_t)wk.tn * tile_w + half * 128;
            // columns of this 128-column half that belong to the tile and to the matrix
            const int64_t n_end = min((int64_t)(wk.tn + 1) * tile_w, N);
            float* out;
            int64_t ld;
            if (ws) {
                out = ws + (int64_t)wk.split * ws_split_stride + (int64_t)wk.batch * M * N;
                ld = N;
            } else {
                out = C + (int64_t)wk.batch * strideC;
                ld = ldc;
            }
            const bool add = accumulate && !ws;
            const bool vec_ok = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
            if (row < M) {
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const int64_t col0 = n0 + c * 32;
                    if (col0 >= n_end) continue;
                    float* p = out + row * ld + col0;
                    if (vec_ok && col0 + 32 <= n_end) {
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            float4 o = make_float4(sum[c * 32 + 4 * j], sum[c * 32 + 4 * j + 1], sum[c * 32 + 4 * j + 2],
                                                   sum[c * 32 + 4 * j + 3]);
                            if (add) {
                                float4 old = *reinterpret_cast<const float4*>(p + 4 * j);
                                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                            }
                            *reinterpret_cast<float4*>(p + 4 * j) = o;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j++)
                            if (col0 + j < n_end) p[j] = add ? p[j] + sum[c * 32 + j] : sum[c * 32 + j];
                    }
                }
            }
            if (threadIdx.x == 128) TS(8);
            if (row < M) {
                if (rep.A && repair_seen) {  // rare: an operand held an Inf / NaN / >= 2^127 value
                    const float* a = rep.A + (int64_t)wk.batch * rep.sAb + row * rep.sAr;
                    for (int64_t col = n0; col < n0 + 128 && col < n_end; col++) {
                        float* pc = out + row * ld + col;
                        if (*pc == *pc) continue;  // only NaNs can be artefacts of the split
                        const float* bb = rep.B + (int64_t)wk.batch * rep.sBb + col * rep.sBc;
                        float acc = 0.0f;
                        for (int64_t k = 0; k < rep.K; k++) acc = fmaf(a[k * rep.sAc], bb[k * rep.sBr], acc);
                        *pc = acc;
                    }
                }
                if constexpr (EPI == 1) {
                    // the elementwise chain on the tile just stored (one batch, dense outputs with row stride N)
                    const bool reread = rep.A && repair_seen;  // (repaired elements live in C, not in the registers)
                    const bool ovec = (N & 3) == 0;
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const int64_t col0 = n0 + c * 32;
                        if (col0 >= N) continue;
                        const float* zrow = out + row * ld + col0;
#pragma unroll
                        for (int j4 = 0; j4 < 8; j4++) {
                            float y[4][4];  // [op][lane of the vector]
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                const int j = 4 * j4 + e;
                                float x = sum[c * 32 + j], d;
                                if (reread && col0 + j < N) x = zrow[j];
                                unary_rule<float, EML_OP_NEG>(epi.c[0], x, y[0][e], d);
                                unary_rule<float, EML_OP_EXP>(epi.c[1], y[0][e], y[1][e], d);
                                unary_rule<float, EML_OP_ADD_C>(epi.c[2], y[1][e], y[2][e], d);
                                unary_rule<float, EML_OP_C_DIV>(epi.c[3], y[2][e], y[3][e], d);
                            }
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                if (!epi.out[k]) continue;
                                float* q = epi.out[k] + row * N + col0 + 4 * j4;
                                if (ovec && col0 + 4 * j4 + 4 <= N) {
                                    *reinterpret_cast<float4*>(q) = make_float4(y[k][0], y[k][1], y[k][2], y[k][3]);
                                } else {
#pragma unroll
                                    for (int e = 0; e < 4; e++)
                                        if (col0 + 4 * j4 + e < N) q[e] = y[k][e];
                                }
                            }
                        }
                    }
                }
            }
        }
    }

    __syncwarp();  // the elected lanes rejoin their warps: the cluster barrier below is warp-aligned
    tc_fence_before();
    cluster_sync_all();  // no CTA frees TMEM or exits while its peer may still address it
    if (warp == 1) tmem_dealloc_pair(tmem_acc, TMEM_COLS);
#ifdef EML_PAIR_TIMING
    if (threadIdx.x == 0 && blockIdx.x == EML_PAIR_TIMING)
        printf("pair timing kb=%d: setup %lld wait %lld sync %lld | first full %lld 4th full %lld last commit %lld | last acc_full %lld drained %lld stored %lld | end %lld\n",
               kblocks_per_split, ts[0], ts[1], ts[2], ts[3], ts[4], ts[5], ts[6], ts[7], ts[8], clock64() - t_entry);
#endif
#undef TS
}

int launch(const Maps& maps, float* C, int64_t batch, int64_t M, int64_t N, int64_t ldc, int64_t strideC, int kblocks,
           int splits, int sms, int accumulate /* ACC_*; ACC_ZERO_PLUS only with ws */, float* ws, cudaStream_t stream,
           bool upper_only = false, const RepairArgs* repair = nullptr, const EpilogueChain* epi = nullptr, bool a_kin = true,
           bool b_kin = true, bool skip_reduce = false, int tile_w = PN) {
    // tile_w: columns of a work item's tile (a multiple of 16, <= 256)
    static thread_local int configured_dev = -1;
    int dev = 0;
    EML_CUDA(cudaGetDevice(&dev));
    using Kernel = void (*)(const Maps, float*, int64_t, int64_t, int64_t, int64_t, int, int, int, int, int, int, int, float*, int64_t,
                            const RepairArgs, const EpilogueChain, int);
    static const Kernel kernels[2][2][2] = {
        {{tf32x3_pair_kernel<0, false, false>, tf32x3_pair_kernel<0, false, true>},
         {tf32x3_pair_kernel<0, true, false>, tf32x3_pair_kernel<0, true, true>}},
        {{tf32x3_pair_kernel<1, false, false>, tf32x3_pair_kernel<1, false, true>},
         {tf32x3_pair_kernel<1, true, false>, tf32x3_pair_kernel<1, true, true>}}};
    if (configured_dev != dev) {
        for (int e = 0; e < 2; e++)
            for (int a = 0; a < 2; a++)
                for (int b = 0; b < 2; b++)
                    EML_CUDA(cudaFuncSetAttribute(kernels[e][a][b], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        configured_dev = dev;
    }
    if (upper_only || epi) tile_w = PN;  // (the host never combines these with narrow tiles)
    const int tiles_m = (int)((M + PM - 1) / PM), tiles_n = upper_only ? 0 : (int)((N + tile_w - 1) / tile_w);
    const int per_split = (kblocks + splits - 1) / splits;
    const int used_splits = (kblocks + per_split - 1) / per_split;
    const int64_t tiles = upper_only ? (int64_t)tiles_m * (tiles_m + 1) / 2 : (int64_t)tiles_m * tiles_n;
    const int64_t items = (int64_t)batch * tiles * used_splits;
    if (items > 0x7fffffffLL) EML_BAD_ARG("tf32x3: too many work items");
    int clusters = sms / 2;
    if (items < clusters) clusters = (int)items;
    launch_k(kernels[epi ? 1 : 0][a_kin ? 1 : 0][b_kin ? 1 : 0], dim3(2 * clusters), dim3(THREADS), SMEM_BYTES, stream, maps, C, M, N, ldc,
             strideC, kblocks, per_split, used_splits, tiles_m, tiles_n, (int)items, accumulate == ACC_ADD ? 1 : 0, ws, batch * M * N,
             repair ? *repair : RepairArgs{}, epi ? *epi : EpilogueChain{}, tile_w);
    count_launch();
    EML_TRY(check_launch("tf32x3_tcgen05 (CTA pair)"));
    if (ws && !skip_reduce)
        EML_TRY(launch_splitk_reduce<float>(ws, used_splits, C, batch, M, N, Strides3{strideC, ldc, 1}, accumulate, stream));
    return EML_OK;
}

}  // namespace pair

int plane_map(CUtensorMap* map, const float* base, int64_t MNp, int64_t Kp, int64_t batch, int box_rows) {
    uint64_t dims[3] = {(uint64_t)Kp, (uint64_t)MNp, (uint64_t)batch};
    uint64_t strides[2] = {(uint64_t)Kp * 4, (uint64_t)MNp * (uint64_t)Kp * 4};
    uint32_t box[3] = {(uint32_t)BK, (uint32_t)box_rows, 1};
    return make_tensor_map(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box,
                           CU_TENSOR_MAP_SWIZZLE_128B);
}

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

template <int BN>
int launch_tile(const Maps& maps, float* C, int64_t batch, int64_t M, int64_t N, int64_t ldc, int64_t strideC,
                int kblocks, int splits, int accumulate /* ACC_*; ACC_ZERO_PLUS only with ws */, float* ws, cudaStream_t stream) {
    auto kern = tf32x3_kernel<BN>;
    static thread_local int configured_dev = -1;
    int dev = 0;
    EML_CUDA(cudaGetDevice(&dev));
    if (configured_dev != dev) {
        EML_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<BN>::SMEM_BYTES));
        configured_dev = dev;
    }
    int64_t tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    int per_split = (kblocks + splits - 1) / splits;
    int used_splits = (kblocks + per_split - 1) / per_split;  // every split owns >= 1 k block
    dim3 grid((unsigned)tiles, (unsigned)batch, (unsigned)used_splits);
    launch_k(kern, dim3(grid), dim3(GEMM_THREADS), Cfg<BN>::SMEM_BYTES, stream, maps, C, M, N, ldc, strideC, kblocks, per_split,
                                                              accumulate == ACC_ADD ? 1 : 0, ws, batch * M * N);
    count_launch();
    EML_TRY(check_launch("tf32x3_tcgen05"));
    if (ws) EML_TRY(launch_splitk_reduce<float>(ws, used_splits, C, batch, M, N, Strides3{strideC, ldc, 1}, accumulate, stream));
    return EML_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// split/pack cache for immutable operands (see common.h)
// ------------------------------------------------------------------------------------------------
namespace {
std::atomic<uint64_t> g_next_identity{1};

struct PackKey {
    uint64_t id;
    int device;
    cudaStream_t stream;
    int64_t mn, k, mnp, kp, batch, sb, smn, sk;
    bool operator<(const PackKey& o) const {
        return std::tie(id, device, stream, mn, k, mnp, kp, batch, sb, smn, sk) <
               std::tie(o.id, o.device, o.stream, o.mn, o.k, o.mnp, o.kp, o.batch, o.sb, o.smn, o.sk);
    }
};
struct PackEntry {
    float* planes;  // big plane followed by the small plane
    size_t bytes;
};
std::mutex g_pack_mu;
std::map<PackKey, PackEntry> g_pack;
std::deque<PackKey> g_pack_order;  // insertion order, for eviction
size_t g_pack_bytes = 0;
constexpr size_t PACK_CACHE_LIMIT = size_t(4) << 30;

void pack_evict_locked(const PackKey& key) {
    auto it = g_pack.find(key);
    if (it == g_pack.end()) return;
    stream_free(it->second.planes, it->second.bytes, key.stream);
    g_pack_bytes -= it->second.bytes;
    g_pack.erase(it);
}

// big / small planes of one operand: from the cache when `id` is set (packing on a miss), else into `scratch`
// *flag: where this operand's "non-finite input seen" word lives -- the stream's shared word for a scratch pack (the
// repair kernel clears it), a word behind the planes for a cached operand (it describes the cached planes for as long
// as they live)
int packed_operand(uint64_t id, const float* src, int64_t batch, int64_t MN, int64_t K, int64_t MNp, int64_t Kp,
                   int64_t sb, int64_t smn, int64_t sk, TempBuf& scratch, cudaStream_t stream, float** big,
                   float** small, unsigned* stream_flag, const unsigned** flag) {
    const size_t plane = (size_t)(batch * MNp * Kp);
    const size_t bytes = sizeof(float) * 2 * plane + 256;  // + the cached operand's flag word
    if (id == 0) {
        EML_TRY(scratch.alloc(bytes, stream));
        *big = scratch.as<float>();
        *small = *big + plane;
        *flag = stream_flag;
        return launch_split_pack(src, *big, *small, batch, MN, K, MNp, Kp, sb, smn, sk, stream_flag, stream);
    }
    int dev = 0;
    EML_CUDA(cudaGetDevice(&dev));
    PackKey key{id, dev, stream, MN, K, MNp, Kp, batch, sb, smn, sk};
    {
        std::lock_guard<std::mutex> lock(g_pack_mu);
        auto it = g_pack.find(key);
        if (it != g_pack.end()) {
            *big = it->second.planes;
            *small = *big + plane;
            *flag = reinterpret_cast<const unsigned*>(*big + 2 * plane);
            return EML_OK;
        }
    }
    void* p = nullptr;
    EML_TRY(stream_alloc(&p, bytes, stream));
    *big = static_cast<float*>(p);
    *small = *big + plane;
    unsigned* own_flag = reinterpret_cast<unsigned*>(*big + 2 * plane);
    *flag = own_flag;
    EML_CUDA(cudaMemsetAsync(own_flag, 0, sizeof(unsigned), stream));
    EML_TRY(launch_split_pack(src, *big, *small, batch, MN, K, MNp, Kp, sb, smn, sk, own_flag, stream));
    std::lock_guard<std::mutex> lock(g_pack_mu);
    while (g_pack_bytes + bytes > PACK_CACHE_LIMIT && !g_pack_order.empty()) {
        pack_evict_locked(g_pack_order.front());
        g_pack_order.pop_front();
    }
    g_pack[key] = PackEntry{*big, bytes};
    g_pack_order.push_back(key);
    g_pack_bytes += bytes;
    return EML_OK;
}

// ------------------------------------------------------------------------------------------------
// "Direct" operands of the CTA-pair kernel: the tensor core reads the caller's f32 matrix itself as the big part
// (kind::tf32 truncates the low 13 mantissa bits, which is how tf32_split.cuh defines big), in place, whichever of
// its two indexes is contiguous (K-major / MN-major tiles, see sm100.cuh).  Only the SMALL parts are materialised:
// one plane with the matrix's own layout -- an elementwise pass that reads 4 and writes 4 bytes per element (the
// full pack: 4 + 8, through a transposing tile when the operand is m / n-contiguous).  Needs what TMA needs: a
// 16-byte aligned base, a leading dimension (and batch stride) that is a multiple of 4 elements, rows that do not
// overlap.  Everything else takes the packed planes.  EML_SGEMM_DIRECT=0 switches it off (comparison path: same bits).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
small_plane_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t rows, int64_t cols, int64_t ld_src,
                   int64_t ld_dst, int64_t sb_src, int64_t sb_dst, unsigned* nonfinite) {
    pdl_prologue();
    const int64_t vec_per_row = ld_dst / 4;  // ld_dst is a multiple of 4 >= cols
    const int64_t total = rows * vec_per_row;
    const float* s = src + (int64_t)blockIdx.y * sb_src;
    float* d = dst + (int64_t)blockIdx.y * sb_dst;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / vec_per_row, c = (i % vec_per_row) * 4;
        const float4 v = load4_guarded(s + r * ld_src + c, c, cols);
        float4 hi, lo;
        if (tf32::ordinary4(v)) {
            tf32::split4_ordinary(v, hi, lo);
        } else {
            tf32::split4_any(v, hi, lo);
            *nonfinite = 1u;
        }
        *reinterpret_cast<float4*>(d + r * ld_dst + c) = lo;
    }
}

struct DirectOperand {
    bool kin = true;         // k contiguous (K-major tiles) or mn contiguous (MN-major tiles)
    int64_t rows = 0, cols = 0, ld = 0, ld_small = 0;  // as stored: kin: rows = MN, cols = K; else rows = K, cols = MN
};

bool direct_enabled() {  // (read per call: the tests compare both paths in one process)
    const char* e = getenv("EML_SGEMM_DIRECT");
    return !(e && e[0] == '0');
}

// element (b, mn, k) of the operand at src[b*sb + mn*smn + k*sk]
bool direct_possible(const float* src, int64_t batch, int64_t MN, int64_t K, int64_t sb, int64_t smn, int64_t sk, DirectOperand* out) {
    if (!direct_enabled()) return false;
    const bool kin = sk == 1 && (smn != 1 || MN == 1), min = smn == 1 && sk != 1;
    if (!kin && !min) return false;
    DirectOperand d;
    d.kin = kin;
    d.rows = kin ? MN : K;
    d.cols = kin ? K : MN;
    d.ld = kin ? smn : sk;
    if ((reinterpret_cast<uintptr_t>(src) & 15) || (d.ld & 3) || d.ld < d.cols) return false;
    if (batch > 1 && ((sb & 3) || sb < (d.rows - 1) * d.ld + d.cols)) return false;
    if (d.ld > (int64_t(1) << 36) || MN > 0x7fffffffLL || K > 0x7fffffffLL) return false;
    d.ld_small = (d.cols + 3) / 4 * 4;
    *out = d;
    return true;
}

// tensor map of a direct operand (the caller's matrix, or its small plane with leading dimension ld)
int direct_map(CUtensorMap* map, const float* base, const DirectOperand& d, int64_t ld, int64_t batch, int64_t sb) {
    uint64_t dims[3] = {(uint64_t)d.cols, (uint64_t)d.rows, (uint64_t)batch};
    uint64_t strides[2] = {(uint64_t)ld * 4, (uint64_t)(batch > 1 ? sb : 4) * 4};
    uint32_t box[3] = {(uint32_t)(d.kin ? BK : 32), (uint32_t)(d.kin ? pair::HALF : BK), 1};
    return make_tensor_map(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box,
                           d.kin ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
}

int launch_small_plane(const float* src, float* dst, const DirectOperand& d, int64_t batch, int64_t sb, unsigned* nonfinite,
                       cudaStream_t stream) {
    const int64_t total = d.rows * (d.ld_small / 4);
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    launch_k(small_plane_kernel, dim3((unsigned)blocks, (unsigned)batch), dim3(256), 0, stream, src, dst, d.rows, d.cols, d.ld, d.ld_small,
             sb, d.rows * d.ld_small, nonfinite);
    count_launch();
    return check_launch("tf32x3 small-part plane");
}

// the small plane of a direct operand: from the cache when `id` is set (computing it on a miss), else into `scratch`
int small_operand(uint64_t id, const float* src, const DirectOperand& d, int64_t batch, int64_t sb, TempBuf& scratch,
                  cudaStream_t stream, float** small, unsigned* stream_flag, const unsigned** flag) {
    const size_t plane = (size_t)(batch * d.rows * d.ld_small);
    const size_t bytes = sizeof(float) * plane + 256;  // + the cached operand's flag word
    if (id == 0) {
        EML_TRY(scratch.alloc(bytes, stream));
        *small = scratch.as<float>();
        *flag = stream_flag;
        return launch_small_plane(src, *small, d, batch, sb, stream_flag, stream);
    }
    int dev = 0;
    EML_CUDA(cudaGetDevice(&dev));
    // (kind of entry: mnp = -1 marks a small-only plane with the operand's own layout)
    PackKey key{id, dev, stream, d.rows, d.cols, -1, d.ld_small, batch, sb, d.ld, d.kin ? 1 : 0};
    {
        std::lock_guard<std::mutex> lock(g_pack_mu);
        auto it = g_pack.find(key);
        if (it != g_pack.end()) {
            *small = it->second.planes;
            *flag = reinterpret_cast<const unsigned*>(*small + plane);
            return EML_OK;
        }
    }
    void* p = nullptr;
    EML_TRY(stream_alloc(&p, bytes, stream));
    *small = static_cast<float*>(p);
    unsigned* own_flag = reinterpret_cast<unsigned*>(*small + plane);
    *flag = own_flag;
    {
        ZeroRanges z;
        z.count = 1, z.p[0] = own_flag, z.bytes[0] = 16;
        EML_TRY(launch_zero_ranges(z, stream));
    }
    EML_TRY(launch_small_plane(src, *small, d, batch, sb, own_flag, stream));
    std::lock_guard<std::mutex> lock(g_pack_mu);
    while (g_pack_bytes + bytes > PACK_CACHE_LIMIT && !g_pack_order.empty()) {
        pack_evict_locked(g_pack_order.front());
        g_pack_order.pop_front();
    }
    g_pack[key] = PackEntry{*small, bytes};
    g_pack_order.push_back(key);
    g_pack_bytes += bytes;
    return EML_OK;
}
}  // namespace

// Reserves (or finds) the small-part plane that a later product will look up for the dense row-major matrix `src`
// [rows x cols] under identity `id` when it reads it in place as its B operand (k = row index, n contiguous), so
// that the kernel which PRODUCES src can write the plane itself, element i of the plane next to element i of src --
// a weight update, the sweep kernel of an activation -- and the product finds it in the cache instead of launching
// the small-part pass.  *plane = nullptr: not possible for this matrix (cols not a multiple of 4, alignment, ...).
// The flag word behind the plane is the reader's "non-finite input seen" word: the producer must set it.
int tf32x3_reserve_small_plane_b(uint64_t id, const float* src, int64_t rows, int64_t cols, cudaStream_t stream, float** plane,
                                 unsigned** flag) {
    *plane = nullptr;
    if (flag) *flag = nullptr;
    DirectOperand d;
    if (!id || (cols & 3) || !direct_possible(src, 1, cols, rows, 0, 1, cols, &d) || d.kin || d.ld_small != cols) return EML_OK;
    int dev = 0;
    EML_CUDA(cudaGetDevice(&dev));
    PackKey key{id, dev, stream, d.rows, d.cols, -1, d.ld_small, 1, 0, d.ld, 0};
    const size_t n = (size_t)(d.rows * d.ld_small);
    const size_t bytes = sizeof(float) * n + 256;
    std::lock_guard<std::mutex> lock(g_pack_mu);
    auto it = g_pack.find(key);
    if (it == g_pack.end()) {
        void* p = nullptr;
        EML_TRY(stream_alloc(&p, bytes, stream));
        while (g_pack_bytes + bytes > PACK_CACHE_LIMIT && !g_pack_order.empty()) {
            pack_evict_locked(g_pack_order.front());
            g_pack_order.pop_front();
        }
        it = g_pack.emplace(key, PackEntry{static_cast<float*>(p), bytes}).first;
        g_pack_order.push_back(key);
        g_pack_bytes += bytes;
    }
    *plane = it->second.planes;
    if (flag) *flag = reinterpret_cast<unsigned*>(*plane + n);
    return EML_OK;
}

uint64_t new_buffer_identity() { return g_next_identity.fetch_add(1); }
void pack_cache_drop(uint64_t id) {
    if (!id) return;
    std::lock_guard<std::mutex> lock(g_pack_mu);
    for (auto it = g_pack.begin(); it != g_pack.end();) {
        if (it->first.id == id) {
            stream_free(it->second.planes, it->second.bytes, it->first.stream);
            g_pack_bytes -= it->second.bytes;
            it = g_pack.erase(it);
        } else {
            ++it;
        }
    }
}
void pack_cache_release_all() {
    std::lock_guard<std::mutex> lock(g_pack_mu);
    for (auto& kv : g_pack) stream_free(kv.second.planes, kv.second.bytes, kv.first.stream);
    g_pack.clear();
    g_pack_order.clear();
    g_pack_bytes = 0;
}

int launch_gemm_tf32x3_f32(DeviceCtx* ctx, const float* A, const float* B, float* C, int64_t batch, int64_t M,
                           int64_t N, int64_t K, Strides3 sA, Strides3 sB, Strides3 sC, bool accumulate,
                           cudaStream_t stream, bool* handled, GemmOptions* opt) {
    *handled = false;
    if (batch == 0 || M == 0 || N == 0) {
        *handled = true;
        return EML_OK;
    }
    if (sC.c != 1 || batch > 65535) return EML_OK;  // C rows must be contiguous (the caller transposes otherwise)
    const char* forced = getenv("EML_SGEMM_KERNEL");
    if (forced && std::string(forced) == "simt") return EML_OK;
    *handled = true;

    const int sms = ctx ? ctx->sm_count : 148;
    const int64_t Kp = round_up(K, BK);
    const int kblocks = (int)(Kp / BK);
    const int64_t tiles_m = (M + BM - 1) / BM;
    // CTA-pair persistent kernel for outputs wider and taller than one 128-row tile
    const bool use_pair = M > 128 && N > 128 && !(forced && std::string(forced) == "1cta");
    // C = G * G^T (requested by contract_dev): tiles on and above the diagonal only, and one packed copy of G
    const bool upper_only = opt && opt->upper_only && use_pair && M == N && !accumulate;
    int BN, splits = 1;
    int64_t Mp, Np;
    if (use_pair) {
        BN = pair::PN;
        Mp = round_up(M, pair::PM);
        Np = round_up(N, pair::PN);
        const int64_t tm_ = Mp / pair::PM;
        const int64_t tiles = (upper_only ? tm_ * (tm_ + 1) / 2 : tm_ * (Np / pair::PN)) * batch;
        const int clusters = sms / 2;
        if (tiles * 2 <= clusters && kblocks >= 8) {  // too few tiles for 74 SM pairs: deterministic split-K
            splits = (int)(clusters / tiles);
            if (splits > kblocks / 4) splits = kblocks / 4;
            if (splits < 1) splits = 1;
        }
    } else {
        // tile width: 256 when that still fills the machine, else 128
        const int64_t tiles256 = tiles_m * ((N + 255) / 256) * batch;
        BN = (N > 128 && tiles256 * 10 >= (int64_t)sms * 8) ? 256 : 128;
        const int64_t tiles = tiles_m * ((N + BN - 1) / BN) * batch;
        // deterministic split-K when the output has too few tiles for 148 SMs (e.g. dW = X^T dH, K = batch size)
        if (tiles * 2 <= sms && kblocks >= 8) {
            splits = (int)((sms + tiles - 1) / tiles);
            if (splits > kblocks / 4) splits = kblocks / 4;
            if (splits < 1) splits = 1;
        }
        // bound the truncating TMEM accumulation chain (see pair::FLUSH_KB): at most 8 k blocks per split
        const int min_splits = (kblocks + pair::FLUSH_KB - 1) / pair::FLUSH_KB;
        if (splits < min_splits) splits = min_splits;
        Mp = round_up(M, BM);
        Np = round_up(N, BN);
    }

    const uint64_t a_id = opt ? opt->a_id : 0, b_id = opt ? opt->b_id : 0;
    // Small outputs: the pack pass moves 12 (M + N) K bytes at HBM speed while the GEMM does 2 M N K flops, so its
    // share grows as M N / (M + N) shrinks (a third of the time for batched 1024^3).  There the variant that splits
    // inside the kernel wins although its main loop is slower (shared-memory-bandwidth bound, measured 240 vs
    // 280 TFLOP/s):
    //   2 M N K (1/240e12 - 1/280e12)  <  12 (M + N) K / 6e12      <=>      M N / (M + N) < 1680
    // (measured break-even: 4096^2, where M N / (M + N) = 2048).  It needs >= 24 k blocks per tile to amortise its
    // unoverlapped epilogue, and is pointless when a packed operand is cached (identity hints from the tape).
    // EML_SGEMM_FUSED=1 forces it wherever it can run, =0 disables it.
    // ACC_ZERO_PLUS (GemmOptions::zero_plus): a split-K launch of the packed kernels ends in the reduction, which writes
    // 0 + sum; a launch whose epilogue adds into C gets its zeros first
    const bool zero_plus = accumulate && opt && opt->zero_plus;
    bool zeroed = false;
    auto zero_c = [&]() -> int {
        zeroed = true;
        const size_t bytes = sizeof(float) * (size_t)(M * N);
        if (((reinterpret_cast<uintptr_t>(C) | bytes) & 15) == 0) {
            ZeroRanges z;
            z.count = 1, z.p[0] = C, z.bytes[0] = bytes;
            return launch_zero_ranges(z, stream);
        }
        EML_CUDA(cudaMemsetAsync(C, 0, bytes, stream));
        return EML_OK;
    };
    if (use_pair && !upper_only) {
        const char* fe = getenv("EML_SGEMM_FUSED");
        const bool pays = a_id == 0 && b_id == 0 && kblocks >= 24 * splits && M * N < 1680 * (M + N);
        const bool want = fe ? fe[0] == '1' : pays;
        if (want) {
            bool fused = false;
            unsigned* tk = nullptr;
            EML_TRY(stream_tickets(stream, &tk));
            if (zero_plus) EML_TRY(zero_c());
            EML_TRY(launch_gemm_tf32x3_fused(ctx, A, B, C, batch, M, N, K, sA, sB, sC, splits, accumulate, stream, &fused,
                                             tk + TICKET_NONFINITE));
            if (fused) {
                if (!accumulate)
                    EML_TRY(launch_tf32_repair(A, B, C, batch, M, N, K, sA, sB, sC, nullptr, nullptr, tk + TICKET_NONFINITE, stream));
                return EML_OK;
            }
        }
    }
    // ---- tile width of the CTA-pair kernel (one batch, no epilogue extras) ----
    // The pair kernel's M tile is fixed (2 x 128 rows); its N tile is whatever the MMA's N field says, a multiple of 16
    // up to 256.  When the output is a few waves of work, narrower tiles can mean less MMA issue time for the busiest
    // CTA pair: 300 columns are two tiles of 160 instead of 256 + 44 (zero padding multiplied for nothing), 784 columns
    // four of 208.  For outputs of ONE wave, the width with the least  k blocks x width  wins when it beats 256 by 8 % WITHOUT changing
    // the split-K partition -- every element's accumulation chain is then what it was, so the bits are those of the
    // plain tiling (and of the in-kernel-split kernel, which always uses it).  Measured: 1000 x 300 x 2048 30.1 -> 26.2 us,
    // 4096 x 784 x 512 32.6 -> 30.6 us (scripts/probe/balance_ab.py).  EML_SGEMM_BALANCE=0: always 256.
    // (Also tried: running C^T = B^T A^T when the ragged extent is M -- X^T dZ of the NN step pads 784 rows to 1024 --
    //  with the tile stored transposed.  4 and 3 us faster for the step's two big products on their own, 5 us SLOWER
    //  inside the step, and different bits than the in-kernel-split kernel; not kept.)
    int tile_w = pair::PN;
    if (use_pair && !upper_only && batch == 1 && !(opt && opt->epilogue)) {
        const char* be = getenv("EML_SGEMM_BALANCE");
        if (!(be && be[0] == '0')) {
            const int clusters = sms / 2;
            auto plan = [&](int w, int* sp_out) {
                const int64_t tiles = ((M + pair::PM - 1) / pair::PM) * ((N + w - 1) / w);
                int sp = 1;
                if (tiles * 2 <= clusters && kblocks >= 8) {
                    sp = (int)(clusters / tiles);
                    if (sp > kblocks / 4) sp = kblocks / 4;
                    if (sp < 1) sp = 1;
                }
                const int per = (kblocks + sp - 1) / sp, used = (kblocks + per - 1) / per;
                const int64_t waves = (tiles * used + clusters - 1) / clusters;
                *sp_out = sp;
                // one wave only: with several tiles per CTA pair the narrower ones lose more to operand re-reads than
                // they save in padding (4096^3 with 240-column tiles: 553 -> 700 us)
                return waves == 1 ? (double)per * w : 1e30;
            };
            int sp0 = 1;
            double best = plan(pair::PN, &sp0) * 0.92;
            for (int w = pair::PN - 16; w >= 128; w -= 16) {
                int sp = 1;
                const double c = plan(w, &sp);
                if (sp == sp0 && c < best) best = c, tile_w = w;
            }
        }
    }
    int acc_mode = accumulate ? ACC_ADD : ACC_WRITE;
    if (zero_plus && !zeroed) {
        if (splits > 1) acc_mode = ACC_ZERO_PLUS;
        else EML_TRY(zero_c());
    }
    TempBuf a_scratch, b_scratch, wsbuf;
    float *a_big = nullptr, *a_small = nullptr, *b_big = nullptr, *b_small = nullptr;
    unsigned* tickets = nullptr;
    EML_TRY(stream_tickets(stream, &tickets));
    unsigned* stream_flag = tickets + TICKET_NONFINITE;
    const unsigned *a_flag = nullptr, *b_flag = nullptr;
    // CTA-pair kernel: an operand TMA can read in place is not packed -- the matrix itself is the big part (see
    // DirectOperand); the 128-row kernels and operands TMA cannot describe take the packed planes
    DirectOperand da, db;
    const bool a_direct = use_pair && direct_possible(A, batch, M, K, sA.b, sA.r, sA.c, &da);
    // B element (k, n) at B[k*sB.r + n*sB.c]: its "mn" index is n
    const bool b_direct = use_pair && direct_possible(B, batch, N, K, sB.b, sB.c, sB.r, &db);
    if (a_direct) {
        EML_TRY(small_operand(a_id, A, da, batch, sA.b, a_scratch, stream, &a_small, stream_flag, &a_flag));
    } else {
        EML_TRY(packed_operand(a_id, A, batch, M, K, Mp, Kp, sA.b, sA.r, sA.c, a_scratch, stream, &a_big, &a_small, stream_flag, &a_flag));
    }
    if (upper_only && a_direct == b_direct) {
        // B is the same matrix read through mirrored strides: the plane(s) just made serve both sides
        b_big = a_big;
        b_small = a_small;
        b_flag = a_flag;
    } else if (b_direct) {
        EML_TRY(small_operand(b_id, B, db, batch, sB.b, b_scratch, stream, &b_small, stream_flag, &b_flag));
    } else {
        EML_TRY(packed_operand(b_id, B, batch, N, K, Np, Kp, sB.b, sB.c, sB.r, b_scratch, stream, &b_big, &b_small, stream_flag, &b_flag));
    }
    float* ws = nullptr;
    if (splits > 1) {
        EML_TRY(wsbuf.alloc(sizeof(float) * (size_t)splits * (size_t)(batch * M * N), stream));
        ws = wsbuf.as<float>();
    }

    Maps maps;
    const int b_box = use_pair ? pair::HALF : BN;  // rows of B one TMA box brings in
    bool maps_ok = true;
    if (a_direct) {
        maps_ok = direct_map(&maps.a_big, A, da, da.ld, batch, sA.b) == EML_OK &&
                  direct_map(&maps.a_small, a_small, da, da.ld_small, batch, da.rows * da.ld_small) == EML_OK;
    } else {
        EML_TRY(plane_map(&maps.a_big, a_big, Mp, Kp, batch, BM));
        EML_TRY(plane_map(&maps.a_small, a_small, Mp, Kp, batch, BM));
    }
    if (b_direct) {
        maps_ok = maps_ok && direct_map(&maps.b_big, B, db, db.ld, batch, sB.b) == EML_OK &&
                  direct_map(&maps.b_small, b_small, db, db.ld_small, batch, db.rows * db.ld_small) == EML_OK;
    } else {
        EML_TRY(plane_map(&maps.b_big, b_big, Np, Kp, batch, b_box));
        EML_TRY(plane_map(&maps.b_small, b_small, Np, Kp, batch, b_box));
    }
    if (!maps_ok) {
        set_error("tf32x3: the driver refused a tensor map of an operand read in place (set EML_SGEMM_DIRECT=0)");
        return EML_ERR_CUDA;
    }
    const bool a_kin = a_direct ? da.kin : true, b_kin = b_direct ? db.kin : true;
    set_last_kernel("tf32x3_tcgen05");
    bool repair_folded = false;
    if (use_pair) {
        pair::RepairArgs rep;
        if (!accumulate && !upper_only && !ws) {  // the kernel stores C itself: the repair rides in its epilogue
            rep.A = A, rep.B = B, rep.K = K;
            rep.sAb = sA.b, rep.sAr = sA.r, rep.sAc = sA.c, rep.sBb = sB.b, rep.sBr = sB.r, rep.sBc = sB.c;
            rep.flag_a = a_flag == stream_flag ? nullptr : a_flag;
            rep.flag_b = b_flag == stream_flag ? nullptr : b_flag;
            rep.stream_flag = stream_flag;
            rep.ticket = tickets;
            repair_folded = true;
        }
        // an elementwise chain handed over by the tape: only the launch that stores C itself, one batch, the whitelisted
        // chain (anything else: the caller runs the group as its own kernel)
        const EpilogueChain* epi = nullptr;
        if (opt && opt->epilogue && !accumulate && !upper_only && !ws && batch == 1) {
            const EpilogueChain& e = *opt->epilogue;
            if (e.n_ops == 4 && e.op[0] == EML_OP_NEG && e.op[1] == EML_OP_EXP && e.op[2] == EML_OP_ADD_C && e.op[3] == EML_OP_C_DIV) {
                epi = &e;
                opt->epilogue_done = true;
            }
        }
        // the caller adds the partial sums itself (GemmOptions::defer_reduce): a first contribution to a dense C whose
        // reduction would be the sequential kernel
        const int used_splits = (kblocks + (kblocks + splits - 1) / splits - 1) / ((kblocks + splits - 1) / splits);
        const bool defer = ws && opt && opt->defer_reduce && acc_mode != ACC_ADD && !upper_only && batch == 1 && sC.r == N &&
                           used_splits > 1 && splitk_reduce_is_sequential(M * N, used_splits);
        EML_TRY(pair::launch(maps, C, batch, M, N, sC.r, sC.b, kblocks, splits, sms, acc_mode, ws, stream, upper_only,
                             repair_folded ? &rep : nullptr, epi, a_kin, b_kin, defer, tile_w));
        if (defer) {
            opt->deferred_ws = wsbuf.p, opt->deferred_bytes = wsbuf.bytes;
            opt->deferred_splits = used_splits, opt->deferred_stride = M * N;
            wsbuf.p = nullptr;  // (ownership passes to the caller)
        }
        if (upper_only) opt->upper_done = true;
    } else if (BN == 256) {
        EML_TRY(launch_tile<256>(maps, C, batch, M, N, sC.r, sC.b, kblocks, splits, acc_mode, ws, stream));
    } else {
        EML_TRY(launch_tile<128>(maps, C, batch, M, N, sC.r, sC.b, kblocks, splits, acc_mode, ws, stream));
    }
    // (an accumulating launch has already folded the product into C's previous contents: nothing to recompute from;
    //  the upper-triangle launch is mirrored afterwards by the caller, and repaired NaNs stay symmetric only if both
    //  halves are recomputed: run the repair on the full matrix after the mirror is not possible here, so skip it)
    if (!accumulate && !upper_only && !repair_folded)
        EML_TRY(launch_tf32_repair(A, B, C, batch, M, N, K, sA, sB, sC, a_flag, b_flag, stream_flag, stream));
    return EML_OK;
}

namespace {
__global__ void __launch_bounds__(256)
tf32_repair_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int64_t batch, int64_t M,
                   int64_t N, int64_t K, Strides3 sA, Strides3 sB, Strides3 sC, const unsigned* flag_a,
                   const unsigned* flag_b, unsigned* stream_flag, unsigned* ticket) {
    pdl_prologue();
    // every block reads the flags, then draws a ticket; the block drawing the last one clears the stream's shared word
    // (all the others have read it by then)
    __shared__ unsigned seen;
    if (threadIdx.x == 0) {
        seen = ((flag_a && *flag_a) || (flag_b && *flag_b) || *stream_flag) ? 1u : 0u;
        __threadfence();
        if (atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1) *stream_flag = 0u;
    }
    __syncthreads();
    if (!seen) return;
    const int64_t total = batch * M * N;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = e / (M * N), i = (e / N) % M, j = e % N;
        float* c = C + b * sC.b + i * sC.r + j * sC.c;
        if (*c == *c) continue;  // only NaNs can be artefacts of the split (0 * inf in a cross term, inf - inf between terms)
        const float* a = A + b * sA.b + i * sA.r;
        const float* bb = B + b * sB.b + j * sB.c;
        float acc = 0.0f;
        for (int64_t k = 0; k < K; k++) acc = fmaf(a[k * sA.c], bb[k * sB.r], acc);
        *c = acc;
    }
}
}  // namespace

int launch_tf32_repair(const float* A, const float* B, float* C, int64_t batch, int64_t M, int64_t N, int64_t K, Strides3 sA,
                       Strides3 sB, Strides3 sC, const unsigned* flag_a, const unsigned* flag_b, unsigned* stream_flag,
                       cudaStream_t stream) {
    int64_t blocks = (batch * M * N + 255) / 256;
    if (blocks > 148) blocks = 148;  // nothing to do in the common case; the rare repair is a slow path anyway
    unsigned* tickets = nullptr;
    EML_TRY(stream_tickets(stream, &tickets));
    launch_k(tf32_repair_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, A, B, C, batch, M, N, K, sA, sB, sC,
             flag_a == stream_flag ? nullptr : flag_a, flag_b == stream_flag ? nullptr : flag_b, stream_flag, tickets);
    count_launch();
    return check_launch("tf32 non-finite repair");
}

}  // namespace eml
