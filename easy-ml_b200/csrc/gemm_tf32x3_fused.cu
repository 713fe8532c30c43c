// gemm_tf32x3_fused.cu -- 3xTF32 f32 GEMM with the big/small split done INSIDE the kernel (no split/pack pass).
//
// The split/pack pass of gemm_tf32x3.cu is HBM-bound; it costs 6 % at K = 8192 but a third of the time at K = 1024
// (BASELINE config 3: batch 64, 1024^3).  Here TMA brings in the RAW f32 operand tiles and the tensor core reads them
// as they are: tcgen05.mma.kind::tf32 ignores the low 13 mantissa bits of an operand word, i.e. it sees
// big = trunc_tf32(x) (tf32_split.cuh; measured, scripts/probe/tf32_operand_probe.cu).  Fourteen warps only compute the
// SMALL tiles, small = rna(x - big), elementwise and layout preserving, and the same tcgen05.mma.cta_group::2 pipeline
// forms As Bb + Ab Bs + Ab Bb with Ab, Bb read from the raw tiles.
//
// Operand layouts in shared memory (both understood by TMA and by the MMA's shared-memory descriptors):
//   k index contiguous in memory ("K-major"):  [128 rows][32 k], 128-byte rows, SWIZZLE_128B; one TMA box (32 k, 128 rows)
//   m / n index contiguous ("MN-major"):       [4 groups][32 k][32 rows], 128-byte rows of 32 consecutive m / n,
//       CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B <-> UMMA layout type 1 (SWIZZLE_128B_BASE32B: 32-byte units XOR (k & 3),
//       groups of 4 k at SBO = 512 B, the next 32 rows at LBO = 4096 B) -- the only MN-major form kind::tf32 accepts
//       (the probe: every other layout type reads zeros); four TMA boxes (32 rows, 32 k) per tile.
//   So an n-contiguous B (the ordinary row-major right operand) needs no transposing conversion any more.
//
// Shared-memory pipe per 32-wide k block and SM: 32 KB of raw tiles written by TMA + 32 KB read by the converters +
// 32 KB of small tiles written + 96 KB read by the MMAs = 192 KB = 1536 cycles of the 128 B/clk pipe -- the same as the
// 1536 cycles of MMA issue (round 1's version wrote big tiles too: 224 KB = 1792 cycles, tensor pipe 66 % active).
//
// Per CTA of a pair:    warp 0      one lane issues the TMA loads of the raw tiles (3-stage ring; a stage holds the raw
//                                   A and B tiles and their small tiles, and is freed by the commit of its MMAs)
//                       warp 1      TMEM allocation; leader CTA: one lane issues the MMAs (M=256, N=256, K=8)
//                       warps 2,3,12..15   converters
//                       warps 4..11 converters too, and the epilogue: between two conversions they drain a finished
//                                   512-k accumulator chunk into a register-resident running sum (see gemm_tf32x3.cu)
// Ragged M, N, K need no padding: TMA zero-fills out-of-range box elements.
#include <cstdlib>
#include <string>

#include "common.h"
#include "sm100.cuh"
#include "tf32_split.cuh"

namespace eml {

namespace {

using namespace sm100;

namespace fz {

constexpr int PM = 256, PN = 256, HALF = 128, BK = 32;
constexpr int TILE_BYTES = HALF * BK * 4;        // 16 KB
constexpr int STAGE_BYTES = 4 * TILE_BYTES;      // A raw, B raw, A small, B small
constexpr int STAGES = 3, ACC_STAGES = 2;
constexpr uint32_t TMEM_COLS = 512;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int CONV_WARPS = 14, EPI_WARPS = 8, THREADS = 16 * 32;
constexpr int LIGHT_REGS = 64, EPILOGUE_REGS = 192;  // 256 light + 256 epilogue threads: 65536 registers
constexpr int FLUSH_KB = 8;  // as in gemm_tf32x3.cu (pair::FLUSH_KB): the two kernels must chunk alike to give the same bits
constexpr int GROUP_BYTES = 32 * BK * 4;         // MN-major: one group of 32 rows x 32 k = 4 KB

struct Maps {
    CUtensorMap a, b;
};

struct Work {
    int tm, tn, batch, split;
};
__device__ __forceinline__ Work decode(int w, int tiles_m, int tiles_n, int splits) {
    Work k;
    k.split = w % splits;
    int t = w / splits;
    const int tiles = tiles_m * tiles_n;
    k.batch = t / tiles;
    t %= tiles;
    constexpr int GROUP = 8;
    const int per_group = GROUP * tiles_n;
    const int first_m = (t / per_group) * GROUP;
    const int rows = tiles_m - first_m < GROUP ? tiles_m - first_m : GROUP;
    const int in_group = t % per_group;
    k.tm = first_m + in_group % rows;
    k.tn = in_group / rows;
    return k;
}

template <bool KIN>
__device__ __forceinline__ uint64_t tile_desc(uint32_t tile_addr, int ks) { return umma_tile_desc_f32<KIN>(tile_addr, ks); }

// One k block: the raw A and B tiles -> their small tiles, elementwise and layout preserving (whatever the layout).
// 64 warp tasks (a task = 32 lanes x 16 bytes of one tile); converter warp `cw` takes tasks cw, cw + CONV_WARPS, ...
// The loads of a group of tasks are issued before the first split so that their latencies overlap.
// A vector holding an Inf / NaN / >= 2^127 value (rare) also rewrites its raw words with the big parts of
// tf32::split_special, so that the packed path and this one feed the tensor core the same numbers.
constexpr int TASKS = 64, MAX_TASKS = (TASKS + CONV_WARPS - 1) / CONV_WARPS, GROUP_TASKS = 3;
__device__ __forceinline__ void convert_block(uint8_t* __restrict__ stage, int cw, int lane, unsigned* nonfinite) {
#pragma unroll
    for (int g = 0; g < MAX_TASKS; g += GROUP_TASKS) {
        float4 v[GROUP_TASKS];
#pragma unroll
        for (int i = 0; i < GROUP_TASKS; i++) {
            const int t = cw + CONV_WARPS * (g + i);
            if (g + i >= MAX_TASKS || t >= TASKS) continue;
            v[i] = *reinterpret_cast<const float4*>(stage + (t >> 5) * TILE_BYTES + ((t & 31) * 32 + lane) * 16);
        }
        // one test for the whole group: the common path below it is branch free, so the group's twelve
        // independent split chains interleave (a per-vector branch serialised them: measured 0.15 IPC per warp)
        bool ordinary = true;
#pragma unroll
        for (int i = 0; i < GROUP_TASKS; i++) {
            const int t = cw + CONV_WARPS * (g + i);
            if (g + i >= MAX_TASKS || t >= TASKS) continue;
            ordinary &= tf32::ordinary4(v[i]);
        }
        if (!ordinary) *nonfinite = 1u;  // the repair kernel after the product recomputes the NaNs the split can cause
#pragma unroll
        for (int i = 0; i < GROUP_TASKS; i++) {
            const int t = cw + CONV_WARPS * (g + i);
            if (g + i >= MAX_TASKS || t >= TASKS) continue;
            const int off = (t >> 5) * TILE_BYTES + ((t & 31) * 32 + lane) * 16;
            float4 hi, lo;
            if (ordinary) {
                tf32::split4_ordinary(v[i], hi, lo);
            } else {
                tf32::split4_any(v[i], hi, lo);
                *reinterpret_cast<float4*>(stage + off) = hi;
            }
            *reinterpret_cast<float4*>(stage + 2 * TILE_BYTES + off) = lo;
        }
    }
}

template <bool A_KIN, bool B_KIN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
tf32x3_fused_kernel(const __grid_constant__ Maps maps, float* __restrict__ C, int64_t M, int64_t N, int64_t ldc,
                    int64_t strideC, int kblocks_total, int per_split, int splits, int tiles_m, int tiles_n,
                    int work_items, int accumulate, float* __restrict__ ws, int64_t ws_split_stride, unsigned* nonfinite) {
    pdl_prologue();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* ring = base;
    uint64_t* raw_full = reinterpret_cast<uint64_t*>(ring + (size_t)STAGES * STAGE_BYTES);
    uint64_t* conv_full = raw_full + STAGES;
    uint64_t* empty = conv_full + STAGES;
    uint64_t* acc_full = empty + STAGES;
    uint64_t* acc_empty = acc_full + ACC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + ACC_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster = blockIdx.x >> 1, clusters = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        prefetch_tensormap(&maps.a);
        prefetch_tensormap(&maps.b);
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&raw_full[s], 1);                // this CTA's producer arrives, this CTA's bytes land
            mbar_init(&conv_full[s], 2 * CONV_WARPS);  // leader's copy: every converter warp of both CTAs
            mbar_init(&empty[s], 1);                   // one multicast commit per use
        }
        for (int a = 0; a < ACC_STAGES; a++) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], 2 * EPI_WARPS);
        }
        fence_barrier_init();
    }
    __syncwarp();
    if (warp == 1) tmem_alloc_pair(tmem_slot, TMEM_COLS);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    const bool is_epilogue = warp >= 4 && warp < 12;
    if (!is_epilogue) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(LIGHT_REGS));

    if (warp == 1) {
        // ---------------- MMA issuer (leader CTA only) ----------------
        if (rank == 0 && lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(PM, PN) | (A_KIN ? 0u : UMMA_IDESC_A_MN) | (B_KIN ? 0u : UMMA_IDESC_B_MN);
            int n = 0, it = 0;
            for (int w = cluster; w < work_items; w += clusters) {
                const Work wk = decode(w, tiles_m, tiles_n, splits);
                const int kb0 = wk.split * per_split, kb1 = min(kblocks_total, kb0 + per_split);
                for (int c0 = kb0; c0 < kb1; c0 += FLUSH_KB, it++) {
                    const int c1 = min(kb1, c0 + FLUSH_KB);
                    const int as = it & 1;
                    if (it >= ACC_STAGES) mbar_wait(&acc_empty[as], ((it >> 1) - 1) & 1);
                    tc_fence_after();
                    const uint32_t d = tmem_acc + (uint32_t)as * PN;
                    for (int kb = c0; kb < c1; kb++, n++) {
                        const int s = n % STAGES;
                        mbar_wait(&conv_full[s], (n / STAGES) & 1);
                        tc_fence_after();
                        const uint32_t st = smem_u32(ring + (size_t)s * STAGE_BYTES);
                        const uint32_t a_big = st, b_big = st + TILE_BYTES, a_small = st + 2 * TILE_BYTES,
                                       b_small = st + 3 * TILE_BYTES;
                        const uint32_t first = kb != c0;
#pragma unroll
                        for (int k = 0; k < BK / 8; k++)
                            umma_tf32_pair(d, tile_desc<A_KIN>(a_small, k), tile_desc<B_KIN>(b_big, k), idesc, first | (uint32_t)k);
#pragma unroll
                        for (int k = 0; k < BK / 8; k++)
                            umma_tf32_pair(d, tile_desc<A_KIN>(a_big, k), tile_desc<B_KIN>(b_small, k), idesc, 1);
#pragma unroll
                        for (int k = 0; k < BK / 8; k++)
                            umma_tf32_pair(d, tile_desc<A_KIN>(a_big, k), tile_desc<B_KIN>(b_big, k), idesc, 1);
                        umma_commit_pair(&empty[s], 0b11);  // frees the stage (raw and small tiles) in both CTAs
                    }
                    umma_commit_pair(&acc_full[as], 0b11);
                }
            }
        }
    } else if (warp == 0) {
        // ---------------- TMA producer (both CTAs): raw f32 tiles of this CTA's 128 rows of A and of B ----------------
        if (lane == 0) {
            int n = 0;
            for (int w = cluster; w < work_items; w += clusters) {
                const Work wk = decode(w, tiles_m, tiles_n, splits);
                const int kb0 = wk.split * per_split, kb1 = min(kblocks_total, kb0 + per_split);
                const int row_a = wk.tm * PM + (int)rank * HALF, row_b = wk.tn * PN + (int)rank * HALF;
                for (int kb = kb0; kb < kb1; kb++, n++) {
                    const int s = n % STAGES;
                    if (n >= STAGES) mbar_wait(&empty[s], ((n / STAGES) - 1) & 1);
                    uint8_t* st = ring + (size_t)s * STAGE_BYTES;
                    mbar_expect_tx(&raw_full[s], 2 * TILE_BYTES);
                    const int kc = kb * BK;
                    if (A_KIN) {
                        tma_load_3d(st, &maps.a, &raw_full[s], kc, row_a, wk.batch);
                    } else {
#pragma unroll
                        for (int g = 0; g < 4; g++) tma_load_3d(st + g * GROUP_BYTES, &maps.a, &raw_full[s], row_a + 32 * g, kc, wk.batch);
                    }
                    if (B_KIN) {
                        tma_load_3d(st + TILE_BYTES, &maps.b, &raw_full[s], kc, row_b, wk.batch);
                    } else {
#pragma unroll
                        for (int g = 0; g < 4; g++)
                            tma_load_3d(st + TILE_BYTES + g * GROUP_BYTES, &maps.b, &raw_full[s], row_b + 32 * g, kc, wk.batch);
                    }
                }
            }
        }
    } else {
        // ---------------- converters: small tiles of the stage the TMA has just filled ----------------
        // warps 2, 3, 12..15 do nothing else; the epilogue warps 4..11 convert too and drain an accumulator
        // chunk two k blocks after its last block was converted (its MMAs have completed by then: the conversion of
        // block n + 3 waits for the MMAs of block n anyway, through the producer's wait on the stage)
        int total = 0;  // k blocks this cluster processes, over all its work items
        for (int w = cluster; w < work_items; w += clusters) {
            const int kb0 = (w % splits) * per_split;
            total += min(kblocks_total, kb0 + per_split) - kb0;
        }
        auto convert_step = [&](int n, int cw) {
            const int s = n % STAGES;
            mbar_wait(&raw_full[s], (n / STAGES) & 1);
            convert_block(ring + (size_t)s * STAGE_BYTES, cw, lane, nonfinite);
            fence_proxy_async();  // this lane's tile writes become visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&conv_full[s], 0);
        };
        if (!is_epilogue) {
            const int cw = warp < 4 ? warp - 2 : warp - 10;  // 0..5
            for (int n = 0; n < total; n++) convert_step(n, cw);
        } else {
            asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(EPILOGUE_REGS));
            const int cw = warp + 2;                          // 6..13
            const int q = warp & 3, half = (warp - 4) >> 2;   // TMEM lane quarter, 128-column half
            // cursor over accumulator chunks: chunk [c0, min(kb1, c0 + FLUSH_KB)) of work item cur_w, whose first
            // k block has index n_base in this cluster's sequence
            int cur_w = cluster, n_base = 0, kb0 = 0, kb1 = 0, c0 = 0, it = 0;
            Work wk{};
            auto load_item = [&]() {
                if (cur_w < work_items) {
                    wk = decode(cur_w, tiles_m, tiles_n, splits);
                    kb0 = wk.split * per_split;
                    kb1 = min(kblocks_total, kb0 + per_split);
                    c0 = kb0;
                }
            };
            load_item();
            float sum[128];  // this thread's row, 128 columns: the IEEE-rounded running sum over k chunks
            for (int n = 0; n < total + 2; n++) {
                const int c1 = min(kb1, c0 + FLUSH_KB);
                if (cur_w < work_items && n == n_base + (c1 - kb0) + 1) {
                    // ---- drain the chunk whose last block was converted two steps ago ----
                    const int as = it & 1;
                    mbar_wait(&acc_full[as], (it >> 1) & 1);
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
                    it++;
                    if (c1 < kb1) {
                        c0 = c1;
                    } else {
                        // ---- the work item is complete: one store of the tile ----
                        const int64_t row = (int64_t)wk.tm * PM + (int64_t)rank * HALF + q * 32 + lane;
                        const int64_t n0 = (int64_t)wk.tn * PN + half * 128;
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
                                if (col0 >= N) continue;
                                float* p = out + row * ld + col0;
                                if (vec_ok && col0 + 32 <= N) {
#pragma unroll
                                    for (int j = 0; j < 8; j++) {
                                        float4 o = make_float4(sum[c * 32 + 4 * j], sum[c * 32 + 4 * j + 1],
                                                               sum[c * 32 + 4 * j + 2], sum[c * 32 + 4 * j + 3]);
                                        if (add) {
                                            float4 old = *reinterpret_cast<const float4*>(p + 4 * j);
                                            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                                        }
                                        *reinterpret_cast<float4*>(p + 4 * j) = o;
                                    }
                                } else {
#pragma unroll
                                    for (int j = 0; j < 32; j++)
                                        if (col0 + j < N) p[j] = add ? p[j] + sum[c * 32 + j] : sum[c * 32 + j];
                                }
                            }
                        }
                        n_base += kb1 - kb0;
                        cur_w += clusters;
                        load_item();
                    }
                }
                if (n < total) convert_step(n, cw);
            }
        }
    }

    __syncwarp();
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_pair(tmem_acc, TMEM_COLS);
}

// raw operand tensor map.  KIN: element (row, k) at base[b*sb + row*ld + k], box (32 k, 128 rows), SWIZZLE_128B;
// else at base[b*sb + k*ld + row], box (32 rows, 32 k), 128-byte swizzle with 32-byte atoms (see the header)
int raw_map(CUtensorMap* map, const float* base, bool kin, int64_t rows, int64_t K, int64_t ld, int64_t batch,
            int64_t sb) {
    uint64_t dims[3], strides[2];
    uint32_t box[3];
    if (kin) {
        dims[0] = (uint64_t)K; dims[1] = (uint64_t)rows;
        box[0] = BK; box[1] = HALF;
    } else {
        dims[0] = (uint64_t)rows; dims[1] = (uint64_t)K;
        box[0] = 32; box[1] = BK;
    }
    dims[2] = (uint64_t)batch;
    box[2] = 1;
    strides[0] = (uint64_t)ld * 4;
    strides[1] = (uint64_t)(batch > 1 ? sb : 4) * 4;
    return make_tensor_map(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box,
                           kin ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
}

template <bool A_KIN, bool B_KIN>
int launch(const Maps& maps, float* C, int64_t batch, int64_t M, int64_t N, int64_t ldc, int64_t strideC, int kblocks,
           int splits, int sms, bool accumulate, float* ws, cudaStream_t stream, unsigned* nonfinite) {
    auto kern = tf32x3_fused_kernel<A_KIN, B_KIN>;
    static thread_local int configured_dev = -1;
    int dev = 0;
    EML_CUDA(cudaGetDevice(&dev));
    if (configured_dev != dev) {
        EML_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        configured_dev = dev;
    }
    const int tiles_m = (int)((M + PM - 1) / PM), tiles_n = (int)((N + PN - 1) / PN);
    const int per_split = (kblocks + splits - 1) / splits;
    const int used_splits = (kblocks + per_split - 1) / per_split;
    const int64_t items = (int64_t)batch * tiles_m * tiles_n * used_splits;
    if (items > 0x7fffffffLL) EML_BAD_ARG("tf32x3: too many work items");
    int clusters = sms / 2;
    if (items < clusters) clusters = (int)items;
    EML_CUDA(launch_k(kern, dim3(2 * clusters), dim3(THREADS), SMEM_BYTES, stream, maps, C, M, N, ldc, strideC, kblocks,
                      per_split, used_splits, tiles_m, tiles_n, (int)items, accumulate ? 1 : 0, ws, batch * M * N, nonfinite));
    count_launch();
    EML_TRY(check_launch("tf32x3_tcgen05 (fused split)"));
    if (ws) EML_TRY(launch_splitk_reduce<float>(ws, used_splits, C, batch, M, N, Strides3{strideC, ldc, 1}, accumulate, stream));
    return EML_OK;
}

}  // namespace fz

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

// Returns with *handled = false when the operands do not meet TMA's alignment rules (the packed path takes over).
int launch_gemm_tf32x3_fused(DeviceCtx* ctx, const float* A, const float* B, float* C, int64_t batch, int64_t M,
                             int64_t N, int64_t K, Strides3 sA, Strides3 sB, Strides3 sC, int splits, bool accumulate,
                             cudaStream_t stream, bool* handled, unsigned* nonfinite) {
    *handled = false;
    const bool a_kin = sA.c == 1, a_min = sA.r == 1;
    const bool b_nin = sB.c == 1, b_kin = sB.r == 1;
    if (!(a_kin || a_min) || !(b_nin || b_kin) || sC.c != 1 || batch > 65535) return EML_OK;
    const bool A_KIN = a_kin, B_KIN = !b_nin;
    const int64_t lda = A_KIN ? sA.r : sA.c, ldb = B_KIN ? sB.c : sB.r;
    if (!al16(A) || !al16(B) || (lda & 3) || (ldb & 3) || (batch > 1 && ((sA.b & 3) || (sB.b & 3)))) return EML_OK;
    if (lda < (A_KIN ? K : M) || ldb < (B_KIN ? K : N)) return EML_OK;  // overlapping rows (a broadcast): pack path
    // a broadcast operand (batch stride 0) or overlapping batches cannot be described to TMA: the pack pass can
    if (batch > 1 && (sA.b < (A_KIN ? (M - 1) * lda + K : (K - 1) * lda + M) ||
                      sB.b < (B_KIN ? (N - 1) * ldb + K : (K - 1) * ldb + N)))
        return EML_OK;
    fz::Maps maps;
    // (a descriptor the driver refuses -- extents or strides outside its limits -- also means "not this kernel")
    if (fz::raw_map(&maps.a, A, A_KIN, M, K, lda, batch, sA.b) != EML_OK) return EML_OK;
    if (fz::raw_map(&maps.b, B, B_KIN, N, K, ldb, batch, sB.b) != EML_OK) return EML_OK;
    const int kblocks = (int)((K + fz::BK - 1) / fz::BK);
    TempBuf wsbuf;
    float* ws = nullptr;
    if (splits > 1) {
        EML_TRY(wsbuf.alloc(sizeof(float) * (size_t)splits * (size_t)(batch * M * N), stream));
        ws = wsbuf.as<float>();
    }
    const int sms = ctx ? ctx->sm_count : 148;
    *handled = true;
    set_last_kernel("tf32x3_tcgen05");
    if (A_KIN && !B_KIN) return fz::launch<true, false>(maps, C, batch, M, N, sC.r, sC.b, kblocks, splits, sms, accumulate, ws, stream, nonfinite);
    if (A_KIN && B_KIN) return fz::launch<true, true>(maps, C, batch, M, N, sC.r, sC.b, kblocks, splits, sms, accumulate, ws, stream, nonfinite);
    if (!A_KIN && !B_KIN) return fz::launch<false, false>(maps, C, batch, M, N, sC.r, sC.b, kblocks, splits, sms, accumulate, ws, stream, nonfinite);
    return fz::launch<false, true>(maps, C, batch, M, N, sC.r, sC.b, kblocks, splits, sms, accumulate, ws, stream, nonfinite);
}

}  // namespace eml
