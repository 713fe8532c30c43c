// gemm_tf32x3_fused.cu -- 3xTF32 f32 GEMM with the big/small split done INSIDE the kernel (no split/pack pass).
//
// The split/pack pass of gemm_tf32x3.cu is HBM-bound; it costs 6 % at K = 8192 but a third of the time at K = 1024
// (BASELINE config 3: batch 64, 1024^3).  Here the raw f32 operand tiles are brought in by TMA, fourteen
// warps split them in shared memory into the big / small TF32 tiles (K-major, 128-byte swizzled: exactly what TMA
// with SWIZZLE_128B would have written from packed planes) and the same tcgen05.mma.cta_group::2 pipeline consumes
// them.  Shared-memory traffic per 32-wide k block and SM rises from 160 KB (packed planes in by TMA, MMA operand
// reads) to 224 KB (raw in by TMA, raw read, big/small written, MMA reads): 1792 cycles of the 128 B/clk shared
// memory pipe against 1536 cycles of MMA issue, so the tensor pipe cannot exceed 86 % here (measured 75 %, with the
// shared-memory pipe 87 % busy).  That is why the dispatcher only uses this kernel where not running the pack pass
// buys more than the slower main loop costs (small outputs), and keeps the packed path elsewhere.
//
// Per CTA of a pair:    warp 0      one lane issues the TMA loads of the raw tiles (3-stage ring)
//                       warp 1      TMEM allocation; leader CTA: one lane issues the MMAs (M=256, N=256, K=8)
//                       warps 2,3,12..15   converters
//                       warps 4..11 converters too, and the epilogue: between two conversions they drain a finished
//                                   512-k accumulator chunk into a register-resident running sum (see gemm_tf32x3.cu)
// Raw tile of an operand whose k index is contiguous in memory: TMA box (32 k, 128 rows), SWIZZLE_128B -- the
// converter is a layout-preserving elementwise pass.  Otherwise (m / n contiguous): box (128 rows, 32 k), no swizzle,
// i.e. [k][row] in shared memory; a converter lane owns one row, gathers four consecutive k and writes one 16-byte
// chunk at the swizzled position (conflict free on both sides).
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
constexpr int RAW_STAGE_BYTES = 2 * TILE_BYTES;  // A raw, B raw
constexpr int OP_STAGE_BYTES = 4 * TILE_BYTES;   // A big, A small, B big, B small
constexpr int RAW_STAGES = 3, OP_STAGES = 2, ACC_STAGES = 2;
constexpr uint32_t TMEM_COLS = 512;
constexpr size_t SMEM_BYTES =
    (size_t)RAW_STAGES * RAW_STAGE_BYTES + (size_t)OP_STAGES * OP_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int CONV_WARPS = 14, EPI_WARPS = 8, THREADS = 16 * 32;
constexpr int LIGHT_REGS = 64, EPILOGUE_REGS = 192;  // 256 light + 256 epilogue threads: 65536 registers
constexpr int FLUSH_KB = 16;

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

// One k block: the raw A and B tiles -> big / small tiles.  64 warp tasks (a task = 32 lanes x 16 bytes of one
// tile); converter warp `cw` takes tasks cw, cw + CONV_WARPS, ...  The loads of a group of tasks are issued before
// the first split so that their latencies overlap.
constexpr int TASKS = 64, MAX_TASKS = (TASKS + CONV_WARPS - 1) / CONV_WARPS, GROUP_TASKS = 3;
template <bool A_KIN, bool B_KIN>
__device__ __forceinline__ void convert_block(const uint8_t* __restrict__ raw, uint8_t* __restrict__ op, int cw,
                                              int lane, unsigned* nonfinite) {
#pragma unroll
    for (int g = 0; g < MAX_TASKS; g += GROUP_TASKS) {
        float4 v[GROUP_TASKS];
#pragma unroll
        for (int i = 0; i < GROUP_TASKS; i++) {
            const int t = cw + CONV_WARPS * (g + i);
            if (g + i >= MAX_TASKS || t >= TASKS) continue;
            const bool is_b = t >= 32;
            const bool kin = is_b ? B_KIN : A_KIN;
            const int tt = t & 31;
            const uint8_t* src = raw + (is_b ? TILE_BYTES : 0);
            if (kin) {
                // the tile is already [row][32 k] 128-byte swizzled: elementwise, layout preserving
                v[i] = *reinterpret_cast<const float4*>(src + (tt * 32 + lane) * 16);
            } else {
                // the tile is [32 k][128 rows]; a lane owns row 32*(tt/8) + lane and gathers k = 4*k4 .. 4*k4+3
                const float* rf = reinterpret_cast<const float*>(src);
                const int row = (tt >> 3) * 32 + lane, k4 = tt & 7;
                v[i] = make_float4(rf[(4 * k4 + 0) * HALF + row], rf[(4 * k4 + 1) * HALF + row],
                                   rf[(4 * k4 + 2) * HALF + row], rf[(4 * k4 + 3) * HALF + row]);
            }
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
            const bool is_b = t >= 32;
            const bool kin = is_b ? B_KIN : A_KIN;
            const int tt = t & 31;
            float4 hi, lo;
            if (ordinary)
                tf32::split4_ordinary(v[i], hi, lo);
            else
                tf32::split4_any(v[i], hi, lo);
            int off;
            if (kin) {
                off = (tt * 32 + lane) * 16;
            } else {
                const int row = (tt >> 3) * 32 + lane, k4 = tt & 7;
                off = row * 128 + ((k4 ^ (row & 7)) << 4);  // SWIZZLE_128B position of chunk k4 of this row
            }
            uint8_t* dst = op + (is_b ? 2 * TILE_BYTES : 0);  // A big, A small, B big, B small
            *reinterpret_cast<float4*>(dst + off) = hi;
            *reinterpret_cast<float4*>(dst + TILE_BYTES + off) = lo;
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
    uint8_t* raw_ring = base;
    uint8_t* op_ring = base + (size_t)RAW_STAGES * RAW_STAGE_BYTES;
    uint64_t* raw_full = reinterpret_cast<uint64_t*>(op_ring + (size_t)OP_STAGES * OP_STAGE_BYTES);
    uint64_t* raw_empty = raw_full + RAW_STAGES;
    uint64_t* op_full = raw_empty + RAW_STAGES;
    uint64_t* op_empty = op_full + OP_STAGES;
    uint64_t* acc_full = op_empty + OP_STAGES;
    uint64_t* acc_empty = acc_full + ACC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + ACC_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster = blockIdx.x >> 1, clusters = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        prefetch_tensormap(&maps.a);
        prefetch_tensormap(&maps.b);
        for (int s = 0; s < RAW_STAGES; s++) {
            mbar_init(&raw_full[s], 1);
            mbar_init(&raw_empty[s], CONV_WARPS);
        }
        for (int s = 0; s < OP_STAGES; s++) {
            mbar_init(&op_full[s], 2 * CONV_WARPS);  // leader's copy: every converter warp of both CTAs
            mbar_init(&op_empty[s], 1);
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
            constexpr uint32_t idesc = umma_idesc_tf32(PM, PN);
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
                        const int s = n % OP_STAGES;
                        mbar_wait(&op_full[s], (n / OP_STAGES) & 1);
                        tc_fence_after();
                        const uint32_t st = smem_u32(op_ring + (size_t)s * OP_STAGE_BYTES);
                        const uint64_t a_big = umma_desc_k_sw128(st), a_small = umma_desc_k_sw128(st + TILE_BYTES);
                        const uint64_t b_big = umma_desc_k_sw128(st + 2 * TILE_BYTES);
                        const uint64_t b_small = umma_desc_k_sw128(st + 3 * TILE_BYTES);
                        const uint32_t first = kb != c0;
#pragma unroll
                        for (int k = 0; k < BK / 8; k++)
                            umma_tf32_pair(d, a_small + 2 * k, b_big + 2 * k, idesc, first | (uint32_t)k);
#pragma unroll
                        for (int k = 0; k < BK / 8; k++) umma_tf32_pair(d, a_big + 2 * k, b_small + 2 * k, idesc, 1);
#pragma unroll
                        for (int k = 0; k < BK / 8; k++) umma_tf32_pair(d, a_big + 2 * k, b_big + 2 * k, idesc, 1);
                        umma_commit_pair(&op_empty[s], 0b11);
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
                    const int s = n % RAW_STAGES;
                    if (n >= RAW_STAGES) mbar_wait(&raw_empty[s], ((n / RAW_STAGES) - 1) & 1);
                    uint8_t* st = raw_ring + (size_t)s * RAW_STAGE_BYTES;
                    mbar_expect_tx(&raw_full[s], RAW_STAGE_BYTES);
                    const int kc = kb * BK;
                    if (A_KIN)
                        tma_load_3d(st, &maps.a, &raw_full[s], kc, row_a, wk.batch);
                    else
                        tma_load_3d(st, &maps.a, &raw_full[s], row_a, kc, wk.batch);
                    if (B_KIN)
                        tma_load_3d(st + TILE_BYTES, &maps.b, &raw_full[s], kc, row_b, wk.batch);
                    else
                        tma_load_3d(st + TILE_BYTES, &maps.b, &raw_full[s], row_b, kc, wk.batch);
                }
            }
        }
    } else {
        // ---------------- converters: raw ring -> big / small operand ring ----------------
        // warps 2, 3, 12..15 do nothing else; the epilogue warps 4..11 convert too and drain an accumulator
        // chunk two k blocks after its last block was converted (its MMAs have completed by then: the operand
        // ring is two stages deep, so the conversion of block n + 2 waits for the MMAs of block n anyway)
        int total = 0;  // k blocks this cluster processes, over all its work items
        for (int w = cluster; w < work_items; w += clusters) {
            const int kb0 = (w % splits) * per_split;
            total += min(kblocks_total, kb0 + per_split) - kb0;
        }
        auto convert_step = [&](int n, int cw) {
            const int rs = n % RAW_STAGES, os = n % OP_STAGES;
            mbar_wait(&raw_full[rs], (n / RAW_STAGES) & 1);
            if (n >= OP_STAGES) mbar_wait(&op_empty[os], ((n / OP_STAGES) - 1) & 1);
            convert_block<A_KIN, B_KIN>(raw_ring + (size_t)rs * RAW_STAGE_BYTES, op_ring + (size_t)os * OP_STAGE_BYTES,
                                        cw, lane, nonfinite);
            fence_proxy_async();  // this lane's tile writes become visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&raw_empty[rs]);
                mbar_arrive_cluster(&op_full[os], 0);
            }
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

// raw operand tensor map.  KIN: element (row, k) at base[b*sb + row*ld + k]; else at base[b*sb + k*ld + row]
int raw_map(CUtensorMap* map, const float* base, bool kin, int64_t rows, int64_t K, int64_t ld, int64_t batch,
            int64_t sb) {
    uint64_t dims[3], strides[2];
    uint32_t box[3];
    if (kin) {
        dims[0] = (uint64_t)K; dims[1] = (uint64_t)rows;
        box[0] = BK; box[1] = HALF;
    } else {
        dims[0] = (uint64_t)rows; dims[1] = (uint64_t)K;
        box[0] = HALF; box[1] = BK;
    }
    dims[2] = (uint64_t)batch;
    box[2] = 1;
    strides[0] = (uint64_t)ld * 4;
    strides[1] = (uint64_t)(batch > 1 ? sb : 4) * 4;
    return make_tensor_map(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box,
                           kin ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE);
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
